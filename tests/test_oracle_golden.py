"""The oracle (CPU restatement) must reproduce what the UNMODIFIED reference computed (tests/golden/*.npz, written
by oracle/make_golden.py in the build container).  CPU only; runs on the GPU box too (no /root/reference needed)."""
import os

import numpy as np
import pytest
import torch

import oracle.whisper_ref as W
from oracle import c_oracle
from oracle import stable_path as SP
from oracle.whisper_ref.model import ModelDimensions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLD, f"{name}.npz"))
    dims = ModelDimensions(*[int(v) for v in z["dims"]])
    model = W.build_model(dims, seed=int(z["model_seed"]))
    tk = W.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en",
                                   task="transcribe")
    script = z["script"].tolist()
    wts, i = [], 0
    for n in z["word_lens"]:
        wts.append(script[i:i + int(n)])
        i += int(n)
    audio = SP.synth_audio(int(z["n_samples"]), seed=1234)
    return z, model, tk, script, wts, audio


@pytest.mark.parametrize("name", ["mini_en", "mini_ml"])
def test_align_window_reproduces_reference(name):
    z, model, tk, script, wts, audio = load_case(name)
    out, inter = SP.align_audio_window(model, tk, wts, audio, return_intermediates=True)
    assert np.array_equal(inter["jumps"], z["jumps"])
    np.testing.assert_allclose(inter["matrix"].numpy(), z["matrix"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(inter["token_probs"], z["token_probs"], rtol=1e-6)
    assert np.array_equal([w["start"] for w in out], z["word_start"])
    assert np.array_equal([w["end"] for w in out], z["word_end"])
    np.testing.assert_allclose([w["probability"] for w in out], z["word_prob"], rtol=1e-6)


@pytest.mark.parametrize("name", ["mini_en", "mini_ml"])
def test_refine_and_decode_reproduce_reference(name):
    z, model, tk, script, wts, audio = load_case(name)
    a2 = torch.stack([audio, SP.synth_audio(int(z["n_samples"]), seed=99)])
    p, rank = SP.prob_and_rank(SP.refine_token_probs(model, tk, a2, script), script)
    np.testing.assert_allclose(p.numpy(), z["refine_p"], rtol=1e-5)
    assert np.array_equal(rank.numpy(), z["refine_rank"])
    mel = W.pad_or_trim(W.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000 - len(audio)), 3000)
    assert abs(float(mel.double().sum()) - float(z["mel_checksum"])) < 1e-6 * abs(float(z["mel_checksum"]))
    mask = torch.zeros(1501, dtype=torch.bool)
    mask[100:400] = True
    res, _, _ = SP.decode_window(model, mel, ts_token_mask=mask, language="en", sample_len=24)
    assert res.tokens == z["decode_tokens"].tolist()
    assert abs(res.avg_logprob - float(z["decode_avg_logprob"])) < 1e-5
    assert abs(res.no_speech_prob - float(z["decode_no_speech"])) < 1e-9


def test_dtw_goldens_python_and_c():
    z = np.load(os.path.join(GOLD, "dtw_cases.npz"))
    for i in range(3):
        x, p = z[f"x{i}"], z[f"p{i}"]
        mine = W.timing.dtw(torch.from_numpy(-x))
        assert np.array_equal(mine, p)
        cp, jumps = c_oracle.dtw(-x)
        assert np.array_equal(cp, p)
        assert np.array_equal(jumps, SP.jumps_from_matrix(torch.from_numpy(x)))


@pytest.mark.parametrize("name", ["mini_en", "mini_ml"])
def test_transcribe_window_reproduces_reference_driver(name):
    """SP.transcribe_window == first window of the unmodified transcribe_stable (tests/golden/*_transcribe.json)."""
    import json
    z, model, tk, script, wts, audio = load_case(name)
    gold = json.load(open(os.path.join(GOLD, f"{name}_transcribe.json")))
    segs, _ = SP.transcribe_window(model, tk, audio, language="en", sample_len=40)
    assert [s["tokens"] for s in segs] == [s["tokens"] for s in gold["segments"]]
    for s, g in zip(segs, gold["segments"]):
        assert s["start"] == g["start"] and s["end"] == g["end"]
        assert [(w["word"], w["tokens"], w["start"], w["end"]) for w in s["words"]] == \
               [(w["word"], w["tokens"], w["start"], w["end"]) for w in g["words"]]
        np.testing.assert_allclose([w["probability"] for w in s["words"]], [w["probability"] for w in g["words"]], rtol=1e-6)
