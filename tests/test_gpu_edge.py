"""GPU edge cases the reference's own driver can produce: maximum token count (token_step =
n_text_ctx - 6, alignment.py:181-185), single-token windows, sub-second audio, ragged batches, and loud failures on
invalid arguments (no silent fallback)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(name, seed):
    import oracle.whisper_ref as W
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    model = W.build_model(name, seed=seed)
    gm = from_oracle(model)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    otk = W.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en", task="transcribe")
    return W, model, gm, tk, otk


def test_maximum_token_window_and_tiny_windows_in_one_batch():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import stable_path as SP
    from stable_ts_b200.alignment import align_words_batch
    W, model, gm, tk, otk = _mk("tiny", 21)
    audios = [SP.synth_audio(480000, seed=1), SP.synth_audio(8000, seed=2), SP.synth_audio(480000, seed=3)]
    scripts = [SP.synth_token_script(442, tk.eot, seed=5), SP.synth_token_script(1, tk.eot, seed=6),
               SP.synth_token_script(3, tk.eot, seed=7)]
    wts = [SP.words_from_script(s, seed=i) for i, s in enumerate(scripts)]
    got = align_words_batch(gm, tk, audios, wts)
    worst, bad, total = 0.0, 0, 0
    for a, wt, g in zip(audios, wts, got):
        ref = SP.align_audio_window(model, otk, wt, a)
        assert len(ref) == len(g)
        for r, w in zip(ref, g):
            d = max(abs(r["start"] - w["start"]), abs(r["end"] - w["end"]))
            worst, total, bad = max(worst, d), total + 1, bad + (d > 0.0201)
    print(f"max-size/tiny ragged batch: {total} words, worst |dt| {worst:.3f}s, outside +-20ms: {bad}")
    assert bad == 0


def test_refine_probs_short_audio():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import stable_path as SP
    from stable_ts_b200.alignment import refine_probs
    W, model, gm, tk, otk = _mk("tiny.en", 22)
    script = SP.synth_token_script(5, tk.eot, seed=1)
    a2 = torch.stack([SP.synth_audio(24000, seed=1), SP.synth_audio(24000, seed=2)])
    p_ref, r_ref = SP.prob_and_rank(SP.refine_token_probs(model, otk, a2, script), script)
    p, r = refine_probs(gm, tk, a2, script)
    np.testing.assert_allclose(p.cpu().numpy(), p_ref.numpy(), rtol=2e-3)
    assert (r.cpu().long() - r_ref).abs().max().item() <= 2


def test_invalid_arguments_fail_loudly():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200 import _lib as L
    W, model, gm, tk, otk = _mk("tiny.en", 23)
    lib = L.lib()
    enc = gm.encode(torch.zeros(1, 80, 3000))
    ckv = gm.cross_kv(enc)
    with pytest.raises(L.StbError):                                   # more tokens than n_text_ctx
        gm.decode_forced(torch.zeros(1, 449, dtype=torch.int32), ckv)
    x = torch.zeros(1, 500, 64, device="cuda")
    with pytest.raises(L.StbError):                                   # DTW rows beyond the supported 480
        gm.dtw(x)
    a = torch.zeros(8, 64, dtype=torch.float16, device="cuda")
    opa = L.Operand(L.ptr(a), None, 8, 64, 64, 0, 0)
    opb = L.Operand(L.ptr(a), None, 8, 32, 64, 0, 0)
    ep = L.Epilogue()
    out = torch.zeros(8, 8, device="cuda")
    ep.out_f32, ep.ld_out, ep.alpha = L.ptr(out), 8, 1.0
    with pytest.raises(L.StbError):                                   # K mismatch
        L.check(lib.stb_gemm(ctypes.byref(opa), ctypes.byref(opb), 1, 1, ctypes.byref(ep), L.stream_ptr()))
    assert b"K mismatch" in lib.stb_last_error()
