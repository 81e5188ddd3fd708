"""GPU parity tests, model level: encoder / cross-KV / teacher-forced decoder with QK capture against the CPU oracle
(oracle.whisper_ref, fp32) on random-init weights at reduced and real Whisper shapes.  Tolerances: logits within 1e-3
relative (BASELINE.json north_star); observed errors are printed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dims_or_name, seed, precision="fp16x3", n_samples=300000, n_tok=40):
    import oracle.whisper_ref as W
    from oracle import stable_path as SP
    from stable_ts_b200.model import from_oracle
    model = W.build_model(dims_or_name, seed=seed)
    tk = W.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en",
                                   task="transcribe")
    gm = from_oracle(model, precision=precision)
    audio = SP.synth_audio(n_samples, seed=1234)
    script = SP.synth_token_script(n_tok, tk.eot, seed=4321)
    return W, SP, model, tk, gm, audio, script


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("case", ["mini", "tiny.en", "base"])
def test_forward_matches_oracle(case):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.whisper_ref.model import ModelDimensions
    dims = ModelDimensions(80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2) if case == "mini" else case
    W, SP, model, tk, gm, audio, script = _setup(dims, seed=11)
    n = len(audio)
    mel_ref = W.pad_or_trim(W.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000 - n), 3000)
    mel = gm.log_mel(audio.cuda()[None])
    e_mel = (mel[0].cpu() - mel_ref).abs().max().item()
    with torch.no_grad():
        xa_ref, qks_ref, logits_ref, probs_ref = SP.window_qks(model, tk, script, mel_ref)
    enc = gm.encode(mel)
    e_xa = _rel(enc["f32"][0], xa_ref[0])
    ckv = gm.cross_kv(enc)
    row = torch.tensor([SP.alignment_token_row(tk, script)], dtype=torch.int32)
    logits, qk = gm.decode_forced(row, ckv, heads="all")
    torch.cuda.synchronize()
    e_logits = _rel(logits[0], logits_ref)
    H = model.dims.n_text_head
    e_qk = max(_rel(qk[0, l * H:(l + 1) * H, :, :1500], qks_ref[l][0]) for l in range(model.dims.n_text_layer))
    print(f"[{case}] mel abs {e_mel:.2e} | xa rel {e_xa:.2e} | logits rel {e_logits:.2e} | qk rel {e_qk:.2e}")
    assert e_mel < 2e-4 and e_xa < 1e-3 and e_logits < 1e-3 and e_qk < 1e-3
    # token probabilities (timing.py:62-64)
    S = len(tk.sot_sequence)
    p, _ = gm.token_probs(logits[0, S:S + len(script)], tk.eot, torch.tensor(script))
    np.testing.assert_allclose(p.cpu().numpy(), np.array(probs_ref), rtol=2e-3)
    # argmax agreement of the teacher-forced rows (proxy for "token ids bit-exact at temperature 0")
    assert torch.equal(logits[0].argmax(-1).cpu(), logits_ref.argmax(-1))


def test_forward_batched_and_fp16_mode():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    W, SP, model, tk, gm, audio, script = _setup("tiny", seed=5)
    from stable_ts_b200.model import from_oracle
    a2 = torch.stack([audio, SP.synth_audio(len(audio), seed=77)])
    mel = gm.log_mel(a2.cuda())
    enc = gm.encode(mel)
    ckv = gm.cross_kv(enc)
    rows = torch.tensor([SP.alignment_token_row(tk, script), SP.alignment_token_row(tk, script[::-1])], dtype=torch.int32)
    logits, qk = gm.decode_forced(rows, ckv, heads=gm.alignment_head_pairs)
    torch.cuda.synchronize()
    for b in range(2):
        mel_ref = W.pad_or_trim(W.log_mel_spectrogram(a2[b], 80, padding=480000 - a2.shape[1]), 3000)
        with torch.no_grad():
            _, qks_ref, logits_ref, _ = SP.window_qks(model, tk, rows[b, 4:-1].tolist(), mel_ref)
        e = _rel(logits[b], logits_ref)
        eq = max(_rel(qk[b, i, :, :1500], qks_ref[l][0, h]) for i, (l, h) in enumerate(gm.alignment_head_pairs))
        print(f"batched item {b}: logits rel {e:.2e} qk rel {eq:.2e}")
        assert e < 1e-3 and eq < 1e-3
    g16 = from_oracle(model, precision="fp16")
    enc16 = g16.encode(mel)
    l16, _ = g16.decode_forced(rows, g16.cross_kv(enc16))
    torch.cuda.synchronize()
    e16 = _rel(l16[0], logits[0])
    print(f"fp16 single-pass vs fp16x3 logits rel {e16:.2e}")
    assert e16 < 5e-2
