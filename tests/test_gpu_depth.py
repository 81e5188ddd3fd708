"""GPU parity AT THE BENCHMARKED DEPTH (VERDICT r1 "weak" item 1): the CUDA path against the CPU oracle on random-init
weights at the full `small` (12+12 layers, BASELINE config 3) and `large-v3` (32+32 layers, configs 4/5) shapes.

Gates (BASELINE.json north_star): logits / QK / encoder output within 1e-3 relative (max |diff| / max |ref|), teacher-
forced argmax rows equal, greedy token ids bit-exact, word start/end within +-20 ms, token probabilities 2e-3.
The same tests print what one fp16 tensor-core pass instead of three would cost at this depth, so the precision choices
of DESIGN.md section 3 rest on measurements at 32 layers, not on an extrapolation from 6.
CPU cost: one large-v3 window through the oracle is ~15-30 s on the GPU box's host cores."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def _build(name, seed, precision="fp16x3"):
    import oracle.whisper_ref as W
    from oracle import stable_path as SP
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    model = W.build_model(name, seed=seed)
    otk = W.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en", task="transcribe")
    gm = from_oracle(model, precision=precision)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    return W, SP, model, otk, gm, tk


@pytest.mark.parametrize("name", ["small", "large-v3"])
def test_full_depth_forward_matches_oracle(name):
    """a1-a4 at full depth: mel, encoder output, all-head cross-attention QK of every layer, logits, token probabilities."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200.model import from_oracle
    W, SP, model, otk, gm, tk = _build(name, seed=11)
    audio = SP.synth_audio(480000, seed=1234)
    script = SP.synth_token_script(60, otk.eot, seed=4321)
    mel_ref = W.pad_or_trim(W.log_mel_spectrogram(audio, model.dims.n_mels), 3000)
    with torch.no_grad():
        xa_ref, qks_ref, logits_ref, probs_ref = SP.window_qks(model, otk, script, mel_ref)
    row = torch.tensor([SP.alignment_token_row(otk, script)], dtype=torch.int32)
    H, Ld = model.dims.n_text_head, model.dims.n_text_layer
    errs = {}
    for prec in ("fp16x3", "fp16"):
        g = gm if prec == "fp16x3" else from_oracle(model, precision="fp16")
        mel = g.log_mel(audio.cuda()[None])
        enc = g.encode(mel)
        logits, qk = g.decode_forced(row, g.cross_kv(enc), heads="all")
        torch.cuda.synchronize()
        errs[prec] = dict(mel=(mel[0].cpu() - mel_ref).abs().max().item(), xa=_rel(enc["f32"][0], xa_ref[0]),
                          logits=_rel(logits[0], logits_ref),
                          qk=max(_rel(qk[0, l * H:(l + 1) * H, :, :1500], qks_ref[l][0]) for l in range(Ld)),
                          argmax_equal=bool(torch.equal(logits[0].argmax(-1).cpu(), logits_ref.argmax(-1))))
        if prec == "fp16x3":
            S = len(otk.sot_sequence)
            p, _ = g.token_probs(logits[0, S:S + len(script)], otk.eot, torch.tensor(script))
            p_err = float(np.max(np.abs(p.cpu().numpy() - np.array(probs_ref)) / np.array(probs_ref)))
        del logits, qk, enc, mel
        if g is not gm:
            del g
        torch.cuda.empty_cache()
    for prec, e in errs.items():
        print(f"[{name} depth {Ld}] {prec:7s}: mel abs {e['mel']:.2e} | xa rel {e['xa']:.2e} | logits rel {e['logits']:.2e} | "
              f"qk rel {e['qk']:.2e} | argmax rows equal {e['argmax_equal']}")
    print(f"[{name}] fp16x3 token-prob rel {p_err:.2e}")
    e = errs["fp16x3"]
    assert e["mel"] < 2e-4 and e["xa"] < 1e-3 and e["logits"] < 1e-3 and e["qk"] < 1e-3 and e["argmax_equal"]
    assert p_err < 2e-3
    assert errs["fp16"]["logits"] < 5e-2          # sanity bound only: the single-pass mode is NOT the parity mode


def test_large_v3_decode_step_vs_oracle():
    """a9 at 32 layers: forced 24-step script, per-step filtered logits and argmax vs the oracle for the decode step (fp16
    cross K/V planes) in the parity mode and, for the record, in one-pass fp16.  Round-2 measurement that decided the
    format (gpurun_out/r2_depth/depth.log): 3-byte hi+int8-residual K/V 2.07e-5, fp16-only K/V 2.15e-5, gate 1e-3."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200.decode import DecodingOptions, decode_windows
    from stable_ts_b200.model import from_oracle
    W, SP, model, otk, gm, tk = _build("large-v3", seed=6)
    audio = SP.synth_audio(480000, seed=41)
    steps = 24
    script = SP.synth_token_script(steps, otk.eot, seed=5)
    mel_ref = W.pad_or_trim(W.log_mel_spectrogram(audio, model.dims.n_mels), 3000)
    ref, _, ex = SP.decode_window(model, mel_ref, forced_tokens=script, return_step_logits=True, sample_len=steps, language="en")

    def run(g):
        enc = g.encode(g.log_mel(audio.cuda()[None]))
        res, gx = decode_windows(g, tk, enc, DecodingOptions(language="en", sample_len=steps),
                                 forced_tokens=torch.tensor(script)[:, None], return_step_logits=True)
        worst = 0.0
        for i in range(steps):
            r, o = ex["step_logits"][i], gx["step_logits"][i][0].cpu()
            fin = r > -1e30
            assert torch.equal(fin, o > -1e30), f"mask mismatch at step {i}"
            worst = max(worst, ((o[fin] - r[fin]).abs().max() / r[fin].abs().max()).item())
        return worst, gx["step_argmax"][:, 0].tolist() == ex["step_argmax"], res[0].avg_logprob

    out = {"fp16x3 (parity mode)": run(gm)}
    out["fp16 single pass"] = run(from_oracle(model, precision="fp16"))
    for k, (w, eq, lp) in out.items():
        print(f"[large-v3 decode step, {steps} steps] {k}: worst step-logit rel {w:.2e}, argmax bit-exact {eq}, "
              f"avg_logprob {lp:.5f} (oracle {ref.avg_logprob:.5f})")
    w, eq, lp = out["fp16x3 (parity mode)"]
    assert w < 1e-3 and eq and abs(lp - ref.avg_logprob) < 1e-3


@pytest.mark.parametrize("name,steps", [("small", 64), ("large-v3", 224)])
def test_full_depth_transcribe_window_matches_oracle(name, steps):
    """The bench workload on ONE window, GPU public path vs the oracle's transcribe_window: forced `steps`-token script
    (KV-cached decode, step argmax bit-exact), then gap-padded word timestamps (+-20 ms, probabilities 2e-3)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200.decode import DecodingOptions
    from stable_ts_b200.transcribe import transcribe_windows
    W, SP, model, otk, gm, tk = _build(name, seed=0)
    audio = SP.synth_audio(480000, seed=1000)
    script = SP.synth_token_script(steps, otk.eot, seed=4321 + 1000)
    ref_segs, ex = SP.transcribe_window(model, otk, audio, forced_tokens=script, sample_len=steps, language="en")
    segs, info = transcribe_windows(gm, tk, audio[None].contiguous(), forced_tokens=torch.tensor(script, dtype=torch.int32)[:, None],
                                    options=DecodingOptions(language="en", sample_len=steps, max_initial_timestamp=None))
    assert info["step_argmax"][:, 0].tolist() == ex["step_argmax"], "greedy token ids differ from the oracle"
    ref_words = [w for s in ref_segs for w in s["words"]]
    words = [w for s in segs[0] for w in s["words"]]
    assert len(words) == len(ref_words) and len(words) > 0
    worst_t = worst_p = 0.0
    for a, b in zip(words, ref_words):
        assert a["tokens"] == b["tokens"] and a["word"] == b["word"]
        worst_t = max(worst_t, abs(a["start"] - b["start"]), abs(a["end"] - b["end"]))
        worst_p = max(worst_p, abs(a["probability"] - b["probability"]) / b["probability"])
    print(f"[{name}] transcribe window, {steps} steps: {len(words)} words, worst |dt| {worst_t * 1e3:.0f} ms, prob rel {worst_p:.1e}, "
          f"avg_logprob {info['decode'][0].avg_logprob:.5f} vs {ex['decode'].avg_logprob:.5f}")
    assert worst_t <= 0.0201 and worst_p <= 2e-3
    assert abs(info["decode"][0].avg_logprob - ex["decode"].avg_logprob) < 1e-3


def test_large_v3_free_running_greedy_tokens_bit_exact():
    """No forcing: 32 free-running greedy steps at 32 layers (timestamp rules + silent-timestamp mask active)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200.decode import DecodingOptions, decode_windows
    W, SP, model, otk, gm, tk = _build("large-v3", seed=4)
    audio = SP.synth_audio(480000, seed=31)
    mask = torch.zeros(1501, dtype=torch.bool)
    mask[700:1501] = True
    mel_ref = W.pad_or_trim(W.log_mel_spectrogram(audio, model.dims.n_mels), 3000)
    ref, _, _ = SP.decode_window(model, mel_ref, ts_token_mask=mask, language="en", sample_len=32)
    res, _ = decode_windows(gm, tk, gm.encode(gm.log_mel(audio.cuda()[None])), DecodingOptions(language="en", sample_len=32),
                            ts_token_mask=mask)
    print(f"[large-v3] free-running: {len(ref.tokens)} tokens, avg_logprob {res[0].avg_logprob:.5f} vs {ref.avg_logprob:.5f}")
    assert res[0].tokens == ref.tokens
    assert abs(res[0].avg_logprob - ref.avg_logprob) < 1e-3
