"""GPU parity tests, path level: the B200 alignment closure / batched drivers against (a) fixtures the UNMODIFIED
reference produced (tests/golden, oracle/make_golden.py) and (b) the live CPU oracle.  North-star gates: word start/end
within +-20 ms, logits within 1e-3 relative; DTW bit-exact on identical inputs (tests/test_gpu_kernels.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _WT:
    def __init__(self, word, tokens):
        self.word, self.tokens = word, tokens


def _golden_case(name):
    import oracle.whisper_ref as W
    from oracle import stable_path as SP
    from oracle.whisper_ref.model import ModelDimensions
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    z = np.load(os.path.join(GOLD, f"{name}.npz"))
    dims = ModelDimensions(*[int(v) for v in z["dims"]])
    model = W.build_model(dims, seed=int(z["model_seed"]))
    gm = from_oracle(model)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    script = z["script"].tolist()
    wts, i = [], 0
    for n in z["word_lens"]:
        wts.append(script[i:i + int(n)])
        i += int(n)
    audio = SP.synth_audio(int(z["n_samples"]), seed=1234)
    return z, model, gm, tk, script, wts, audio


@pytest.mark.parametrize("name", ["mini_en", "mini_ml"])
def test_alignment_closure_matches_reference_fixture(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200.alignment import align_words_batch, get_b200_alignment_func
    z, model, gm, tk, script, wts, audio = _golden_case(name)
    f = get_b200_alignment_func(gm, tk)
    out = f(audio, [_WT(tk.decode(w), w) for w in wts])
    assert len(out) == len(z["word_start"])
    ds = np.abs(np.array([w["start"] for w in out]) - z["word_start"]).max()
    de = np.abs(np.array([w["end"] for w in out]) - z["word_end"]).max()
    dp = np.abs(np.array([w["probability"] for w in out]) / z["word_prob"] - 1).max()
    _, inter = align_words_batch(gm, tk, [audio], [wts], return_intermediates=True)
    dm = np.abs(inter["matrices"][0].numpy() - z["matrix"]).max()
    print(f"[{name}] |dstart| {ds:.3f}s |dend| {de:.3f}s prob rel {dp:.2e} matrix abs {dm:.2e}")
    assert ds <= 0.0201 and de <= 0.0201 and dp < 2e-3 and dm < 1e-3


@pytest.mark.parametrize("name,n_win", [("tiny.en", 5), ("base", 3)])
def test_batched_windows_match_live_oracle(name, n_win):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle.whisper_ref as W
    from oracle import stable_path as SP
    from stable_ts_b200.alignment import align_words_batch
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    model = W.build_model(name, seed=3)
    gm = from_oracle(model)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    otk = W.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en", task="transcribe")
    lens = [480000, 300000, 123456, 480000, 64000][:n_win]
    ntok = [60, 25, 12, 60, 5][:n_win]
    audios = [SP.synth_audio(n, seed=10 + i) for i, n in enumerate(lens)]
    wts = [SP.words_from_script(SP.synth_token_script(k, tk.eot, seed=20 + i), seed=i) for i, k in enumerate(ntok)]
    got = align_words_batch(gm, tk, audios, wts)
    worst, total, bad = 0.0, 0, 0
    for a, wt, g in zip(audios, wts, got):
        ref = SP.align_audio_window(model, otk, wt, a)
        assert len(ref) == len(g)
        for r, w in zip(ref, g):
            d = max(abs(r["start"] - w["start"]), abs(r["end"] - w["end"]))
            worst = max(worst, d)
            total += 1
            bad += d > 0.0201
            assert abs(r["probability"] - w["probability"]) <= 2e-3 * r["probability"] + 1e-12
    print(f"[{name}] {total} words, worst |dt| {worst:.3f}s, outside +-20ms: {bad}")
    assert bad == 0


def test_refine_probs_and_rank_match_oracle():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle.whisper_ref as W
    from oracle import stable_path as SP
    from stable_ts_b200.alignment import get_b200_refinement_func, refine_probs
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    model = W.build_model("tiny", seed=2)
    gm = from_oracle(model)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    otk = W.tokenizer.get_tokenizer(True, num_languages=model.num_languages, language="en", task="transcribe")
    script = SP.synth_token_script(30, tk.eot)
    a2 = torch.stack([SP.synth_audio(200000, seed=1), SP.synth_audio(200000, seed=2)])
    a2[1, 50000:90000] = 0                                     # the Refiner mutes spans of one row
    probs3_ref = SP.refine_token_probs(model, otk, a2, script)
    p_ref, r_ref = SP.prob_and_rank(probs3_ref, script)
    p, r = refine_probs(gm, tk, a2, script)
    np.testing.assert_allclose(p.cpu().numpy(), p_ref.numpy(), rtol=2e-3)
    # 3-D form (what the unmodified Refiner consumes, refinement.py:305-325): probabilities of every class on the device
    probs3 = get_b200_refinement_func(gm, tk)(a2, script)
    assert probs3.shape == (2, len(script), tk.eot) and probs3.is_cuda
    np.testing.assert_allclose(probs3.cpu().numpy(), probs3_ref.numpy(), rtol=2e-3, atol=1e-12)
    # (1) integer self-consistency, EXACT: the counting kernel's rank == the Refiner's own derivation (ascending sort
    #     position of the target) applied to the probabilities this path returns -- ties aside (none at these magnitudes)
    tgt = torch.tensor(script, device=probs3.device)
    idx = torch.arange(len(script), device=probs3.device)
    own = torch.stack([(probs3[i, idx].sort().indices == tgt.unsqueeze(1)).nonzero()[:, -1] for i in range(2)])
    tgt_p = probs3[:, idx, tgt]
    tie = torch.stack([(probs3[i] == tgt_p[i][:, None]).sum(-1) > 1 for i in range(2)])
    assert torch.equal(own[~tie], r.long()[~tie]), "counting rank differs from the sort position of the same probabilities"
    # (2) against the fp32 CPU oracle: the rank is an integer function of floats that agree to ~1e-5, so it may differ only
    #     where the oracle itself has classes within that tolerance of the target -- bound it by the oracle's own counts
    eps = 2e-3
    lo = (probs3_ref < (p_ref * (1 - eps))[:, :, None]).sum(-1)
    hi = (probs3_ref <= (p_ref * (1 + eps))[:, :, None]).sum(-1) - 1
    rc = r.cpu().long()
    dr = (rc - r_ref).abs().max().item()
    print(f"refine: prob rel {((p.cpu() - p_ref).abs() / p_ref).max():.2e}, ranks equal {int((rc == r_ref).sum())}/{rc.numel()}, "
          f"max delta {dr} (all inside the oracle's near-tie interval)")
    assert bool(((rc >= lo) & (rc <= hi)).all())
    p2 = get_b200_refinement_func(gm, tk, form="2d")(a2, script)
    assert p2.shape == (2, len(script)) and torch.allclose(p2, p.cpu())


@pytest.mark.parametrize("dyn,aligner", [(True, "legacy"), ("4,2", "legacy"), (None, "new")])
def test_dynamic_heads_and_new_aligner_match_live_oracle(dyn, aligner):
    """a5 variants end to end (stable_whisper/timing.py:85-103,115-163) through the alignment closure."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle.whisper_ref as W
    from oracle import stable_path as SP
    from stable_ts_b200.alignment import align_words_batch
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    model = W.build_model("tiny", seed=9)
    gm = from_oracle(model)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    otk = W.tokenizer.get_tokenizer(True, num_languages=model.num_languages, language="en", task="transcribe")
    audios = [SP.synth_audio(480000, seed=71), SP.synth_audio(200000, seed=72)]
    wts = [SP.words_from_script(SP.synth_token_script(40, tk.eot, seed=73)), SP.words_from_script(SP.synth_token_script(15, tk.eot, seed=74))]
    got = align_words_batch(gm, tk, audios, wts, dynamic_heads=dyn, aligner=aligner)
    worst, bad, total = 0.0, 0, 0
    for a, wt, g in zip(audios, wts, got):
        ref = SP.align_audio_window(model, otk, wt, a, dynamic_heads=dyn, aligner=aligner)
        assert len(ref) == len(g)
        for r, w in zip(ref, g):
            d = max(abs(r["start"] - w["start"]), abs(r["end"] - w["end"]))
            worst, total, bad = max(worst, d), total + 1, bad + (d > 0.0201)
    print(f"dynamic_heads={dyn} aligner={aligner}: {total} words, worst |dt| {worst:.3f}s, outside +-20ms: {bad}")
    assert bad == 0
