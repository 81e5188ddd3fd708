"""GPU tests of the sampling side of the decode path (SURVEY.md section 8 row a9 beyond temperature 0):

* ``stb_sample`` at temperature > 0 on given logits == inverse-CDF draw restated in float64 numpy (the draw is a pure function of
  the caller's uniforms; whisper's ``Categorical(logits / T).sample()`` has no cross-device reproducible stream);
* ragged initial tokens (per-window prompts) in ONE batch == the oracle decoding each window alone with its prompt
  (whisper DecodingTask._get_initial_tokens; stable_whisper original_whisper.py:533);
* per-sequence stop at ``tokens.shape[-1] > n_ctx`` (decode.py:60) for long prompts.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_temperature_draw_is_inverse_cdf_of_the_uniforms():
    _gpu()
    from stable_ts_b200 import _lib as L
    lib = L.lib()
    B, V, rows, T = 6, 51866, 4, 0.7
    ld = (V + 7) // 8 * 8
    g = torch.Generator().manual_seed(7)
    base = torch.randn(B, ld, generator=g) * 1.5
    base[1, 100:40000] = float("-inf")                   # masked ranges are skipped by the running sum
    base[2, 5] = 30.0                                    # one dominant token
    base[3, 0] = base[3, V - 1] = 2.0
    u = torch.rand(rows, B, generator=g)
    u[0, 3], u[1, 3] = 0.0, 1.0 - 2.0 ** -24             # first / last token with probability > 0
    logits = base.cuda()
    states = torch.zeros(B, 6, dtype=torch.int32, device="cuda")
    states[:, 3] = -1
    nxt = torch.zeros(B, dtype=torch.int32, device="cuda")
    tok = torch.zeros(rows, B, dtype=torch.int32, device="cuda")
    arg = torch.zeros(rows, B, dtype=torch.int32, device="cuda")
    ud = u.cuda()
    cap = torch.tensor([rows, rows, rows, rows, 2, rows], dtype=torch.int32, device="cuda")
    eot = 50256
    for _ in range(rows):
        L.check(lib.stb_sample(L.ptr(logits), ld, B, V, eot, 50363, 50362, None, None, None, 0, -1, 0, None, L.ptr(states),
                               L.ptr(nxt), L.ptr(tok), L.ptr(arg), rows, T, L.ptr(ud), L.ptr(cap), L.stream_ptr()))
    torch.cuda.synchronize()
    tok, arg, st = tok.cpu().numpy(), arg.cpu().numpy(), states.cpu()
    lg = np.nan_to_num(base[:, :V].double().numpy(), neginf=-np.finfo(np.float32).max)
    sum_lp = np.zeros(B)
    for b in range(B):
        x = lg[b]
        p = np.exp((x - x.max()) / T)
        c = np.cumsum(p)
        lsm = x - x.max() - np.log(np.exp(x - x.max()).sum())
        done = False
        for r in range(rows):
            want = int(np.argmax(c > float(u[r, b]) * c[-1]))
            if done or (b == 4 and r >= 2):
                want = eot
            else:
                sum_lp[b] += lsm[want]
            assert tok[r, b] == want, (r, b, tok[r, b], want)
            assert arg[r, b] == int(np.argmax(x))
            done = done or want == eot
    got = st[:, 5].contiguous().view(torch.float32).numpy()
    np.testing.assert_allclose(got, sum_lp, rtol=2e-5, atol=1e-4)
    assert tok[0, 3] == 0 and tok[1, 3] == V - 1


def _setup(name="tiny.en", seed=3):
    import oracle.whisper_ref as W
    from stable_ts_b200.model import from_oracle
    from stable_ts_b200.tokenizer import get_tokenizer
    om = W.build_model(name, seed=seed)
    gm = from_oracle(om)
    tk = get_tokenizer(gm, language="en", task="transcribe", synthetic=True)
    return W, om, gm, tk


def test_ragged_prompts_in_one_batch_match_the_oracle_window_by_window():
    _gpu()
    from oracle import stable_path as SP
    from stable_ts_b200.decode import DecodingOptions, decode_windows
    W, om, gm, tk = _setup()
    audios = [SP.synth_audio(480000, seed=31 + i) for i in range(4)]
    g = torch.Generator().manual_seed(5)
    prompts = [[], torch.randint(300, 40000, (5,), generator=g).tolist(), torch.randint(300, 40000, (37,), generator=g).tolist(),
               torch.randint(300, 40000, (260,), generator=g).tolist()]          # the last one is cut to n_ctx // 2 - 1 = 223
    enc = gm.encode(gm.log_mel(torch.stack(audios).cuda()))
    res, ex = decode_windows(gm, tk, enc, DecodingOptions(language="en", sample_len=20), prompts=prompts)
    for b, (a, p) in enumerate(zip(audios, prompts)):
        mel = W.pad_or_trim(W.log_mel_spectrogram(a, om.dims.n_mels), 3000)
        ref, _, _ = SP.decode_window(om, mel, language="en", sample_len=20, prompt=p or None)
        assert res[b].tokens == ref.tokens, (b, res[b].tokens, ref.tokens)
        assert abs(res[b].avg_logprob - ref.avg_logprob) < 1e-3
        assert abs(res[b].no_speech_prob - ref.no_speech_prob) <= 2e-3 * ref.no_speech_prob + 1e-9
    print(f"ragged prompts {[len(p) for p in prompts]}: tokens bit-exact vs the oracle, window by window")


def test_long_prompt_stops_at_n_ctx_like_the_reference():
    """1 + 223 + 1 = 225 initial tokens (tiny.en: sot_sequence is one token): the loop stops once 449 > n_ctx tokens exist, i.e.
    after 224 sampled tokens, although sample_len is 230; the window without a prompt in the same batch samples all 230."""
    _gpu()
    from oracle import stable_path as SP
    from stable_ts_b200.decode import DecodingOptions, decode_windows
    W, om, gm, tk = _setup()
    audios = [SP.synth_audio(480000, seed=41 + i) for i in range(2)]
    g = torch.Generator().manual_seed(6)
    prompts = [torch.randint(300, 40000, (223,), generator=g).tolist(), []]
    forced = torch.randint(300, 40000, (230, 2), generator=g, dtype=torch.int32)
    enc = gm.encode(gm.log_mel(torch.stack(audios).cuda()))
    res, ex = decode_windows(gm, tk, enc, DecodingOptions(language="en", sample_len=230), prompts=prompts, forced_tokens=forced)
    for b, (a, p) in enumerate(zip(audios, prompts)):
        mel = W.pad_or_trim(W.log_mel_spectrogram(a, om.dims.n_mels), 3000)
        ref, _, rex = SP.decode_window(om, mel, language="en", sample_len=230, prompt=p or None, forced_tokens=forced[:, b].tolist())
        n = len(rex["step_argmax"])
        assert n == (224 if p else 230)
        assert ex["step_argmax"][:n, b].tolist() == rex["step_argmax"]
        assert res[b].tokens == ref.tokens
        assert abs(res[b].avg_logprob - ref.avg_logprob) < 1e-3
