"""GPU parity tests, kernel level: every C-ABI entry point against the CPU oracle on seeded inputs.

Run on the B200 box:  python -m pytest tests -m gpu -q
The oracle (oracle/) is the checker only; the thing under test is libstablets_b200.so called through ctypes.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stable_ts_b200 import _lib
    _lib.lib()
    return _lib


def _split(x, lo=True):
    hi = x.half()
    return hi, ((x - hi.float()).half() if lo else None)


def _gemm(L, A, B, passes=3, bias=None, act=0, residual=None, alpha=1.0, out="f32", transposed=False):
    """A [M,K], B [N,K] fp32 cuda -> D via stb_gemm."""
    M, K = A.shape
    N = B.shape[0]
    ah, al = _split(A, passes == 3)
    bh, bl = _split(B, passes == 3)
    opa = L.Operand(L.ptr(ah), L.ptr(al), M, K, K, 0, 0)
    opb = L.Operand(L.ptr(bh), L.ptr(bl), N, K, K, 0, 0)
    shape = (N, M) if transposed else (M, N)
    ld = shape[1]
    ldp = (ld + 7) // 8 * 8
    of = torch.full((shape[0], ldp), float("nan"), device="cuda") if out == "f32" else None
    oh = torch.zeros(shape[0], ldp, dtype=torch.float16, device="cuda") if out == "split" else None
    ol = torch.zeros(shape[0], ldp, dtype=torch.float16, device="cuda") if out == "split" else None
    ep = L.Epilogue()
    ep.out_f32, ep.out_hi, ep.out_lo = L.ptr(of), L.ptr(oh), L.ptr(ol)
    ep.ld_out = ldp
    ep.transposed = int(transposed)
    ep.bias = L.ptr(bias)
    ep.residual = L.ptr(residual)
    ep.ld_res = residual.stride(0) if residual is not None else 0
    ep.alpha = alpha
    ep.act = act
    L.check(L.lib().stb_gemm(ctypes.byref(opa), ctypes.byref(opb), 1, 1, ctypes.byref(ep), L.stream_ptr()))
    torch.cuda.synchronize()
    if out == "f32":
        return of[:, :ld]
    return oh[:, :ld].float() + ol[:, :ld].float()


def _ref(A, B, bias=None, act=0, residual=None, alpha=1.0):
    D = alpha * (A.double().cpu() @ B.double().cpu().T)
    if bias is not None:
        D = D + bias.double().cpu()
    if act == 1:
        D = torch.nn.functional.gelu(D)
    if residual is not None:
        D = D + residual.double().cpu()
    return D


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (300, 200, 192), (1500, 384, 384), (77, 1000, 128),
                                   (256, 64, 1504), (130, 16, 64), (200, 40, 320), (64, 51866, 128)])
@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_plain(L, M, N, K, passes):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.3
    D = _gemm(L, A, B, passes).double().cpu()
    R = _ref(A, B)
    err = (D - R).abs().max().item() / R.abs().max().item()
    print(f"gemm {M}x{N}x{K} passes={passes}: rel err {err:.3e}")
    assert torch.isfinite(D).all()
    assert err < (max(3e-6, 2e-7 * K ** 0.5 * 2) if passes == 3 else 3e-3)   # grows ~sqrt(K)


def _gemv(L, X, W, passes=3, bias=None, act=0, residual=None, out="f32"):
    """X [B,K], W [N,K] fp32 cuda -> act(X W^T + bias) + residual via stb_gemv (decode-step linear)."""
    Bn, K = X.shape
    N = W.shape[0]
    xh, xl = _split(X, passes == 3)
    wh, wl = _split(W, passes == 3)
    ld = (N + 7) // 8 * 8
    of = torch.full((Bn, ld), float("nan"), device="cuda") if out == "f32" else None
    oh = torch.zeros(Bn, ld, dtype=torch.float16, device="cuda") if out == "split" else None
    ol = torch.zeros(Bn, ld, dtype=torch.float16, device="cuda") if out == "split" else None
    L.check(L.lib().stb_gemv(L.ptr(xh), L.ptr(xl), Bn, K, L.ptr(wh), L.ptr(wl), N, L.ptr(bias), act, L.ptr(residual),
                             residual.stride(0) if residual is not None else 0, L.ptr(of), L.ptr(oh), L.ptr(ol), ld,
                             L.stream_ptr()))
    torch.cuda.synchronize()
    return of[:, :N] if out == "f32" else oh[:, :N].float() + ol[:, :N].float()


@pytest.mark.parametrize("Bn,N,K", [(64, 1280, 1280), (50, 1280, 5120), (33, 5120, 1280), (16, 3840, 1280), (1, 384, 384),
                                    (7, 1002, 384), (18, 384, 1536), (3, 130, 128), (64, 51866, 128), (40, 512, 2048)])
@pytest.mark.parametrize("passes", [3, 1])
def test_gemv_decode_linear_vs_fp64(L, Bn, N, K, passes):
    """every width of the Whisper family (K = 128 .. 5120: one or several K chunks, odd/even slices per CTA, idle cluster
    ranks), ragged sequence groups and a ragged last feature tile"""
    g = torch.Generator(device="cuda").manual_seed(Bn * 31 + N + K)
    X = torch.randn(Bn, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.3
    D = _gemv(L, X, W, passes).double().cpu()
    R = _ref(X, W)
    err = (D - R).abs().max().item() / R.abs().max().item()
    print(f"gemv B={Bn} N={N} K={K} passes={passes}: rel err {err:.3e}")
    assert torch.isfinite(D).all()
    assert err < (max(3e-6, 2e-7 * K ** 0.5 * 2) if passes == 3 else 3e-3)


def test_gemv_epilogues_and_in_place_residual(L):
    g = torch.Generator(device="cuda").manual_seed(11)
    Bn, N, K = 37, 640, 1280
    X = torch.randn(Bn, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.2
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(Bn, N, device="cuda", generator=g)
    for kw in [dict(bias=bias), dict(bias=bias, act=1), dict(bias=bias, residual=res), dict(bias=bias, act=1, residual=res)]:
        D = _gemv(L, X, W, 3, **kw).double().cpu()
        R = _ref(X, W, **kw)
        assert (D - R).abs().max().item() / R.abs().max().item() < 5e-6
    D = _gemv(L, X, W, 3, bias=bias, act=1, out="split").double().cpu()
    assert (D - _ref(X, W, bias=bias, act=1)).abs().max().item() / _ref(X, W, bias=bias, act=1).abs().max().item() < 5e-6
    # the decoder's residual stream is updated in place: out_f32 == res
    xh, xl = _split(X)
    wh, wl = _split(W)
    stream = res.clone()
    L.check(L.lib().stb_gemv(L.ptr(xh), L.ptr(xl), Bn, K, L.ptr(wh), L.ptr(wl), N, L.ptr(bias), 0, L.ptr(stream), N,
                             L.ptr(stream), None, None, N, L.stream_ptr()))
    torch.cuda.synchronize()
    R = _ref(X, W, bias=bias, residual=res)
    assert (stream.double().cpu() - R).abs().max().item() / R.abs().max().item() < 5e-6


def test_gemm_epilogues(L):
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 333, 264, 448
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.2
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    for kw in [dict(bias=bias), dict(bias=bias, act=1), dict(bias=bias, residual=res), dict(alpha=0.125),
               dict(bias=bias, act=1, residual=res)]:
        D = _gemm(L, A, B, 3, **kw).double().cpu()
        R = _ref(A, B, **kw)
        err = (D - R).abs().max().item() / R.abs().max().item()
        print("epilogue", {k: (v if not torch.is_tensor(v) else "T") for k, v in kw.items()}, f"{err:.3e}")
        assert err < 5e-6
    D = _gemm(L, A, B, 3, bias=bias, act=1, out="split").double().cpu()
    R = _ref(A, B, bias=bias, act=1)
    assert (D - R).abs().max().item() / R.abs().max().item() < 5e-6
    D = _gemm(L, A, B, 3, bias=bias, out="split", transposed=True).double().cpu()
    assert (D - _ref(A, B, bias=bias).T).abs().max().item() < 5e-5
    D = _gemm(L, A, B, 3, out="f32", transposed=True).double().cpu()
    assert (D - _ref(A, B).T).abs().max().item() < 5e-5


def test_gemm_batched_head_views(L):
    """q k^T over a fused [B*T][2d] buffer: 4-D operand views with head and batch strides (attention scores)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    Bn, H, T, d = 2, 3, 150, 192
    qk = torch.randn(Bn * T, 2 * d, device="cuda", generator=g)
    hi, lo = _split(qk)
    opa = L.Operand(L.ptr(hi), L.ptr(lo), T, 64, 2 * d, 64, T * 2 * d)
    opb = L.Operand(hi.data_ptr() + d * 2, lo.data_ptr() + d * 2, T, 64, 2 * d, 64, T * 2 * d)
    ldk = (T + 7) // 8 * 8
    S = torch.full((Bn, H, T, ldk), float("nan"), device="cuda")
    ep = L.Epilogue()
    ep.out_f32 = L.ptr(S)
    ep.ld_out, ep.out_h_stride, ep.out_b_stride = ldk, T * ldk, H * T * ldk
    ep.alpha = 0.125
    L.check(L.lib().stb_gemm(ctypes.byref(opa), ctypes.byref(opb), Bn, H, ctypes.byref(ep), L.stream_ptr()))
    torch.cuda.synchronize()
    x = qk.double().cpu().view(Bn, T, 2, H, 64)
    R = 0.125 * torch.einsum("bthc,bshc->bhts", x[:, :, 0], x[:, :, 1])
    err = (S[..., :T].double().cpu() - R).abs().max().item() / R.abs().max().item()
    print(f"batched head-view gemm rel err {err:.3e}")
    assert err < 3e-6


@pytest.mark.parametrize("stride_mult,C", [(1, 80), (2, 128)])
def test_gemm_overlapping_rows_is_conv(L, stride_mult, C):
    """conv1d(k=3, pad=1, stride s) as a GEMM over an overlapping-row TMA view of the time-major padded input."""
    g = torch.Generator(device="cuda").manual_seed(11)
    Tin, Cout = 600, 96
    x = torch.randn(1, C, Tin, device="cuda", generator=g)
    w = torch.randn(Cout, C, 3, device="cuda", generator=g) * 0.1
    Tout = (Tin + 2 - 3) // stride_mult + 1
    xt = torch.zeros(Tin + 2, C, device="cuda")
    xt[1:Tin + 1] = x[0].T
    hi, lo = _split(xt)
    wh, wl = _split(w.permute(0, 2, 1).reshape(Cout, 3 * C).contiguous())
    opa = L.Operand(L.ptr(hi), L.ptr(lo), Tout, 3 * C, stride_mult * C, 0, 0)
    opb = L.Operand(L.ptr(wh), L.ptr(wl), Cout, 3 * C, 3 * C, 0, 0)
    out = torch.full((Tout, Cout), float("nan"), device="cuda")
    ep = L.Epilogue()
    ep.out_f32, ep.ld_out, ep.alpha = L.ptr(out), Cout, 1.0
    L.check(L.lib().stb_gemm(ctypes.byref(opa), ctypes.byref(opb), 1, 1, ctypes.byref(ep), L.stream_ptr()))
    torch.cuda.synchronize()
    R = torch.nn.functional.conv1d(x.double().cpu(), w.double().cpu(), padding=1, stride=stride_mult)[0].T
    err = (out.double().cpu() - R).abs().max().item() / R.abs().max().item()
    print(f"conv-as-gemm stride {stride_mult} C {C}: rel err {err:.3e}")
    assert err < 3e-6


@pytest.mark.parametrize("Bn,H,Mq,Mk,passes", [(1, 2, 128, 128, 3), (2, 3, 300, 300, 3), (1, 2, 1500, 1500, 3),
                                                (2, 2, 200, 1500, 3), (1, 2, 1500, 1500, 1)])
def test_fused_attention_vs_fp64(L, Bn, H, Mq, Mk, passes):
    """tcgen05 fused attention (scores in TMEM, exact two-pass softmax) vs softmax(q k^T / 8) v in fp64."""
    g = torch.Generator(device="cuda").manual_seed(Mq + Mk + H)
    d = 64 * H
    q = torch.randn(Bn, Mq, d, device="cuda", generator=g) * 1.5
    k = torch.randn(Bn, Mk, d, device="cuda", generator=g) * 1.5
    v = torch.randn(Bn, Mk, d, device="cuda", generator=g)
    ldk = (Mk + 7) // 8 * 8
    vT = torch.zeros(Bn, H, 64, ldk, device="cuda")
    vT[..., :Mk] = v.view(Bn, Mk, H, 64).permute(0, 2, 3, 1)
    lo = passes == 3
    qh, ql = _split(q, lo)
    kh, kl = _split(k, lo)
    vh, vl = _split(vT, lo)
    oq = L.Operand(L.ptr(qh), L.ptr(ql), Mq, 64, d, 64, Mq * d)
    ok = L.Operand(L.ptr(kh), L.ptr(kl), Mk, 64, d, 64, Mk * d)
    ov = L.Operand(L.ptr(vh), L.ptr(vl), 64, ldk, ldk, 64 * ldk, d * ldk)
    out_hi = torch.zeros(Bn, Mq, d, dtype=torch.float16, device="cuda")
    out_lo = torch.zeros_like(out_hi)
    L.check(L.lib().stb_attention(ctypes.byref(oq), ctypes.byref(ok), ctypes.byref(ov), Bn, H, Mq, Mk, L.ptr(out_hi),
                                  L.ptr(out_lo) if lo else None, d, 64, Mq * d, L.stream_ptr()))
    torch.cuda.synchronize()
    out = (out_hi.float() + out_lo.float()).double().cpu()
    qd, kd, vd = [x.double().cpu().view(Bn, -1, H, 64).permute(0, 2, 1, 3) for x in (q, k, v)]
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) / 8.0, dim=-1) @ vd).permute(0, 2, 1, 3).reshape(Bn, Mq, d)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"fused attention B{Bn} H{H} {Mq}x{Mk} passes={passes}: rel err {err:.3e}")
    assert torch.isfinite(out).all()
    assert err < (1e-5 if passes == 3 else 5e-3)


# --------------------------------------------------------------------------------------------------------- DTW
def _dtw_gpu(L, x, negate=False, want_path=True):
    from stable_ts_b200 import _lib
    B, R, F = x.shape
    jumps = torch.empty(B, R, dtype=torch.int32, device="cuda")
    path = torch.empty(B, 2, R + F, dtype=torch.int32, device="cuda")
    plen = torch.empty(B, dtype=torch.int32, device="cuda")
    ws = torch.empty(L.lib().stb_dtw_ws_bytes(B, R, F), dtype=torch.uint8, device="cuda")
    L.check(L.lib().stb_dtw(L.ptr(x), B, R, F, x.stride(1), int(negate), L.ptr(jumps), L.ptr(path), L.ptr(plen), L.ptr(ws),
                            ws.numel(), L.stream_ptr()))
    torch.cuda.synchronize()
    return jumps.cpu().numpy(), path.cpu().numpy(), plen.cpu().numpy()


@pytest.mark.parametrize("R,F,quant", [(1, 7, False), (9, 1, False), (7, 31, False), (33, 64, False), (41, 333, True),
                                       (101, 937, False), (106, 1500, False), (230, 1500, False), (449, 1500, False),
                                       (449, 1500, True), (64, 33, False)])
def test_dtw_bit_exact_vs_oracle(L, R, F, quant):
    from oracle import c_oracle
    rng = np.random.default_rng(R * 10007 + F)
    Bn = 3
    x = rng.standard_normal((Bn, R, F)).astype(np.float32)
    if quant:
        x = np.round(x * 2) / 2            # many exact ties: exercises the strict-'<' rule
    jumps, path, plen = _dtw_gpu(L, torch.from_numpy(x).cuda())
    for b in range(Bn):
        p_ref, j_ref = c_oracle.dtw(x[b])
        n = int(plen[b])
        assert n == p_ref.shape[1], (n, p_ref.shape)
        assert np.array_equal(path[b, :, :n], p_ref)
        assert np.array_equal(jumps[b], j_ref)


def test_dtw_goldens_from_reference(L):
    z = np.load(os.path.join(GOLD, "dtw_cases.npz"))
    for i in range(3):
        x, p = z[f"x{i}"], z[f"p{i}"]
        jumps, path, plen = _dtw_gpu(L, torch.from_numpy(x)[None].cuda(), negate=True)
        assert np.array_equal(path[0, :, :int(plen[0])], p)


def test_dtw_nan_and_strided_rows(L):
    from oracle import c_oracle
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 20, 104)).astype(np.float32)
    x[0, 5] = np.nan
    xt = torch.from_numpy(x).cuda()
    jumps, path, plen = _dtw_gpu(L, xt[:, :, :101])            # row pitch 104, F = 101
    for b in range(2):
        p_ref, j_ref = c_oracle.dtw(x[b, :, :101])
        assert np.array_equal(path[b, :, :int(plen[b])], p_ref)
        assert np.array_equal(jumps[b], j_ref)


# --------------------------------------------------------------------------------------------------------- QK post
@pytest.mark.parametrize("A,M,F,S", [(6, 44, 937, 1), (10, 109, 1500, 3), (3, 12, 50, 3), (2, 6, 3, 1)])
def test_qk_postprocess_vs_oracle(L, A, M, F, S):
    from oracle import stable_path as SP
    g = torch.Generator().manual_seed(A * 100 + M)
    Bn = 2
    qk = torch.randn(Bn, A, M, 1504, generator=g) * 2.0
    lib = L.lib()
    R = M - 1 - S
    ldm = (F + 3) // 4 * 4
    out = torch.empty(Bn, R, ldm, device="cuda")
    ws = torch.empty(lib.stb_qkpost_ws_bytes(Bn, A, R, F), dtype=torch.uint8, device="cuda")
    qkc = qk.cuda()
    L.check(lib.stb_qk_postprocess(L.ptr(qkc), Bn, A, M, 1504, S, R, F, 1.0, 7, L.ptr(out), ldm, L.ptr(ws), ws.numel(),
                                   L.stream_ptr()))
    torch.cuda.synchronize()
    for b in range(Bn):
        qks = [qk[b:b + 1, a:a + 1] for a in range(A)]           # one "layer" per head, head index 0
        w = SP.attention_weights_legacy(qks, [(a, 0) for a in range(A)], S, F * 320)
        ref = w.mean(dim=0)
        got = out[b, :, :F].cpu()
        err = (got - ref).abs().max().item()
        print(f"qkpost A={A} M={M} F={F}: max abs err {err:.3e}")
        assert err < 2e-5


# --------------------------------------------------------------------------------------------------------- log-mel
@pytest.mark.parametrize("n,n_mels", [(480000, 80), (300000, 128), (123457, 80)])
def test_logmel_align_mode_vs_oracle(L, n, n_mels):
    from oracle import stable_path as SP
    from oracle.whisper_ref import audio as A
    from stable_ts_b200.model import mel_filterbank
    assert np.array_equal(mel_filterbank(n_mels), A.mel_filterbank_np(n_mels))
    x = SP.synth_audio(n, seed=n % 1000)
    ref = A.pad_or_trim(A.log_mel_spectrogram(x, n_mels, padding=480000 - n), 3000)
    got = _logmel(L, x[None].cuda(), 480000, n_mels, False)[0].cpu()
    err = (got - ref).abs().max().item()
    print(f"logmel n={n} mels={n_mels}: max abs err {err:.3e}")
    assert err < 2e-4


def test_logmel_refine_mode_batch_global_max(L):
    from oracle import stable_path as SP
    from oracle.whisper_ref import audio as A
    n = 200000
    x = torch.stack([SP.synth_audio(n, seed=1), 0.05 * SP.synth_audio(n, seed=2)])
    ref = A.pad_or_trim(A.log_mel_spectrogram(x, 80), 3000)
    got = _logmel(L, x.cuda(), n, 80, True).cpu()
    err = (got - ref).abs().max().item()
    print(f"logmel refine-mode: max abs err {err:.3e}")
    assert err < 2e-4
    assert (got[:, :, n // 160:] == 0).all()


def _logmel(L, audio, padded, n_mels, global_max):
    from stable_ts_b200.model import mel_filterbank
    B, n = audio.shape
    k = np.arange(400, dtype=np.float64)
    window = torch.from_numpy((0.5 - 0.5 * np.cos(2 * np.pi * k / 400)).astype(np.float32)).cuda()
    dft = torch.from_numpy(np.stack([np.cos(2 * np.pi * k / 400), np.sin(2 * np.pi * k / 400)], 1).astype(np.float32)).cuda()
    filt = torch.from_numpy(mel_filterbank(n_mels)).cuda()
    mel = torch.empty(B, n_mels, 3000, device="cuda")
    ws = torch.empty(B * 376 * 4, dtype=torch.uint8, device="cuda")
    L.check(L.lib().stb_logmel(L.ptr(audio), B, n, padded, n_mels, L.ptr(filt), L.ptr(window), L.ptr(dft), int(global_max),
                               L.ptr(mel), L.ptr(ws), ws.numel(), L.stream_ptr()))
    torch.cuda.synchronize()
    return mel


# --------------------------------------------------------------------------------------------------------- token probs
def test_token_probs_and_rank(L):
    g = torch.Generator().manual_seed(2)
    n, V, eot = 37, 51872, 50257
    logits = torch.randn(n, V, generator=g) * 3
    tgt = torch.randint(0, eot, (n,), generator=g)
    lc = logits.cuda()
    prob = torch.empty(n, device="cuda")
    rank = torch.empty(n, dtype=torch.int32, device="cuda")
    tc = tgt.to(torch.int32).cuda()
    L.check(L.lib().stb_token_probs(L.ptr(lc), V, n, eot, L.ptr(tc), L.ptr(prob), L.ptr(rank), L.stream_ptr()))
    torch.cuda.synchronize()
    p = logits[:, :eot].softmax(-1)
    ref = p[torch.arange(n), tgt]
    np.testing.assert_allclose(prob.cpu().numpy(), ref.numpy(), rtol=2e-5)
    order = p.sort(dim=-1).indices
    ref_rank = (order == tgt[:, None]).nonzero()[:, -1]
    assert np.array_equal(rank.cpu().numpy(), ref_rank.numpy())


# --------------------------------------------------------------------------------------------------------- a5 variants
def _model_stub(L):
    """A B200Whisper is not needed for the post-processing kernels: call the C ABI directly."""
    return L.lib()


@pytest.mark.parametrize("LH,M,F,S,count,iters", [(8, 30, 500, 1, 6, 1), (12, 44, 937, 3, 4, 2), (24, 20, 1500, 3, 6, 3)])
def test_qk_postprocess_dynamic_heads_vs_oracle(L, LH, M, F, S, count, iters):
    from oracle import c_oracle
    from oracle import stable_path as SP
    g = torch.Generator().manual_seed(LH * 31 + M)
    Bn = 2
    qk = torch.randn(Bn, LH, M, 1504, generator=g) * 3.0
    lib = L.lib()
    R = M - 1 - S
    ldm = (F + 3) // 4 * 4
    qkc = qk.cuda()
    out = torch.empty(Bn, R, ldm, device="cuda")
    ws = torch.empty(lib.stb_qkpost_dynamic_ws_bytes(Bn, LH, R, F, count), dtype=torch.uint8, device="cuda")
    jumps_dev = None
    ref_jumps = [None] * Bn
    for it in range(iters):
        L.check(lib.stb_qk_postprocess_dynamic(L.ptr(qkc), Bn, LH, M, 1504, S, R, F, 1.0, 7, count, L.ptr(jumps_dev), int(it > 0),
                                               L.ptr(out), ldm, L.ptr(ws), ws.numel(), L.stream_ptr()))
        torch.cuda.synchronize()
        new_jumps = []
        for b in range(Bn):
            qks = [qk[b:b + 1, h:h + 1] for h in range(LH)]          # one head per "layer"
            w = SP.attention_weights_dynamic(qks, S, F * 320, count, ref_jumps[b])
            ref = w.mean(dim=0)
            err = (out[b, :, :F].cpu() - ref).abs().max().item()
            assert err < 5e-5, (it, b, err)
            ref_jumps[b] = SP.jumps_from_matrix(ref)
            new_jumps.append(torch.from_numpy(c_oracle.dtw(-out[b, :, :F].cpu().numpy())[1]))
            assert np.array_equal(new_jumps[-1].numpy(), ref_jumps[b])
        jumps_dev = torch.stack(new_jumps).to(torch.int32).cuda()
    print(f"dynamic heads LH={LH} M={M} F={F} count={count} iters={iters}: ok")


@pytest.mark.parametrize("LH,H,M,F,S,topk,wcov", [(8, 2, 30, 500, 1, 5, 0.0), (24, 4, 44, 937, 3, 20, 0.0), (12, 3, 20, 1500, 3, 6, 0.5)])
def test_qk_postprocess_new_aligner_vs_oracle(L, LH, H, M, F, S, topk, wcov):
    from oracle import stable_path as SP
    g = torch.Generator().manual_seed(LH * 17 + M)
    Bn = 2
    qk = torch.randn(Bn, LH, M, 1504, generator=g) * 3.0
    lib = L.lib()
    R = M - 1 - S
    ldm = (F + 3) // 4 * 4
    qkc = qk.cuda()
    out = torch.empty(Bn, R, ldm, device="cuda")
    ws = torch.empty(lib.stb_qkpost_new_ws_bytes(Bn, LH, M, F, topk), dtype=torch.uint8, device="cuda")
    L.check(lib.stb_qk_postprocess_new(L.ptr(qkc), Bn, LH, M, 1504, S, R, F, 1.0, 7, topk, 1.0, 1.0, wcov, L.ptr(out), ldm,
                                       L.ptr(ws), ws.numel(), L.stream_ptr()))
    torch.cuda.synchronize()
    for b in range(Bn):
        qks = [qk[b:b + 1, l * H:(l + 1) * H] for l in range(LH // H)]
        ref = SP.attention_matrix_new(qks, S, F * 320, topk=topk, w_coverage=wcov)
        err = (out[b, :, :F].cpu() - ref).abs().max().item() / ref.abs().max().item()
        print(f"new aligner LH={LH} M={M} F={F} topk={topk} wcov={wcov}: rel err {err:.2e}")
        assert err < 1e-4
