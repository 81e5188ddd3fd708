"""Developer micro-benchmarks of single kernels through the C ABI (CUDA-event timed, cold weights: the working set of
rotating weight copies exceeds the 126 MB L2 so every launch streams its weights from HBM, as in a real decode step).

    python tools/microbench.py gemv      # decode-step linear shapes of large-v3 at 16 / 64 sequences
    python tools/microbench.py dtw       # DTW at 16 / 64 windows x [101 x 1500]
    python tools/microbench.py step 120 4   # ms per decode step (120 windows, 4 decoder layers of large-v3 width)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_ts_b200 import _lib as L  # noqa: E402


def _time(fn, iters, graph=True):
    """us per call; the calls are captured into one CUDA graph (no Python / ctypes launch overhead in the timing)"""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    if graph:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(iters):
                    fn(i)
        g.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if graph else 1
    e0.record()
    for _ in range(reps):
        if graph:
            g.replay()
        else:
            for i in range(iters):
                fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3      # us


def bench_splitk(batches=(16, 64), iters=40):
    """decode-step linear as a SWAPPED split-K tcgen05 GEMM through stb_gemm: A = W (features on the 128-row M side), B = x
    (sequences on the N side), the K range cut into `split` slices mapped on the GEMM batch axis; partials [split][B][N]."""
    import ctypes
    lib = L.lib()
    out = []
    for Bn in batches:
        for (N, K, split) in ((1280, 1280, 10), (1280, 1280, 5), (3840, 1280, 5), (5120, 1280, 4), (1280, 5120, 16),
                              (1280, 5120, 8), (51866, 1280, 1)):
            ks = K // split
            copies = max(2, int(300e6 // (N * K * 4)) + 1)
            W = torch.randn(copies, N, K, device="cuda") * 0.05
            wh = W.half()
            wl = (W - wh.float()).half()
            X = torch.randn(Bn, K, device="cuda")
            xh = X.half()
            xl = (X - xh.float()).half()
            P = torch.zeros(split, Bn, N, device="cuda")

            def run(i):
                c = i % copies
                a = L.Operand(L.ptr(wh[c]), L.ptr(wl[c]), N, ks, K, 0, ks)
                b = L.Operand(L.ptr(xh), L.ptr(xl), Bn, ks, K, 0, ks)
                ep = L.Epilogue()
                ep.out_f32 = L.ptr(P)
                ep.ld_out = N
                ep.out_b_stride = Bn * N
                ep.transposed = 1
                ep.alpha = 1.0
                L.check(lib.stb_gemm(ctypes.byref(a), ctypes.byref(b), split, 1, ctypes.byref(ep), L.stream_ptr()))
            run(0)
            torch.cuda.synchronize()
            ref = X.double() @ W[0].double().T
            err = ((P.double().sum(0) - ref).abs().max() / ref.abs().max()).item()
            us = _time(run, iters)
            out.append(dict(B=Bn, N=N, K=K, split=split, us=round(us, 2), weight_GBps=round(N * K * 4 / us / 1e3, 1), rel_err=err))
            print(out[-1], flush=True)
            del W, wh, wl
    return out


def bench_gemv(batches=(16, 64), shapes=((1280, 1280), (3840, 1280), (5120, 1280), (1280, 5120), (51866, 1280)), iters=200):
    lib = L.lib()
    out = []
    for Bn in batches:
        for (N, K) in shapes:
            copies = max(2, int(300e6 // (N * K * 4)) + 1)
            wh = torch.randn(copies, N, K, device="cuda").half()
            wl = torch.randn(copies, N, K, device="cuda").half() * 1e-3
            xh = torch.randn(Bn, K, device="cuda").half()
            xl = torch.randn(Bn, K, device="cuda").half() * 1e-3
            ld = (N + 7) // 8 * 8
            o = torch.empty(Bn, ld, device="cuda")

            def run(i):
                c = i % copies
                L.check(lib.stb_gemv(L.ptr(xh), L.ptr(xl), Bn, K, L.ptr(wh[c]), L.ptr(wl[c]), N, None, 0, None, 0, L.ptr(o),
                                     None, None, ld, L.stream_ptr()))
            us = _time(run, min(iters, 40))
            gbs = N * K * 4 / us / 1e3
            out.append(dict(B=Bn, N=N, K=K, us=round(us, 2), weight_GBps=round(gbs, 1)))
            print(out[-1], flush=True)
            del wh, wl
    return out


def bench_dtw(batches=(16, 64, 148)):
    lib = L.lib()
    out = []
    for Bn in batches:
        R, F = 101, 1500
        x = torch.randn(Bn, R, F, device="cuda")
        jumps = torch.zeros(Bn, R, dtype=torch.int32, device="cuda")
        ws = torch.empty(lib.stb_dtw_ws_bytes(Bn, R, F), dtype=torch.uint8, device="cuda")

        def run(i):
            L.check(lib.stb_dtw(L.ptr(x), Bn, R, F, F, 1, L.ptr(jumps), None, None, L.ptr(ws), ws.numel(), L.stream_ptr()))
        us = _time(run, 50)
        out.append(dict(windows=Bn, R=R, F=F, us=round(us, 1)))
        print(out[-1], flush=True)
    return out


def bench_step(windows=120, layers=4, s1=24, s2=72):
    """ms per KV-cached decode step of a large-v3-WIDTH model with `layers` decoder layers (cross K/V of 120 windows x 32
    layers would not leave room for A/B runs), from the slope between two forced decodes of s1 and s2 steps (graph replay).
    Use with STB_DECODE_SPLITK_LEGACY=1 / STB_STEP_* to A/B the decode-step variants:  microbench.py step [windows] [layers]"""
    import time
    from stable_ts_b200.api import random_state_dict
    from stable_ts_b200.decode import DecodingOptions, decode_windows
    from stable_ts_b200.model import B200Whisper
    from stable_ts_b200.tokenizer import get_tokenizer
    from oracle.whisper_ref.model import ModelDimensions      # dims container only (tools/ is developer tooling)
    dims = ModelDimensions(128, 1500, 1280, 20, 1, 51866, 448, 1280, 20, layers)
    model = B200Whisper(dims, random_state_dict(dims, 0), device="cuda")
    tk = get_tokenizer(model, language="en", task="transcribe", synthetic=True)
    g = torch.Generator().manual_seed(0)
    audio = (torch.randn(windows, 480000, generator=g) * 0.1).cuda()
    enc = model.encode(model.log_mel(audio))
    forced = torch.randint(300, 50000, (s2, windows), generator=g, dtype=torch.int32)

    def run(steps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        decode_windows(model, tk, enc, DecodingOptions(language="en", sample_len=steps), forced_tokens=forced[:steps],
                       reuse_buffers=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t) * 1e3
    run(s1)
    a = min(run(s1) for _ in range(3))
    b = min(run(s2) for _ in range(3))
    per_step = (b - a) / (s2 - s1)
    L.lib().stb_prof_enable(1)
    decode_windows(model, tk, enc, DecodingOptions(language="en", sample_len=8), forced_tokens=forced[:8], use_graph=False,
                   reuse_buffers=True)
    prof = L.prof_report()
    L.lib().stb_prof_enable(0)
    table = {k: {"n": v["n"], "us_avg": round(v["ms"] * 1e3 / max(v["n"], 1), 2)} for k, v in prof.items()}
    out = dict(windows=windows, layers=layers, ms_per_step=round(per_step, 3), ms_per_layer_step=round(per_step / layers, 4),
               fixed_ms=round(a - per_step * s1, 1), eager_kernels=table)
    print(out, flush=True)
    return out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gemv"
    if what == "step":
        res = bench_step(*(int(v) for v in sys.argv[2:4]))
    elif what == "ncu":          # short run for `ncu -k regex:...`: fc1 / fc2 at 64 sequences + one DTW batch
        res = [bench_gemv((64,), ((5120, 1280), (1280, 5120)), iters=4), bench_dtw((16,))]
    else:
        res = {"gemv": bench_gemv, "dtw": bench_dtw, "splitk": bench_splitk}[what]()
    print(json.dumps({what: res, "env": {k: v for k, v in os.environ.items() if k.startswith("STB_")}}))
