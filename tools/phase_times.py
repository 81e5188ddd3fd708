"""Where one transcribe step spends its wall time (host vs device phases), large-v3 by default:

    python tools/phase_times.py [--windows 120] [--model large-v3]

Every wrapped function is bracketed by torch.cuda.synchronize(), so the numbers add up to a SERIALISED step (slightly slower
than the real, asynchronous one); the point is the split between GPU phases and host bookkeeping.
"""
import argparse
import os
import sys
import time
from collections import defaultdict

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import stable_ts_b200.timing as TM  # noqa: E402
import stable_ts_b200.transcribe as TR  # noqa: E402
from stable_ts_b200.api import load_model  # noqa: E402
from stable_ts_b200.decode import DecodingOptions  # noqa: E402
from stable_ts_b200.tokenizer import get_tokenizer  # noqa: E402

ACC = defaultdict(float)
DEPTH = [0]


def wrap(mod, name, label=None):
    fn = getattr(mod, name)
    label = label or name

    def inner(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        DEPTH[0] += 1
        try:
            return fn(*a, **k)
        finally:
            DEPTH[0] -= 1
            torch.cuda.synchronize()
            ACC[label] += time.perf_counter() - t
    setattr(mod, name, inner)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=120)
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--tokens", type=int, default=224)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = load_model(a.model, device=dev, precision="fp16x3", seed=0)
    from stable_ts_b200.api import MODEL_DIMS
    model.alignment_head_pairs = bench.alignment_head_pairs(tuple(MODEL_DIMS[a.model]), 10)
    tk = get_tokenizer(model, language="en", task="transcribe", synthetic=True)
    audios, words = bench.make_windows(a.windows, a.tokens, tk.eot, seed0=1000)
    host = torch.stack(audios).pin_memory()
    script = torch.tensor([[t for w in wt for t in w] for wt in words], dtype=torch.int32).T.contiguous()
    opt = DecodingOptions(language="en", sample_len=a.tokens, max_initial_timestamp=None)
    for _ in range(2):
        TR.transcribe_windows(model, tk, host, options=opt, forced_tokens=script)
    wrap(model, "log_mel", "gpu: h2d + log_mel")
    wrap(model, "encode", "gpu: encoder")
    wrap(TR, "decode_windows", "gpu: decode loop (cross K/V + 224 graph steps)")
    wrap(TR, "slice_segments", "host: slice_segments")
    wrap(TM, "_prepare_word_timestamps", "host: split words / gap padding")
    wrap(TM, "align_windows", "gpu: forced decoder + QK post + DTW + probs (incl. d2h)")
    wrap(TM, "word_timings_from_jumps", "host: word timings")
    wrap(TM, "_finish_word_timestamps", "host: merge punctuation / write words")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        TR.transcribe_windows(model, tk, host, options=opt, forced_tokens=script)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / n
    print(f"{a.model}, {a.windows} windows: serialised step {total * 1e3:.1f} ms")
    for k, v in sorted(ACC.items(), key=lambda kv: -kv[1]):
        print(f"  {v / n * 1e3:9.1f} ms  {k}")
    print(f"  {(total - sum(ACC.values()) / n) * 1e3:9.1f} ms  (unattributed host glue)")


if __name__ == "__main__":
    main()
