"""Multi-GPU sharding of independent 30 s windows: one process per GPU, static contiguous
ranges of windows per rank, full weight replica per rank, NO data-path collective, and ONE all-gather of fixed-stride
word records at the end (NCCL over NVLink on the GPU box; the same code runs over gloo on CPU in tests).

Record buffer (int32, one per rank, identical capacity on every rank so a single all_gather suffices):
    [0]                      number of words n
    [1 : 1+5*cap_words]      n x (window_id, start_ms, end_ms, n_tokens, probability as fp32 bits)
    [1+5*cap_words : ]       token ids of the words, concatenated
"""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

MAX_TOKENS_PER_WINDOW = 448


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first n % world ranks get one extra window."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def capacity(n_windows_total: int, world: int) -> Tuple[int, int]:
    per_rank = -(-n_windows_total // world)
    return per_rank * MAX_TOKENS_PER_WINDOW, per_rank * MAX_TOKENS_PER_WINDOW      # words <= tokens


def pack_records(results: List[List[dict]], first_window: int, cap_words: int, cap_tokens: int) -> torch.Tensor:
    buf = np.zeros(1 + 5 * cap_words + cap_tokens, dtype=np.int32)
    flat = [(first_window + w, wd) for w, words in enumerate(results) for wd in words]
    n = len(flat)
    tok_lists = [wd["tokens"] for _, wd in flat]
    counts = np.fromiter((len(t) for t in tok_lists), dtype=np.int32, count=n)
    t = int(counts.sum())
    assert n <= cap_words and t <= cap_tokens, "word record capacity exceeded"
    if n:
        recs = buf[1:1 + 5 * n].reshape(n, 5)
        recs[:, 0] = np.fromiter((w for w, _ in flat), dtype=np.int32, count=n)
        recs[:, 1] = np.rint(np.fromiter((wd["start"] for _, wd in flat), dtype=np.float64, count=n) * 1000.0).astype(np.int32)
        recs[:, 2] = np.rint(np.fromiter((wd["end"] for _, wd in flat), dtype=np.float64, count=n) * 1000.0).astype(np.int32)
        recs[:, 3] = counts
        recs[:, 4] = np.fromiter((wd["probability"] for _, wd in flat), dtype=np.float32, count=n).view(np.int32)
        buf[1 + 5 * cap_words:1 + 5 * cap_words + t] = np.fromiter((v for tl in tok_lists for v in tl), dtype=np.int32, count=t)
    buf[0] = n
    return torch.from_numpy(buf)


class GatheredWords(Sequence):
    """Word records of all windows as gathered (columnar, zero-copy views of the rank buffers).  Behaves like the
    list-of-lists ``unpack_records`` returns, but the per-word dicts of a window are only built when that window is
    indexed: at 8 ranks x 120 windows a step gathers ~160 k words, and building every dict on every rank would cost
    ~0.3 s of Python per step for results most ranks never look at."""

    def __init__(self, bufs: Sequence[torch.Tensor], n_windows_total: int, cap_words: int):
        self._n = n_windows_total
        self._parts = []                                   # (recs [n,5], token array, token end offsets)
        first = np.full(n_windows_total + 1, -1, dtype=np.int64)
        self._where = {}                                   # window -> (part index, first record, n records)
        self.n_words = 0
        for b in bufs:
            a = b.cpu().numpy()
            n = int(a[0])
            if n == 0:
                continue
            recs = a[1:1 + 5 * n].reshape(n, 5)
            ends = np.cumsum(recs[:, 3], dtype=np.int64)
            toks = a[1 + 5 * cap_words:1 + 5 * cap_words + int(ends[-1])]
            pi = len(self._parts)
            self._parts.append((recs, toks, ends))
            wins, starts, counts = np.unique(recs[:, 0], return_index=True, return_counts=True)   # records are window-sorted
            for w, s0, c in zip(wins.tolist(), starts.tolist(), counts.tolist()):
                self._where[w] = (pi, s0, c)
            self.n_words += n
        del first

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        loc = self._where.get(i)
        if loc is None:
            return []
        pi, r0, c = loc
        recs, toks, ends = self._parts[pi]
        r = recs[r0:r0 + c]
        hi = ends[r0:r0 + c].tolist()
        lo = [int(ends[r0 - 1]) if r0 > 0 else 0] + hi[:-1]
        tl = toks[lo[0]:hi[-1]].tolist()
        base = lo[0]
        probs = np.ascontiguousarray(r[:, 4]).view(np.float32).astype(np.float64).tolist()
        return [dict(start=s0, end=e0, tokens=tl[a0 - base:a1 - base], probability=p)
                for s0, e0, p, a0, a1 in zip((r[:, 1] / 1000.0).tolist(), (r[:, 2] / 1000.0).tolist(), probs, lo, hi)]


def unpack_records(bufs: Sequence[torch.Tensor], n_windows_total: int, cap_words: int, lazy: bool = False):
    g = GatheredWords(bufs, n_windows_total, cap_words)
    return g if lazy else [g[i] for i in range(n_windows_total)]


def run_sharded(process: Callable[[int, int], List[List[dict]]], n_windows_total: int, *, device: Optional[torch.device] = None,
                group=None, lazy: bool = False):
    """``process(lo, hi)`` computes the word lists of windows [lo, hi) on this rank; every rank returns the merged
    result for all windows (``lazy=True``: a ``GatheredWords`` view that builds a window's dicts on access).  The only
    collective is one all_gather of the record buffer."""
    if not (dist.is_available() and dist.is_initialized()):
        return process(0, n_windows_total)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_windows_total, rank, world)
    local = process(lo, hi)
    cap_w, cap_t = capacity(n_windows_total, world)
    buf = pack_records(local, lo, cap_w, cap_t)
    if device is not None:
        buf = buf.to(device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    return unpack_records(gathered, n_windows_total, cap_w, lazy=lazy)


def align_sharded(model, tokenizer, audios: Sequence[torch.Tensor], word_tokens: Sequence[List[List[int]]], group=None):
    """Data-parallel ``align_words_batch`` over the ranks of the default process group."""
    from .alignment import align_words_batch

    def process(lo, hi):
        if hi <= lo:
            return []
        return align_words_batch(model, tokenizer, list(audios[lo:hi]), list(word_tokens[lo:hi]))

    return run_sharded(process, len(audios), device=model.device, group=group)
