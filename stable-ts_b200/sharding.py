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
    recs = buf[1:1 + 5 * cap_words].reshape(cap_words, 5)
    toks = buf[1 + 5 * cap_words:]
    n = t = 0
    for w, words in enumerate(results):
        for wd in words:
            k = len(wd["tokens"])
            assert n < cap_words and t + k <= cap_tokens, "word record capacity exceeded"
            recs[n] = (first_window + w, int(round(wd["start"] * 1000)), int(round(wd["end"] * 1000)), k,
                       np.float32(wd["probability"]).view(np.int32))
            toks[t:t + k] = wd["tokens"]
            n += 1
            t += k
    buf[0] = n
    return torch.from_numpy(buf)


def unpack_records(bufs: Sequence[torch.Tensor], n_windows_total: int, cap_words: int) -> List[List[dict]]:
    out: List[List[dict]] = [[] for _ in range(n_windows_total)]
    for b in bufs:
        a = b.cpu().numpy()
        n = int(a[0])
        recs = a[1:1 + 5 * cap_words].reshape(cap_words, 5)[:n]
        toks = a[1 + 5 * cap_words:]
        t = 0
        for win, s_ms, e_ms, k, pbits in recs.tolist():
            out[win].append(dict(start=s_ms / 1000.0, end=e_ms / 1000.0, tokens=toks[t:t + k].tolist(),
                                 probability=float(np.int32(pbits).view(np.float32))))
            t += k
    return out


def run_sharded(process: Callable[[int, int], List[List[dict]]], n_windows_total: int, *, device: Optional[torch.device] = None,
                group=None) -> List[List[dict]]:
    """``process(lo, hi)`` computes the word lists of windows [lo, hi) on this rank; every rank returns the merged
    result for all windows.  The only collective is one all_gather of the record buffer."""
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
    return unpack_records(gathered, n_windows_total, cap_w)


def align_sharded(model, tokenizer, audios: Sequence[torch.Tensor], word_tokens: Sequence[List[List[int]]], group=None):
    """Data-parallel ``align_words_batch`` over the ranks of the default process group."""
    from .alignment import align_words_batch

    def process(lo, hi):
        if hi <= lo:
            return []
        return align_words_batch(model, tokenizer, list(audios[lo:hi]), list(word_tokens[lo:hi]))

    return run_sharded(process, len(audios), device=model.device, group=group)
