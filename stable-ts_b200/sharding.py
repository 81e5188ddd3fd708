"""Multi-GPU sharding of independent 30 s windows: one process per GPU, static contiguous
ranges of windows per rank, full weight replica per rank, NO data-path collective, and ONE all-gather of fixed-stride
word records at the end (NCCL over NVLink on the GPU box; the same code runs over gloo on CPU in tests).

Record buffer (int32, one per rank, identical capacity on every rank so a single all_gather suffices):
    [0]                      number of words n
    [1 : 1+6*cap_words]      n x (window_id, start_ms, end_ms, n_tokens, probability as fp32 bits, segment index in the window)
    [1+6*cap_words : ]       token ids of the words, concatenated

The gathered records ARE the result wire format: ``gathered_to_result`` rebuilds, on any rank,
the dict of ``WhisperResult.to_dict`` (stable_whisper/result.py:618-636 segment dict, :1398-1406 result dict) -- word text is
re-decoded from the token ids, segment text / span / tokens from the words -- and ``result.make_result`` turns it into the
reference's ``WhisperResult`` (or the schema-compatible stand-in).
"""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

MAX_TOKENS_PER_WINDOW = 448
REC = 6                      # int32 fields per word record


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first n % world ranks get one extra window."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def capacity(n_windows_total: int, world: int) -> Tuple[int, int]:
    per_rank = -(-n_windows_total // world)
    return per_rank * MAX_TOKENS_PER_WINDOW, per_rank * MAX_TOKENS_PER_WINDOW      # words <= tokens


def pack_records(results: List[List[dict]], first_window: int, cap_words: int, cap_tokens: int) -> torch.Tensor:
    buf = np.zeros(1 + REC * cap_words + cap_tokens, dtype=np.int32)
    flat = [(first_window + w, wd) for w, words in enumerate(results) for wd in words]
    n = len(flat)
    tok_lists = [wd["tokens"] for _, wd in flat]
    counts = np.fromiter((len(t) for t in tok_lists), dtype=np.int32, count=n)
    t = int(counts.sum())
    assert n <= cap_words and t <= cap_tokens, "word record capacity exceeded"
    if n:
        recs = buf[1:1 + REC * n].reshape(n, REC)
        recs[:, 0] = np.fromiter((w for w, _ in flat), dtype=np.int32, count=n)
        recs[:, 1] = np.rint(np.fromiter((wd["start"] for _, wd in flat), dtype=np.float64, count=n) * 1000.0).astype(np.int32)
        recs[:, 2] = np.rint(np.fromiter((wd["end"] for _, wd in flat), dtype=np.float64, count=n) * 1000.0).astype(np.int32)
        recs[:, 3] = counts
        recs[:, 4] = np.fromiter((wd["probability"] for _, wd in flat), dtype=np.float32, count=n).view(np.int32)
        recs[:, 5] = np.fromiter((wd.get("segment", 0) for _, wd in flat), dtype=np.int32, count=n)
        buf[1 + REC * cap_words:1 + REC * cap_words + t] = np.fromiter((v for tl in tok_lists for v in tl), dtype=np.int32, count=t)
    buf[0] = n
    return torch.from_numpy(buf)


class GatheredWords(Sequence):
    """Word records of all windows as gathered (columnar, zero-copy views of the rank buffers).  Behaves like the
    list-of-lists ``unpack_records`` returns, but the per-word dicts of a window are only built when that window is
    indexed: at 8 ranks x 120 windows a step gathers ~160 k words, and building every dict on every rank would cost
    ~0.3 s of Python per step for results most ranks never look at."""

    def __init__(self, bufs: Sequence[torch.Tensor], n_windows_total: int, cap_words: int, tokenizer=None):
        self._n = n_windows_total
        self._tk = tokenizer                               # given: every word dict also carries its decoded text ("word")
        self._parts = []                                   # (recs [n,5], token array, token end offsets)
        first = np.full(n_windows_total + 1, -1, dtype=np.int64)
        self._where = {}                                   # window -> (part index, first record, n records)
        self.n_words = 0
        for b in bufs:
            a = b.cpu().numpy()
            n = int(a[0])
            if n == 0:
                continue
            recs = a[1:1 + REC * n].reshape(n, REC)
            ends = np.cumsum(recs[:, 3], dtype=np.int64)
            toks = a[1 + REC * cap_words:1 + REC * cap_words + int(ends[-1])]
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
        out = [dict(start=s0, end=e0, tokens=tl[a0 - base:a1 - base], probability=p, segment=sg)
               for s0, e0, p, sg, a0, a1 in zip((r[:, 1] / 1000.0).tolist(), (r[:, 2] / 1000.0).tolist(), probs, r[:, 5].tolist(),
                                                lo, hi)]
        if self._tk is not None:
            for w in out:
                w["word"] = self._tk.decode(w["tokens"])
        return out


def unpack_records(bufs: Sequence[torch.Tensor], n_windows_total: int, cap_words: int, lazy: bool = False, tokenizer=None):
    g = GatheredWords(bufs, n_windows_total, cap_words, tokenizer=tokenizer)
    return g if lazy else [g[i] for i in range(n_windows_total)]


def gathered_to_result(windows: Sequence[List[dict]], tokenizer, language: Optional[str] = None, window_seconds: float = 30.0) -> dict:
    """Per-window word lists (``GatheredWords`` or plain lists; words carry ``segment`` = index of their segment inside
    the window) -> dict(text, segments, language) with the keys of ``WhisperResult.to_dict``."""
    segments = []
    for wi in range(len(windows)):
        by_seg = {}
        for w in windows[wi]:
            by_seg.setdefault(int(w.get("segment", 0)), []).append(w)
        for si in sorted(by_seg):
            ws = [dict(word=w["word"] if "word" in w else tokenizer.decode(w["tokens"]), start=w["start"], end=w["end"],
                       probability=w["probability"], tokens=list(w["tokens"])) for w in by_seg[si]]
            segments.append(dict(start=ws[0]["start"], end=ws[-1]["end"], text="".join(w["word"] for w in ws),
                                 seek=round(wi * window_seconds, 3), tokens=[t for w in ws for t in w["tokens"]], temperature=0.0,
                                 avg_logprob=None, compression_ratio=None, no_speech_prob=None, words=ws, id=len(segments)))
    return dict(text="".join(s["text"] for s in segments), segments=segments,
                language=language or getattr(tokenizer, "language", None))


def run_sharded(process: Callable[[int, int], List[List[dict]]], n_windows_total: int, *, device: Optional[torch.device] = None,
                group=None, lazy: bool = False, tokenizer=None):
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
    return unpack_records(gathered, n_windows_total, cap_w, lazy=lazy, tokenizer=tokenizer)


def align_sharded(model, tokenizer, audios: Sequence[torch.Tensor], word_tokens: Sequence[List[List[int]]], group=None):
    """Data-parallel ``align_words_batch`` over the ranks of the default process group."""
    from .alignment import align_words_batch

    def process(lo, hi):
        if hi <= lo:
            return []
        return align_words_batch(model, tokenizer, list(audios[lo:hi]), list(word_tokens[lo:hi]))

    return run_sharded(process, len(audios), device=model.device, group=group)


def transcribe_sharded(model, tokenizer, audio: torch.Tensor, *, group=None, **kw):
    """Data-parallel transcription of one long 16 kHz waveform (every rank holds it): 30 s shards are split into contiguous
    ranges per rank, each rank walks its shards with ``transcribe.transcribe``, ONE all_gather of word records, and every rank
    rebuilds the full result (BASELINE config 4: 8 h over 8 GPUs).  -> WhisperResult (result.make_result)."""
    from .result import make_result
    from .transcribe import N_SAMPLES, transcribe
    audio = audio.detach().float().flatten()
    n_win = max(1, -(-int(audio.numel()) // N_SAMPLES))

    def process(lo, hi):
        if hi <= lo:
            return []
        d = transcribe(model, tokenizer, audio[lo * N_SAMPLES: hi * N_SAMPLES], **kw)
        out = [[] for _ in range(hi - lo)]
        for si, seg in enumerate(d["segments"]):
            wi = min(int(seg["seek"] // 30.0), hi - lo - 1)
            for w in seg.get("words") or []:
                out[wi].append(dict(w, start=w["start"] + lo * 30.0, end=w["end"] + lo * 30.0, segment=si))
        return out

    words = run_sharded(process, n_win, device=model.device, group=group, tokenizer=tokenizer)
    return make_result(gathered_to_result(words, tokenizer, getattr(tokenizer, "language", None)))
