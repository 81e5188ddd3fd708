"""CPU oracle: restatement of ``whisper.timing`` {median_filter, dtw, merge_punctuations} (openai-whisper 20250625).

TEST INFRASTRUCTURE.  Reference call sites: stable_whisper/timing.py:110,138 (median_filter), :195 (dtw),
:468 (merge_punctuations); stable_whisper/alignment.py:946.

Only the CPU semantics are restated (the dependency's Triton CUDA DTW uses a different tie rule and is NOT the
oracle).
"""
import numpy as np
import torch
import torch.nn.functional as F

try:                                    # numba is what the dependency uses; optional here
    import numba
    _jit = numba.jit(nopython=True)
except Exception:                       # pragma: no cover
    numba = None

    def _jit(f):
        return f


def median_filter(x: torch.Tensor, filter_width: int) -> torch.Tensor:
    """Median of width ``filter_width`` along the last dim with reflect padding."""
    pad_width = filter_width // 2
    if x.shape[-1] <= pad_width:
        return x                        # reflect padding impossible: returned untouched
    ndim = x.ndim
    if ndim <= 2:
        x = x[None, None, :]
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    x = F.pad(x, (pad_width, pad_width, 0, 0), mode="reflect")
    result = x.unfold(-1, filter_width, 1).sort()[0][..., pad_width]
    if ndim <= 2:
        result = result[0, 0]
    return result


@_jit
def _backtrace(trace: np.ndarray):
    i = trace.shape[0] - 1
    j = trace.shape[1] - 1
    trace[0, :] = 2
    trace[:, 0] = 1
    out = []
    while i > 0 or j > 0:
        out.append((i - 1, j - 1))
        t = trace[i, j]
        if t == 0:
            i -= 1
            j -= 1
        elif t == 1:
            i -= 1
        elif t == 2:
            j -= 1
        else:
            raise ValueError("Unexpected trace[i, j]")
    res = np.array(out)
    return res[::-1, :].T


@_jit
def _dtw_cpu(x: np.ndarray):
    """x: float64 [N, M].  Cost is kept in float32; column-major sweep; strict '<' tie rule."""
    N, M = x.shape
    cost = np.ones((N + 1, M + 1), dtype=np.float32) * np.inf
    trace = -np.ones((N + 1, M + 1), dtype=np.float32)
    cost[0, 0] = 0
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0 = cost[i - 1, j - 1]
            c1 = cost[i - 1, j]
            c2 = cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = x[i - 1, j - 1] + c
            trace[i, j] = t
    return _backtrace(trace)


def dtw_cpu(x: np.ndarray) -> np.ndarray:
    return _dtw_cpu(np.ascontiguousarray(x, dtype=np.float64))


def dtw(x: torch.Tensor) -> np.ndarray:
    """-> int array [2, path_len] = (text_indices, time_indices)."""
    return dtw_cpu(x.double().cpu().numpy())


def merge_punctuations(alignment, prepended: str, appended: str):
    """In-place merge of punctuation-only entries into neighbours; emptied entries keep word='' tokens=[]."""
    i = len(alignment) - 2
    j = len(alignment) - 1
    while i >= 0:
        previous, following = alignment[i], alignment[j]
        if previous.word.startswith(" ") and previous.word.strip() in prepended:
            following.word = previous.word + following.word
            following.tokens = previous.tokens + following.tokens
            previous.word = ""
            previous.tokens = []
        else:
            j = i
        i -= 1
    i, j = 0, 1
    while j < len(alignment):
        previous, following = alignment[i], alignment[j]
        if not previous.word.endswith(" ") and following.word in appended:
            previous.word = previous.word + following.word
            previous.tokens = previous.tokens + following.tokens
            following.word = ""
            following.tokens = []
        else:
            i = j
        j += 1
