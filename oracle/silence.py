"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's non-VAD silence detection, SURVEY.md section 8(f)
row 1: the step either side of the hot path that gates every transcribe window.

Follows, line by line:
    audio2loudness     /root/reference/stable_whisper/stabilization/nonvad.py:16-41
    wav2mask           /root/reference/stable_whisper/stabilization/nonvad.py:44-88
    mask2timing        /root/reference/stable_whisper/stabilization/utils.py:43-86
    timing2mask        /root/reference/stable_whisper/stabilization/utils.py:89-111
    predict_with_nonvad / _silent_mask_test   /root/reference/stable_whisper/stabilization/__init__.py:98-104,238-252

The float arithmetic of the two PyTorch operators on the path is restated exactly as this build's PyTorch CPU kernels
evaluate it (pinned in tests/test_oracle_silence.py against torch itself and against the unmodified reference functions):
    F.interpolate(mode='linear', align_corners=False):  out = fma(l0, x[i0], rn(l1 * x[i1])),
        src = max(scale * (i + 0.5) - 0.5, 0) in fp32, scale = fp32(n) / fp32(size)
    avg_pool1d(k, stride 1) after reflect padding:       out = (((((0 + a0) + a1) + a2) + a3) + a4) / k   in fp32
Pinning: tests/test_oracle_silence.py (live reference import when /root/reference exists) and the fixtures
tests/golden/silence_cases.npz written by oracle/make_golden_silence.py from the unmodified reference.
"""
from typing import Optional, Tuple

import numpy as np

TOKENS_PER_SECOND = 50            # whisper.audio: SAMPLE_RATE // N_SAMPLES_PER_TOKEN
N_SAMPLES_PER_TOKEN = 320
FRAMES_PER_SECOND = 100


def kth_largest_abs(audio: np.ndarray, k: int) -> np.float32:
    """top_values[-1] of torch.topk(audio.abs(), k): the k-th largest magnitude (nonvad.py:22-26)"""
    a = np.abs(audio.astype(np.float32))
    return np.partition(a, len(a) - k)[len(a) - k]


def interpolate_linear(x: np.ndarray, size: int) -> np.ndarray:
    """F.interpolate(x[None, None], size=size, mode='linear', align_corners=False)[0, 0] on fp32 (nonvad.py:34-39)"""
    n = len(x)
    scale = np.float32(n) / np.float32(size)
    i = np.arange(size, dtype=np.float32)
    src = np.maximum(scale * (i + np.float32(0.5)) - np.float32(0.5), np.float32(0))
    i0 = src.astype(np.int64)
    i1 = i0 + (i0 < n - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    t = (l1 * x[i1]).astype(np.float32)
    return (l0.astype(np.float64) * x[i0].astype(np.float64) + t.astype(np.float64)).astype(np.float32)   # one rounding = fma


def audio2loudness(audio: np.ndarray, samples_per_unit: Optional[int] = None) -> Optional[np.ndarray]:
    """nonvad.py:16-41.  -> fp32 [round(n / 320) + 1] or None when that is <= 2 tokens."""
    assert audio.ndim == 1
    a = np.abs(audio.astype(np.float32))
    k = int(a.size * 0.001)
    if k:
        threshold = np.partition(a, a.size - k)[a.size - k]
    else:
        import torch
        threshold = np.float32(torch.from_numpy(a).quantile(0.999, dim=-1).item())
    if samples_per_unit is None:
        samples_per_unit = N_SAMPLES_PER_TOKEN
    token_count = round(a.shape[-1] / samples_per_unit) + 1
    if token_count > 2:
        if threshold < 1e-5:
            return np.zeros(token_count, dtype=np.float32)
        t175 = np.float32(threshold) * np.float32(1.75)
        denom = np.float32(1.0) if 1.0 < t175 else t175               # Python min(1., tensor): the tensor unless 1. is smaller
        return interpolate_linear((a / denom).astype(np.float32), token_count)
    return None


def avg_pool_reflect(x: np.ndarray, k_size: int) -> np.ndarray:
    p = k_size // 2
    a = np.pad(x.astype(np.float32), (p, p), mode="reflect")
    s = np.zeros(len(x), dtype=np.float32)
    for j in range(k_size):
        s = (s + a[j:j + len(x)]).astype(np.float32)
    return (s / np.float32(k_size)).astype(np.float32)


def loudness_to_raw_mask(loudness: np.ndarray, q_levels: int = 20, k_size: int = 5) -> np.ndarray:
    """nonvad.py:58-78: smoothing, quantisation, -> bool (True = sound).  This is the part the CUDA kernel produces."""
    p = k_size // 2 if k_size else 0
    if p and p < loudness.shape[-1]:
        assert k_size % 2, f"kernel_size must be odd but got {k_size}"
        mask = avg_pool_reflect(loudness, k_size)
    else:
        mask = loudness.copy()
    if q_levels:
        mask = np.rint((mask * np.float32(q_levels)).astype(np.float32))
    return mask.astype(bool)


def mask2timing(silence_mask: Optional[np.ndarray], time_offset: float = 0.0, second_per_unit: Optional[float] = None,
                min_start: Optional[float] = None, max_end: Optional[float] = None):
    """stabilization/utils.py:43-86"""
    if silence_mask is None or not silence_mask.any() or not len(silence_mask):
        return None
    assert silence_mask.ndim == 1
    mask = np.concatenate(([False], silence_mask, [False]))
    silent_starts = np.logical_and(~mask[:-2], mask[1:-1]).nonzero()[0]
    silent_ends = (np.logical_and(mask[1:-1], ~mask[2:]).nonzero()[0] + 1)
    clipped = False
    if second_per_unit is None:
        silent_starts = silent_starts / TOKENS_PER_SECOND
        silent_ends = silent_ends / TOKENS_PER_SECOND
    else:
        silent_starts = silent_starts * second_per_unit
        silent_ends = silent_ends * second_per_unit
    if time_offset:
        silent_starts += time_offset
        silent_ends += time_offset
    if min_start is not None and silent_starts[0] < min_start:
        silent_starts.clip(min_start, None, silent_starts)
        clipped = True
    if max_end is not None and silent_ends[-1] > max_end:
        silent_ends.clip(None, max_end, silent_ends)
        clipped = True
    if clipped:
        invalid = silent_starts >= silent_ends
        if invalid.any():
            if invalid.all():
                return None
            silent_starts, silent_ends = silent_starts[~invalid], silent_ends[~invalid]
    return silent_starts, silent_ends


def timing2mask(silent_starts: np.ndarray, silent_ends: np.ndarray, size: int, time_offset: Optional[float] = None,
                units_per_second: Optional[int] = None) -> np.ndarray:
    """stabilization/utils.py:89-111"""
    if units_per_second is None:
        units_per_second = TOKENS_PER_SECOND
    assert len(silent_starts) == len(silent_ends)
    out = np.zeros(size, dtype=bool)
    if time_offset:
        silent_starts = (silent_starts - time_offset).clip(min=0)
        silent_ends = (silent_ends - time_offset).clip(min=0)
    mask_i = (silent_starts * units_per_second).round().astype(np.int32)
    mask_e = (silent_ends * units_per_second).round().astype(np.int32)
    for mi, me in zip(mask_i, mask_e):
        out[mi:me + 1] = True
    return out


def raw_mask_to_silence_mask(mask: np.ndarray) -> Optional[np.ndarray]:
    """nonvad.py:76-88: sound mask -> suppression mask (True = silent timestamp token), None when nothing is silent"""
    if not mask.any():                                   # entirely silent
        return ~mask
    s, e = mask2timing(mask)
    keep = (e - s) > 0.1
    out = ~timing2mask(s[keep], e[keep], mask.shape[-1])
    if not out.any():                                    # no silence
        return None
    return out


def wav2mask(audio: np.ndarray, q_levels: int = 20, k_size: int = 5) -> Optional[np.ndarray]:
    """nonvad.py:44-88 for 16 kHz mono fp32 input (the resampling front door is out of scope)"""
    loud = audio2loudness(audio)
    if loud is None:
        return None
    return raw_mask_to_silence_mask(loudness_to_raw_mask(loud, q_levels, k_size))


def silent_mask_test(mask: Optional[np.ndarray], min_unit_per_word: int) -> bool:
    """stabilization/__init__.py:98-104"""
    if mask is None:
        return False
    return bool(mask.shape[-1] - np.count_nonzero(mask) < min_unit_per_word)


def pad_mask(mask: Optional[np.ndarray], length: int = 1501) -> Optional[np.ndarray]:
    """mask_pad_func = whisper.audio.pad_or_trim(mask, 1501) (original_whisper.py:427-429, stabilization/__init__.py:138-143)"""
    if mask is None:
        return None
    if mask.shape[-1] > length:
        return mask[:length]
    if mask.shape[-1] < length:
        return np.pad(mask, (0, length - mask.shape[-1]))
    return mask


def predict_with_nonvad(audio: np.ndarray, offset: Optional[float] = None, q_levels: int = 20, k_size: int = 5,
                        min_word_dur: float = 0.1) -> dict:
    """stabilization/__init__.py:238-252 with the predictor settings transcribe uses (original_whisper.py:427-440):
    -> dict(timings [2, n] | None, mask bool [1501] | None, is_silent)"""
    mask = wav2mask(audio, q_levels=q_levels, k_size=k_size)
    timings = mask2timing(mask, time_offset=offset)
    if timings is not None:
        timings = np.stack(timings, axis=0)
    min_frames_per_word = max(round(min_word_dur * FRAMES_PER_SECOND), 1)
    is_silent = silent_mask_test(mask, min_frames_per_word)
    return dict(timings=timings, mask=pad_mask(mask), is_silent=is_silent)
