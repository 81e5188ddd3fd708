"""Host-side mirror of the reference's non-VAD silence detection for BATCHES of windows.

    stable_whisper/stabilization/nonvad.py:16-88     audio2loudness, wav2mask
    stable_whisper/stabilization/utils.py:43-111     mask2timing, timing2mask
    stable_whisper/stabilization/__init__.py:98-104,238-252   NonSpeechPredictor._silent_mask_test / predict_with_nonvad

The per-sample work (k-th largest magnitude, down-sampling, moving average, quantisation) runs in ``stb_silence_mask`` for
all windows of the batch at once; what is left on the host is the run-length bookkeeping on <= 1501 booleans per window,
which the reference also keeps in numpy.  There is no CPU fallback: the audio must live on a CUDA device.

``transcribe_windows(..., suppress_ts_tokens=True)`` feeds ``predict_nonvad_batch``'s per-window masks to the sampler
(``ts_token_mask`` [B, 1501], original_whisper.py:504-511 / decode.py:14-16); bit-exact against the oracle and the
reference-written fixtures on hardware (tests/test_gpu_silence.py).
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

TOKENS_PER_SECOND = 50
N_SAMPLES_PER_TOKEN = 320
FRAMES_PER_SECOND = 100
MAX_TOKENS = 1501


def sound_masks(audio: torch.Tensor, q_levels: int = 20, k_size: int = 5, want_loudness: bool = False):
    """audio fp32 [B, n] on a CUDA device (16 kHz mono, every row the same length) ->
    (mask bool ndarray [B, token_count], True = sound; loudness fp32 ndarray [B, token_count] | None), or (None, None) when
    the windows are too short to have more than two tokens (nonvad.py:30,41)."""
    if audio.ndim == 1:
        audio = audio[None]
    if not audio.is_cuda:
        raise RuntimeError("stable_ts_b200.silence: audio must be on a CUDA device (there is no CPU path)")
    audio = audio.to(torch.float32)
    if audio.stride(-1) != 1:
        audio = audio.contiguous()
    B, n = audio.shape
    token_count = round(n / N_SAMPLES_PER_TOKEN) + 1
    if token_count <= 2:
        return None, None
    if token_count > MAX_TOKENS:
        raise ValueError(f"window of {n} samples is longer than 30 s")
    k = int(n * 0.001)
    thr_in = None
    if k == 0:                                             # < 1000 samples: the reference switches to a quantile
        thr_in = audio.abs().quantile(0.999, dim=-1).contiguous()
    mask = torch.empty(B, MAX_TOKENS, dtype=torch.uint8, device=audio.device)
    loud = torch.empty(B, MAX_TOKENS, dtype=torch.float32, device=audio.device) if want_loudness else None
    with torch.cuda.device(audio.device):
        L.check(L.lib().stb_silence_mask(L.ptr(audio), B, n, audio.stride(0), k, token_count, int(q_levels or 0), int(k_size or 0),
                                         L.ptr(thr_in), L.ptr(loud), L.ptr(mask), None, L.stream_ptr()))
    m = mask[:, :token_count].cpu().numpy().astype(bool)
    return m, (loud[:, :token_count].cpu().numpy() if want_loudness else None)


def mask2timing(silence_mask: Optional[np.ndarray], time_offset: float = 0.0, second_per_unit: Optional[float] = None,
                min_start: Optional[float] = None, max_end: Optional[float] = None):
    """(starts, ends) in seconds of the True runs of ``silence_mask`` (stabilization/utils.py:43-86)."""
    if silence_mask is None or not len(silence_mask) or not silence_mask.any():
        return None
    m = np.concatenate(([False], np.asarray(silence_mask, dtype=bool), [False]))
    starts = np.logical_and(~m[:-2], m[1:-1]).nonzero()[0]
    ends = np.logical_and(m[1:-1], ~m[2:]).nonzero()[0] + 1
    if second_per_unit is None:
        starts, ends = starts / TOKENS_PER_SECOND, ends / TOKENS_PER_SECOND
    else:
        starts, ends = starts * second_per_unit, ends * second_per_unit
    if time_offset:
        starts, ends = starts + time_offset, ends + time_offset
    clipped = False
    if min_start is not None and starts[0] < min_start:
        starts = starts.clip(min_start, None)
        clipped = True
    if max_end is not None and ends[-1] > max_end:
        ends = ends.clip(None, max_end)
        clipped = True
    if clipped:
        bad = starts >= ends
        if bad.any():
            if bad.all():
                return None
            starts, ends = starts[~bad], ends[~bad]
    return starts, ends


def timing2mask(silent_starts: np.ndarray, silent_ends: np.ndarray, size: int, time_offset: Optional[float] = None,
                units_per_second: Optional[int] = None) -> np.ndarray:
    """bool [size], True inside [start, end] of every silent span (stabilization/utils.py:89-111)."""
    ups = TOKENS_PER_SECOND if units_per_second is None else units_per_second
    out = np.zeros(size, dtype=bool)
    if time_offset:
        silent_starts = (silent_starts - time_offset).clip(min=0)
        silent_ends = (silent_ends - time_offset).clip(min=0)
    for mi, me in zip((silent_starts * ups).round().astype(np.int32), (silent_ends * ups).round().astype(np.int32)):
        out[mi:me + 1] = True
    return out


def _silence_from_sound(mask: np.ndarray) -> Optional[np.ndarray]:
    """nonvad.py:76-88: sound mask -> suppression mask (True = silent), None when the window has no silence"""
    if not mask.any():
        return ~mask
    s, e = mask2timing(mask)
    keep = (e - s) > 0.1
    out = ~timing2mask(s[keep], e[keep], mask.shape[-1])
    return out if out.any() else None


def wav2mask_batch(audio: torch.Tensor, q_levels: int = 20, k_size: int = 5) -> List[Optional[np.ndarray]]:
    """``wav2mask`` (nonvad.py:44-88) of every row of ``audio`` [B, n]: bool [token_count] (True = silent) or None."""
    sound, _ = sound_masks(audio, q_levels, k_size)
    B = 1 if audio.ndim == 1 else audio.shape[0]
    if sound is None:
        return [None] * B
    return [_silence_from_sound(sound[b]) for b in range(B)]


def predict_nonvad_batch(audio: torch.Tensor, offsets: Optional[Sequence[Optional[float]]] = None, q_levels: int = 20,
                         k_size: int = 5, min_word_dur: float = 0.1) -> List[dict]:
    """``NonSpeechPredictor.predict_with_nonvad`` (stabilization/__init__.py:238-252) with transcribe's settings
    (original_whisper.py:427-440) for every window: dict(timings ndarray [2, n] | None, mask torch.bool [1501] | None --
    the ``ts_token_mask`` the decode step consumes --, is_silent)."""
    masks = wav2mask_batch(audio, q_levels, k_size)
    min_frames_per_word = max(round(min_word_dur * FRAMES_PER_SECOND), 1)
    out = []
    for b, mask in enumerate(masks):
        off = None if offsets is None else offsets[b]
        timings = mask2timing(mask, time_offset=off if off else 0.0)
        if timings is not None:
            timings = np.stack(timings, axis=0)
        is_silent = False if mask is None else bool(mask.shape[-1] - np.count_nonzero(mask) < min_frames_per_word)
        padded = None
        if mask is not None:                                # mask_pad_func = pad_or_trim(mask, 1501)
            padded = torch.zeros(MAX_TOKENS, dtype=torch.bool)
            padded[: min(len(mask), MAX_TOKENS)] = torch.from_numpy(mask[:MAX_TOKENS])
        out.append(dict(timings=timings, mask=padded, is_silent=is_silent))
    return out
