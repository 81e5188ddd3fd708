"""CPU oracle: restatement of ``whisper.audio`` (openai-whisper 20250625).  TEST INFRASTRUCTURE.

Call sites in the reference: stable_whisper/whisper_word_level/original_whisper.py:528-530,
stable_whisper/alignment.py:411-413 (align), :660-661 (refine), :924-925 (locate).
Constants mirror stable_whisper/whisper_compatibility.py:82-90.
"""
from functools import lru_cache
from typing import Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE          # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH              # 3000
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2            # 320
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH   # 100
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN  # 50


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Zero-pad or cut ``array`` to ``length`` along ``axis`` (whisper_compatibility.py:218-241)."""
    if torch.is_tensor(array):
        n = array.shape[axis]
        if n > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if n < length:
            pads = [0, 0] * array.ndim
            # F.pad takes pairs starting from the LAST dim
            pads[2 * (array.ndim - 1 - (axis % array.ndim)) + 1] = length - n
            array = F.pad(array, pads)
    else:
        n = array.shape[axis]
        if n > length:
            array = array.take(indices=range(length), axis=axis)
        if n < length:
            widths = [(0, 0)] * array.ndim
            widths[axis] = (0, length - n)
            array = np.pad(array, widths)
    return array


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


@lru_cache(maxsize=None)
def mel_filterbank_np(n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels=n_mels) (Slaney scale, area-normalised).

    openai-whisper ships this matrix as ``assets/mel_filters.npz``; the asset is not available offline,
    so it is regenerated from the published recipe.  float32 [n_mels, 201].
    """
    n_freqs = N_FFT // 2 + 1
    fft_freqs = np.linspace(0.0, SAMPLE_RATE / 2.0, n_freqs)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(SAMPLE_RATE / 2.0), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fft_freqs[None, :]
    weights = np.zeros((n_mels, n_freqs), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def mel_filters(device, n_mels: int) -> torch.Tensor:
    assert n_mels in (80, 128), f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(mel_filterbank_np(n_mels)).to(device)


def log_mel_spectrogram(audio: Union[np.ndarray, torch.Tensor], n_mels: int = 80, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
    """[..., n_samples] -> [..., n_mels, n_frames]; the max used for the 8 dB floor is over the WHOLE tensor."""
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.asarray(audio))
    if device is not None:
        audio = audio.to(device)
    if padding > 0:
        audio = F.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT).to(audio.device)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    mel_spec = mel_filters(audio.device, n_mels) @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec
