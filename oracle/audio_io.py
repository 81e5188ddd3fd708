"""CPU oracle of the audio ingest (stable-ts_b200/audio_io.py + csrc/resample.cu).  TEST INFRASTRUCTURE ONLY.

Reference being replaced: stable_whisper/audio/utils.py:96-125 (``ffmpeg -ac 1 -ar 16000 -f s16le`` -> int16 -> /32768).
ffmpeg is absent from this image, so the resampling FILTER is this repo's own definition (Kaiser-windowed sinc, parameters
below must equal audio_io.ZEROS / ROLLOFF / BETA); what the oracle pins is the arithmetic: float64 numpy restatement of
down-mix + polyphase FIR + s16 re-quantisation, cross-checked in tests/test_oracle_audio_io.py against
``scipy.signal.upfirdn`` (an independent polyphase implementation) fed with the same prototype filter.  parity: the value
GRID and pipeline order follow the reference (mono mean, resample, round to s16, /32768); the taps are unpinned by it."""
import math

import numpy as np

ZEROS, ROLLOFF, BETA = 24, 0.94, 10.0


def polyphase_table(L: int, M: int) -> np.ndarray:
    if L == M:                                                   # same rate: identity (the reference's ffmpeg would not resample)
        return np.ones((1, 1), dtype=np.float64)
    scale = min(1.0, L / M)
    fc = 0.5 * ROLLOFF * scale
    half_width = ZEROS / scale
    half = int(math.ceil(half_width)) + 1
    taps = 2 * half + 1
    j = np.arange(taps, dtype=np.float64)[None, :] - half
    p = np.arange(L, dtype=np.float64)[:, None] / L
    t = j - p
    w = np.where(np.abs(t) <= half_width, np.i0(BETA * np.sqrt(np.clip(1.0 - (t / half_width) ** 2, 0.0, None))) / np.i0(BETA), 0.0)
    h = 2.0 * fc * np.sinc(2.0 * fc * t) * w
    return h / h.sum(axis=1, keepdims=True)


def to_float(pcm: np.ndarray) -> np.ndarray:
    if pcm.dtype == np.int16:
        return pcm.astype(np.float64) / 32768.0
    if pcm.dtype == np.int32:
        return pcm.astype(np.float64) / 2147483648.0
    return pcm.astype(np.float64)


def resample_to_mono(pcm: np.ndarray, channels: int, in_rate: int, out_rate: int = 16000, quantize_s16: bool = False) -> np.ndarray:
    """interleaved samples [n_in * channels] -> float32 [ceil(n_in * L / M)]; float32 product/accumulate order is NOT
    imitated (float64 here), so compare the kernel with a tolerance of a few fp32 ulps before quantisation."""
    g = math.gcd(in_rate, out_rate)
    L, M = out_rate // g, in_rate // g
    x = to_float(pcm).reshape(-1, channels).astype(np.float32).astype(np.float64).sum(axis=1) / channels
    tab = polyphase_table(L, M).astype(np.float32).astype(np.float64)
    taps = tab.shape[1]
    half = taps // 2
    n_in = len(x)
    n_out = -(-n_in * L // M)
    xp = np.concatenate([np.zeros(half), x, np.zeros(half + taps)])
    m = np.arange(n_out, dtype=np.int64)
    n0 = (m * M) // L
    ph = (m * M) % L
    idx = n0[:, None] + np.arange(taps)[None, :]
    y = (xp[idx] * tab[ph]).sum(axis=1)
    if quantize_s16:
        y = np.clip(np.rint(y * 32768.0), -32768, 32767) / 32768.0
    return y.astype(np.float32)


def make_wav(samples: np.ndarray, rate: int, channels: int) -> bytes:
    """interleaved int16 / int32 / float32 samples -> RIFF/WAVE bytes (with an odd-sized LIST chunk before `data`)."""
    import struct
    tag, bits = {np.dtype(np.int16): (1, 16), np.dtype(np.int32): (1, 32), np.dtype(np.float32): (3, 32)}[samples.dtype]
    body = samples.tobytes()
    fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    junk = b"LIST" + struct.pack("<I", 5) + b"INFOx" + b"\x00"
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(body)) + body
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks
