"""Audio ingest for the B200 path: RIFF/WAVE decode on the host, everything per-sample on the
device.  Replaces ``stable_whisper.audio.utils.load_audio`` (audio/utils.py:63-125: an ffmpeg subprocess that down-mixes,
resamples to 16 kHz and emits s16le) for PCM/WAV input; compressed containers still need ffmpeg and are not handled here.

    bytes / path -> parse_wav (header walk, no copy of the sample block) -> H2D of the raw interleaved PCM ->
    stb_resample_mono (format conversion + down-mix + polyphase FIR L/M + optional s16 re-quantisation) -> fp32 mono 16 kHz

The filter is defined HERE (not "whatever ffmpeg does"): Kaiser-windowed sinc, ``ZEROS`` zero crossings per side at the
lower of the two rates, roll-off ``ROLLOFF`` of that Nyquist, beta ``BETA``, every polyphase branch normalised to DC gain 1;
``oracle/audio_io.py`` restates it in float64 numpy and is cross-checked against ``scipy.signal.resample_poly`` with the same
taps.  Bit-parity with ffmpeg's swresample is not claimed (ffmpeg is not in this image to compare with).
"""
import math
import os
import struct
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib as L

SAMPLE_RATE = 16000
ZEROS, ROLLOFF, BETA = 24, 0.94, 10.0


def parse_wav(data: bytes) -> Tuple[int, int, int, memoryview]:
    """RIFF/WAVE bytes -> (sample_rate, channels, sample_format, interleaved sample bytes).
    sample_format: 0 = s16, 1 = s32, 2 = f32 (the codes of stb_resample_mono); 8-bit and 24-bit PCM are widened on the
    host to s16 / s32.  WAVE_FORMAT_EXTENSIBLE is resolved through its sub-format."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file (compressed containers need ffmpeg; this loader reads PCM WAV only)")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = memoryview(data)[pos + 8: pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, rate, _, align, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == 0xFFFE and size >= 26:
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            payload = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError("WAV file without fmt/data chunk")
    tag, ch, rate, bits = fmt
    if ch < 1 or ch > 8:
        raise ValueError(f"{ch} channels not supported")
    if tag == 1 and bits == 16:
        return rate, ch, 0, payload
    if tag == 1 and bits == 32:
        return rate, ch, 1, payload
    if tag == 3 and bits == 32:
        return rate, ch, 2, payload
    if tag == 1 and bits == 8:                                   # unsigned 8-bit -> s16
        a = (np.frombuffer(payload, np.uint8).astype(np.int16) - 128) << 8
        return rate, ch, 0, memoryview(a.tobytes())
    if tag == 1 and bits == 24:                                  # packed 24-bit -> s32
        raw = np.frombuffer(payload, np.uint8)[: len(payload) // 3 * 3].reshape(-1, 3).astype(np.int32)
        a = (raw[:, 0] << 8) | (raw[:, 1] << 16) | (raw[:, 2] << 24)
        return rate, ch, 1, memoryview(a.astype(np.int32).tobytes())
    if tag == 3 and bits == 64:
        return rate, ch, 2, memoryview(np.frombuffer(payload, np.float64).astype(np.float32).tobytes())
    raise ValueError(f"unsupported WAV encoding (format tag {tag}, {bits} bits)")


def resample_ratio(in_rate: int, out_rate: int = SAMPLE_RATE) -> Tuple[int, int]:
    g = math.gcd(int(in_rate), int(out_rate))
    return out_rate // g, in_rate // g                          # L (up), M (down)


def polyphase_table(L_up: int, M_down: int, zeros: int = ZEROS, rolloff: float = ROLLOFF, beta: float = BETA) -> np.ndarray:
    """[L][taps] fp32: tab[p][j] = h((j - taps//2) - p / L) in INPUT samples, h = windowed sinc with cutoff
    rolloff * min(1, L/M) / 2 cycles per input sample; each branch scaled to unit DC gain."""
    if L_up == M_down:                                           # same rate: no filter at all (ffmpeg inserts no resampler either)
        return np.ones((1, 1), dtype=np.float32)
    scale = min(1.0, L_up / M_down)
    fc = 0.5 * rolloff * scale                                   # cycles / input sample
    half_width = zeros / scale                                   # support in input samples (per side)
    half = int(math.ceil(half_width)) + 1
    taps = 2 * half + 1
    j = np.arange(taps, dtype=np.float64)[None, :] - half
    p = np.arange(L_up, dtype=np.float64)[:, None] / L_up
    t = j - p                                                    # input-sample offset of tap j from the output instant
    w = np.where(np.abs(t) <= half_width, np.i0(beta * np.sqrt(np.clip(1.0 - (t / half_width) ** 2, 0.0, None))) / np.i0(beta), 0.0)
    h = 2.0 * fc * np.sinc(2.0 * fc * t) * w
    h /= h.sum(axis=1, keepdims=True)
    return h.astype(np.float32)


def resample_to_mono(pcm: torch.Tensor, sample_format: int, channels: int, in_rate: int, out_rate: int = SAMPLE_RATE,
                     quantize_s16: bool = False) -> torch.Tensor:
    """``pcm``: flat CUDA tensor of the interleaved samples (int16 / int32 / float32) -> fp32 [n_out] on the same device."""
    if not pcm.is_cuda:
        raise RuntimeError("stable_ts_b200.audio_io: the PCM block must be on a CUDA device (there is no CPU path)")
    n_in = pcm.numel() // channels
    L_up, M_down = resample_ratio(in_rate, out_rate)
    n_out = -(-n_in * L_up // M_down)
    out = torch.empty(n_out, dtype=torch.float32, device=pcm.device)
    tab = torch.from_numpy(polyphase_table(L_up, M_down)).to(pcm.device)
    with torch.cuda.device(pcm.device):
        L.check(L.lib().stb_resample_mono(L.ptr(pcm), int(sample_format), int(channels), n_in, L_up, M_down, L.ptr(tab),
                                          tab.shape[1], L.ptr(out), n_out, int(quantize_s16), L.stream_ptr()))
    return out


def load_audio(file: Union[str, bytes, os.PathLike], sr: int = SAMPLE_RATE, device: Optional[Union[str, torch.device]] = None,
               quantize_s16: bool = True) -> torch.Tensor:
    """path or bytes of a PCM WAV -> fp32 mono waveform at ``sr`` on ``device`` (default cuda).  ``quantize_s16`` keeps the
    value grid of the reference's loader (its ffmpeg pipe emits s16le, audio/utils.py:104-123)."""
    data = file if isinstance(file, (bytes, bytearray)) else open(os.fspath(file), "rb").read()
    rate, ch, fmt, payload = parse_wav(bytes(data))
    dt = {0: np.int16, 1: np.int32, 2: np.float32}[fmt]
    n = len(payload) // (np.dtype(dt).itemsize * ch) * ch
    host = torch.from_numpy(np.frombuffer(payload, dt, count=n).copy())
    dev = torch.device(device or "cuda")
    if dev.type != "cuda":
        raise RuntimeError("stable_ts_b200.audio_io.load_audio decodes on a CUDA device")
    pcm = host.pin_memory().to(dev, non_blocking=True)
    return resample_to_mono(pcm, fmt, ch, rate, sr, quantize_s16=quantize_s16)
