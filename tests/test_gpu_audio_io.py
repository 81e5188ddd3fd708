"""GPU parity of the audio-ingest kernel: stb_resample_mono against the float64 oracle.
fp32 accumulation over <= ~140 taps: |diff| <= 2e-6 before quantisation; on the s16 grid the values are identical except
where the exact result sits within that distance of a rounding boundary (then 1 LSB)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sig(n, rate, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / rate
    return sum(np.sin(2 * np.pi * f * t + p) for f, p in zip(rng.uniform(80, 7000, 5), rng.uniform(0, 6.28, 5))) / 6


@pytest.mark.parametrize("rate,channels,dtype", [(44100, 2, np.int16), (48000, 1, np.int16), (8000, 1, np.float32),
                                                 (22050, 2, np.int32), (16000, 1, np.int16), (96000, 6, np.float32)])
def test_resample_matches_oracle(rate, channels, dtype):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import audio_io as OA
    from stable_ts_b200.audio_io import load_audio
    n = 3 * rate + 17
    x = np.stack([_sig(n, rate, 20 + c) for c in range(channels)], axis=1).reshape(-1)
    pcm = {np.int16: lambda v: np.round(v * 20000).astype(np.int16), np.int32: lambda v: np.round(v * 1.3e9).astype(np.int32),
           np.float32: lambda v: (0.7 * v).astype(np.float32)}[dtype](x)
    wav = OA.make_wav(pcm, rate, channels)
    got = load_audio(wav, quantize_s16=False).cpu().numpy()
    ref = OA.resample_to_mono(pcm, channels, rate)
    assert got.shape == ref.shape
    err = np.abs(got - ref).max()
    gq = load_audio(wav, quantize_s16=True).cpu().numpy()
    rq = OA.resample_to_mono(pcm, channels, rate, quantize_s16=True)
    lsb = np.abs(gq - rq) * 32768
    print(f"{rate} Hz x{channels} {np.dtype(dtype).name}: {len(ref)} samples, max |diff| {err:.2e}, s16 mismatches {int((lsb > 0).sum())}")
    assert err < 2e-6
    assert lsb.max() <= 1.0 and (lsb > 0).mean() < 1e-3
    if rate == 16000 and channels == 1:                       # identity ratio: the centre tap is exactly 1
        np.testing.assert_array_equal(got, pcm.astype(np.float32) / 32768)


def test_ingest_feeds_the_hot_path():
    """WAV bytes -> device waveform -> log-mel, with no host round trip of the samples."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import oracle.whisper_ref as W
    from oracle import audio_io as OA
    from stable_ts_b200.audio_io import load_audio
    from stable_ts_b200.model import from_oracle
    rate = 44100
    pcm = np.round(np.stack([_sig(5 * rate, rate, 1), _sig(5 * rate, rate, 2)], 1).reshape(-1) * 15000).astype(np.int16)
    wave = load_audio(OA.make_wav(pcm, rate, 2))
    assert wave.is_cuda and abs(wave.numel() - 5 * 16000) <= 1
    gm = from_oracle(W.build_model("tiny.en", seed=1))
    mel = gm.log_mel(wave[None])
    # the waveform handed to the hot path is the oracle's up to 1 LSB of the s16 grid on a few samples ...
    ref_wave = OA.resample_to_mono(pcm, 2, rate, quantize_s16=True)
    lsb = np.abs(wave.cpu().numpy() - ref_wave) * 32768
    assert lsb.max() <= 1.0 and (lsb > 0).mean() < 1e-3
    # ... and the log-mel of THAT device waveform equals the oracle's log-mel of the same samples (a 1-LSB difference alone moves
    # bins near the -8 dB floor by ~1e-2, so the two stages are checked separately)
    ref = W.pad_or_trim(W.log_mel_spectrogram(wave.cpu(), 80, padding=480000 - wave.numel()), 3000)
    assert (mel[0].cpu() - ref).abs().max() < 5e-4
