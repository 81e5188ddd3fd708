"""Build-container only: silence-detection fixtures written by the UNMODIFIED reference
(/root/reference/stable_whisper/stabilization) -> tests/golden/silence_cases.npz.

    python oracle/make_golden_silence.py

Each case: the audio generator arguments (the audio itself is regenerated from the seed) and what the reference returned:
loudness, the wav2mask result, and NonSpeechPredictor.predict_with_nonvad's timings / padded mask / is_silent.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"

CASES = [  # (n_samples, seed, floor, scale)
    (480000, 11, 0.0, 1.0), (480000, 12, 1e-4, 1.0), (250001, 13, 3e-4, 1.0), (160000, 14, 0.0, 0.2),
    (480000, 15, 0.0, 1e-6), (100000, 16, 1e-3, 3.0), (3000, 17, 0.0, 1.0), (479999, 18, 2e-4, 1.0), (48000, 19, 0.0, 1.0),
]


def case_audio(n, seed, floor, scale):
    from oracle import stable_path as SP
    if seed == 19:                                         # no gaps at all: wav2mask returns None ("no silence")
        return SP.synth_audio(n, seed=seed) * scale
    return SP.synth_gapped_audio(n, seed=seed, floor=floor) * scale


def main():
    import oracle.whisper_ref as W
    W.install_as_whisper()
    sys.path.insert(0, REFERENCE)
    from stable_whisper.stabilization import NonSpeechPredictor
    from stable_whisper.stabilization.nonvad import audio2loudness, wav2mask
    from stable_whisper.whisper_compatibility import pad_or_trim
    out = {"cases": np.array(CASES, dtype=np.float64)}
    for i, (n, seed, floor, scale) in enumerate(CASES):
        audio = case_audio(int(n), int(seed), floor, scale)
        loud = audio2loudness(audio)
        mask = wav2mask(audio, sr=16000)
        pred = NonSpeechPredictor(vad=False, mask_pad_func=pad_or_trim, get_mask=True, min_word_dur=0.1, sampling_rate=16000,
                                  verbose=None, store_timings=True).predict(audio, offset=12.5)
        out[f"loud_{i}"] = loud.numpy() if loud is not None else np.zeros(0, np.float32)
        out[f"has_mask_{i}"] = np.array(mask is not None)
        out[f"mask_{i}"] = mask.numpy() if mask is not None else np.zeros(0, bool)
        out[f"has_timings_{i}"] = np.array(pred["timings"] is not None)
        out[f"timings_{i}"] = pred["timings"] if pred["timings"] is not None else np.zeros((2, 0))
        out[f"pmask_{i}"] = pred["mask"].numpy() if pred["mask"] is not None else np.zeros(0, bool)
        out[f"silent_{i}"] = np.array(bool(pred["is_silent"]))
        print(i, n, "loud", None if loud is None else tuple(loud.shape), "mask", None if mask is None else int(mask.sum()),
              "timings", None if pred["timings"] is None else pred["timings"].shape, "is_silent", bool(pred["is_silent"]))
    path = os.path.join(ROOT, "tests", "golden", "silence_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
