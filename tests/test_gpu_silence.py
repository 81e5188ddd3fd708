"""GPU parity of the silence-detection kernel against the CPU oracle and the fixtures written by
the unmodified reference: loudness and masks must be BIT-identical (fp32 arithmetic restated operation by operation).

First hardware run: round 2 (gpurun_out/r2_first/silence_tests.log, 3 passed)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "silence_cases.npz")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_loudness_and_sound_mask_bit_exact_vs_oracle():
    _need_gpu()
    from oracle import silence as SIL
    from oracle.make_golden_silence import CASES, case_audio
    from stable_ts_b200.silence import sound_masks
    for (n, seed, floor, scale) in CASES:
        audio = case_audio(int(n), int(seed), floor, scale)
        m, loud = sound_masks(audio.cuda()[None], want_loudness=True)
        ref_loud = SIL.audio2loudness(audio.numpy())
        assert np.array_equal(loud[0], ref_loud), f"n={n} seed={seed}: loudness differs by {np.abs(loud[0] - ref_loud).max()}"
        assert np.array_equal(m[0], SIL.loudness_to_raw_mask(ref_loud))


def test_batch_of_windows_and_kernel_sizes():
    _need_gpu()
    from oracle import silence as SIL
    from oracle import stable_path as SP
    from stable_ts_b200.silence import sound_masks, wav2mask_batch
    audios = torch.stack([SP.synth_gapped_audio(480000, seed=200 + i, floor=(0.0, 1e-4, 1e-3)[i % 3]) for i in range(7)])
    for q, k in ((20, 5), (20, 3), (10, 7), (0, 5), (20, 0)):
        m, _ = sound_masks(audios.cuda(), q_levels=q, k_size=k)
        for b in range(len(audios)):
            assert np.array_equal(m[b], SIL.loudness_to_raw_mask(SIL.audio2loudness(audios[b].numpy()), q, k)), (q, k, b)
    got = wav2mask_batch(audios.cuda())
    for b in range(len(audios)):
        ref = SIL.wav2mask(audios[b].numpy())
        assert (got[b] is None) == (ref is None)
        if ref is not None:
            assert np.array_equal(got[b], ref)


def test_predict_matches_reference_fixtures():
    _need_gpu()
    from oracle.make_golden_silence import CASES, case_audio
    from stable_ts_b200.silence import predict_nonvad_batch
    z = np.load(GOLD)
    for i, (n, seed, floor, scale) in enumerate(CASES):
        audio = case_audio(int(n), int(seed), floor, scale)
        pred = predict_nonvad_batch(audio.cuda()[None], offsets=[12.5])[0]
        assert (pred["timings"] is not None) == bool(z[f"has_timings_{i}"])
        if pred["timings"] is not None:
            assert np.array_equal(pred["timings"], z[f"timings_{i}"])
        assert pred["is_silent"] == bool(z[f"silent_{i}"])
        if pred["mask"] is not None:
            assert np.array_equal(pred["mask"].numpy(), z[f"pmask_{i}"])
        else:
            assert z[f"pmask_{i}"].size == 0
