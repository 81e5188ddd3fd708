"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, importable only in the build
container) on top of ``oracle.whisper_ref`` registered as module ``whisper``.  TEST INFRASTRUCTURE.

    python oracle/make_golden.py          # rewrites tests/golden/

Inputs are regenerated from seeds at test time (oracle.stable_path.synth_audio / synth_token_script /
whisper_ref.model.build_model are deterministic); outputs stored here are what the reference computed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import oracle.whisper_ref as W  # noqa: E402

W.install_as_whisper()
from oracle import stable_path as SP  # noqa: E402
from oracle.whisper_ref.model import ModelDimensions  # noqa: E402

from stable_whisper.alignment import get_whisper_alignment_func, get_whisper_refinement_func  # noqa: E402
from stable_whisper.decode import decode_stable  # noqa: E402
from stable_whisper.non_whisper.alignment import WordToken  # noqa: E402
from stable_whisper import timing as ref_timing  # noqa: E402
from whisper.decoding import DecodingOptions  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (dims, model seed, multilingual language)
CASES = {
    "mini_en": (ModelDimensions(80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2), 11),
    "mini_ml": (ModelDimensions(128, 1500, 128, 2, 2, 51866, 448, 128, 2, 3), 12),
}


class _Opts:
    class align:
        extra_models = None
        dynamic_heads = None
        aligner = "legacy"


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, (dims, seed) in CASES.items():
        model = W.build_model(dims, seed=seed)
        tk = W.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en",
                                       task="transcribe")
        n_samples, n_tok = 300000, 40
        audio = SP.synth_audio(n_samples, seed=1234)
        script = SP.synth_token_script(n_tok, tk.eot, seed=4321)
        wts = SP.words_from_script(script, seed=7)
        words = [tk.decode(w) for w in wts]
        # --- align closure (alignment.py:405-429) with the reference's timing.py underneath; capture intermediates
        captured = {}
        orig = ref_timing._compute_jump_indices

        def spy(model, cache, **kw):
            orig(model, cache, **kw)
            captured["jumps"] = np.asarray(cache["jump_indices"]).copy()
            captured["token_probs"] = np.asarray(cache["text_token_probs"], dtype=np.float64)
            w = ref_timing._compute_atten_weights(model, cache=dict(cache), **{k: v for k, v in kw.items()
                                                                              if k not in ("extra_models", "new")})
            captured["matrix"] = w.mean(dim=0).numpy().copy()
        ref_timing._compute_jump_indices = spy
        try:
            out = get_whisper_alignment_func(model, tk, None, _Opts)(audio, [WordToken(w, t) for w, t in zip(words, wts)])
        finally:
            ref_timing._compute_jump_indices = orig
        # --- refine closure (alignment.py:649-672)
        a2 = torch.stack([audio, SP.synth_audio(n_samples, seed=99)])
        probs3 = get_whisper_refinement_func(model, tk, None)(a2, script)
        p, rank = SP.prob_and_rank(probs3, script)
        # --- decode (decode.py:70-110)
        mel = W.pad_or_trim(W.log_mel_spectrogram(audio, dims.n_mels, padding=480000 - n_samples), 3000)
        mask = torch.zeros(1501, dtype=torch.bool)
        mask[100:400] = True
        res, _ = decode_stable(model, mel, DecodingOptions(language="en", fp16=False, sample_len=24), ts_token_mask=mask)
        # --- full driver, first window (original_whisper.py:492-710): decode -> segment slicing -> word timestamps
        import copy
        import json
        import stable_whisper.whisper_word_level.original_whisper as ow
        first = {}
        orig_awt = ow.add_word_timestamps_stable

        def spy_awt(**kw):
            orig_awt(**kw)
            if "segments" not in first:
                first["segments"] = copy.deepcopy(kw["segments"])
                first["num_samples"] = int(kw["num_samples"])
        ow.add_word_timestamps_stable = spy_awt
        try:
            ow.transcribe_stable(model, audio, language="en", temperature=0.0, condition_on_previous_text=False,
                                 word_timestamps=True, vad=False, suppress_silence=False, suppress_ts_tokens=False,
                                 regroup=False, verbose=None, fp16=False, ignore_compatibility=True, sample_len=40)
        finally:
            ow.add_word_timestamps_stable = orig_awt
        tr = [dict(start=float(sg["start"]), end=float(sg["end"]), tokens=[int(t) for t in sg["tokens"]],
                   words=[dict(word=w["word"], start=float(w["start"]), end=float(w["end"]),
                               probability=float(w["probability"]), tokens=[int(t) for t in w["tokens"]]) for w in sg["words"]])
              for sg in first["segments"]]
        with open(os.path.join(GOLD, f"{name}_transcribe.json"), "w") as f:
            json.dump(dict(num_samples=first["num_samples"], segments=tr), f)
        np.savez_compressed(
            os.path.join(GOLD, f"{name}.npz"),
            dims=np.array([getattr(dims, f) for f in dims.__dataclass_fields__], dtype=np.int64),
            model_seed=seed, n_samples=n_samples, script=np.array(script), word_lens=np.array([len(w) for w in wts]),
            word_start=np.array([w["start"] for w in out]), word_end=np.array([w["end"] for w in out]),
            word_prob=np.array([w["probability"] for w in out]),
            jumps=captured["jumps"], token_probs=captured["token_probs"], matrix=captured["matrix"].astype(np.float32),
            refine_p=p.numpy(), refine_rank=rank.numpy(),
            decode_tokens=np.array(res.tokens), decode_avg_logprob=res.avg_logprob,
            decode_no_speech=res.no_speech_prob, mel_checksum=float(mel.double().sum()),
        )
        print(name, "words", len(out), "jumps", captured["jumps"][:6], "decode", res.tokens[:6])
    # --- DTW-only goldens (whisper.timing.dtw semantics as executed through the reference's import)
    rng = np.random.default_rng(2024)
    mats, paths = [], []
    for (R, F) in [(7, 31), (41, 333), (101, 937)]:
        x = rng.standard_normal((R, F)).astype(np.float32)
        if R == 41:
            x = np.round(x * 4) / 4          # exact ties
        from stable_whisper.whisper_compatibility import dtw as ref_dtw
        ti, tj = ref_dtw(torch.from_numpy(-x))
        mats.append(x)
        paths.append(np.stack([ti, tj]).astype(np.int32))
    np.savez_compressed(os.path.join(GOLD, "dtw_cases.npz"), **{f"x{i}": m for i, m in enumerate(mats)},
                        **{f"p{i}": p for i, p in enumerate(paths)})
    print("wrote", os.listdir(GOLD))


if __name__ == "__main__":
    main()
