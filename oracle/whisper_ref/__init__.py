"""CPU oracle: module-shaped restatement of PyPI ``openai-whisper`` 20250625 (TEST INFRASTRUCTURE).

``install_as_whisper()`` registers this package as ``whisper`` (+ ``whisper.audio/.model/.timing/.tokenizer/
.decoding``) in ``sys.modules`` so the UNMODIFIED reference package (stable_whisper/whisper_compatibility.py:58-76
is its only import point) runs on top of it in this container.
"""
import sys

import torch

from . import audio, decoding, model, timing, tokenizer
from .audio import log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult, decode, detect_language
from .model import MODEL_DIMS, ModelDimensions, Whisper, build_model

__version__ = "20250625"


def available_models():
    return list(MODEL_DIMS.keys())


def load_model(name: str, device=None, download_root=None, in_memory: bool = False, seed: int = 0):
    """Random-init model at the named checkpoint's shapes, or a real ``.pt`` checkpoint if ``name`` is a path."""
    import os
    if os.path.isfile(name):
        ckpt = torch.load(name, map_location="cpu", weights_only=False)
        m = Whisper(ModelDimensions(**ckpt["dims"]))
        m.load_state_dict(ckpt["model_state_dict"])
        m = m.eval()
    else:
        m = build_model(name, seed=seed)
    return m.to(device or "cpu")


def transcribe(*a, **k):            # the dependency's own driver is not on the restated path
    raise NotImplementedError("whisper.transcribe is outside the oracle's scope")


def install_as_whisper():
    me = sys.modules[__name__]
    sys.modules.setdefault("whisper", me)
    for sub in ("audio", "decoding", "model", "timing", "tokenizer"):
        sys.modules.setdefault(f"whisper.{sub}", getattr(me, sub))
    return me
