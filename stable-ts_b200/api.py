"""``load_model`` and model methods (mirror of stable_whisper/whisper_word_level/original_whisper.py:931-1009).

The returned object carries ``align`` / ``align_words`` / ``refine`` (and ``transcribe`` once the decode path is
loaded).  When the reference package is importable (a user switching over has it installed), ``align`` and ``refine``
delegate the window/seek control logic to the reference's own model-agnostic ``Aligner`` / ``Refiner`` with the B200
plugin closures; otherwise the built-in batched drivers are used.
"""
import os
from typing import Optional, Union

import torch

from .model import B200Whisper, ModelDimensions

MODEL_DIMS = {
    "tiny.en": (80, 1500, 384, 6, 4, 51864, 448, 384, 6, 4), "tiny": (80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base.en": (80, 1500, 512, 8, 6, 51864, 448, 512, 8, 6), "base": (80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small.en": (80, 1500, 768, 12, 12, 51864, 448, 768, 12, 12), "small": (80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium.en": (80, 1500, 1024, 16, 24, 51864, 448, 1024, 16, 24),
    "medium": (80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v1": (80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v2": (80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v3": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large-v3-turbo": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
    "turbo": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
}


def random_state_dict(dims: ModelDimensions, seed: int = 0):
    """Seeded random weights at the true Whisper shapes (checkpoints cannot be downloaded offline).  Key names equal
    openai-whisper's.  Generated on the CPU in fp32 so that every device/process sees identical values."""
    import math
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True):
        a = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (2 * torch.rand(out_f, in_f, generator=g) - 1) * a
        if bias:
            sd[name + ".bias"] = (2 * torch.rand(out_f, generator=g) - 1) * a

    def ln(name, d):
        sd[name + ".weight"] = 1.0 + 0.1 * (2 * torch.rand(d, generator=g) - 1)
        sd[name + ".bias"] = 0.05 * (2 * torch.rand(d, generator=g) - 1)

    def block(p, d, cross):
        for n in ("attn",) + (("cross_attn",) if cross else ()):
            lin(f"{p}.{n}.query", d, d)
            lin(f"{p}.{n}.key", d, d, bias=False)
            lin(f"{p}.{n}.value", d, d)
            lin(f"{p}.{n}.out", d, d)
            ln(f"{p}.{n}_ln", d)
        lin(f"{p}.mlp.0", 4 * d, d)
        lin(f"{p}.mlp.2", d, 4 * d)
        ln(f"{p}.mlp_ln", d)

    da, dt = dims.n_audio_state, dims.n_text_state
    a1, a2 = 1.0 / math.sqrt(3 * dims.n_mels), 1.0 / math.sqrt(3 * da)
    sd["encoder.conv1.weight"] = (2 * torch.rand(da, dims.n_mels, 3, generator=g) - 1) * a1
    sd["encoder.conv1.bias"] = (2 * torch.rand(da, generator=g) - 1) * a1
    sd["encoder.conv2.weight"] = (2 * torch.rand(da, da, 3, generator=g) - 1) * a2
    sd["encoder.conv2.bias"] = (2 * torch.rand(da, generator=g) - 1) * a2
    half = da // 2
    inv = torch.exp(-(math.log(10000.0) / (half - 1)) * torch.arange(half))
    st = torch.arange(dims.n_audio_ctx)[:, None] * inv[None, :]
    sd["encoder.positional_embedding"] = torch.cat([torch.sin(st), torch.cos(st)], dim=1)
    for l in range(dims.n_audio_layer):
        block(f"encoder.blocks.{l}", da, False)
    ln("encoder.ln_post", da)
    sd["decoder.token_embedding.weight"] = torch.randn(dims.n_vocab, dt, generator=g) * 0.05
    sd["decoder.positional_embedding"] = torch.randn(dims.n_text_ctx, dt, generator=g) * 0.05
    for l in range(dims.n_text_layer):
        block(f"decoder.blocks.{l}", dt, True)
    ln("decoder.ln", dt)
    return sd


def load_model(name: str = "base", device: Optional[Union[str, torch.device]] = None, download_root: str = None,
               in_memory: bool = False, cpu_preload: bool = True, dq: bool = False, engine: Optional[str] = None, *,
               precision: str = "fp16x3", seed: int = 0) -> B200Whisper:
    """Same signature as the reference's ``load_model`` (+ ``precision`` / ``seed``).

    ``name``: an official model name or a path to an openai-whisper ``.pt`` checkpoint
    (``{"dims": ..., "model_state_dict": ...}``).  For a model NAME, the checkpoint is looked up in ``download_root``
    (default ``~/.cache/whisper``) -- there is no network here, so when it is absent the model is built with seeded
    random weights at the named shapes and ``model.random_init`` is set.
    """
    if dq:
        raise ValueError("dq (CPU dynamic quantisation) does not apply to the B200 path")
    device = device or "cuda"
    path = name if os.path.isfile(name) else os.path.join(download_root or os.path.expanduser("~/.cache/whisper"),
                                                          f"{name}.pt")
    if os.path.isfile(path):
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        dims = ModelDimensions(**ckpt["dims"])
        model = B200Whisper(dims, ckpt["model_state_dict"], device=device, precision=precision)
        model.random_init = False
    else:
        if name not in MODEL_DIMS:
            raise RuntimeError(f"Model {name} not found; available models = {list(MODEL_DIMS)}")
        dims = ModelDimensions(*MODEL_DIMS[name])
        model = B200Whisper(dims, random_state_dict(dims, seed), device=device, precision=precision)
        model.random_init = True
    model.name = name
    return model
