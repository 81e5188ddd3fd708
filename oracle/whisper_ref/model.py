"""CPU oracle: restatement of ``whisper.model`` (openai-whisper 20250625).  TEST INFRASTRUCTURE.

Reference call sites that reach this code: stable_whisper/timing.py:50-61 (encoder + teacher-forced
decoder with forward hooks on ``decoder.blocks[i].cross_attn`` under ``disable_sdpa``),
stable_whisper/decode.py:27-40 (encoder cache + KV-cached decoder steps),
stable_whisper/alignment.py:660-667 (``model(mel, tokens)``), :985 (``install_kv_cache_hooks``).

State-dict key names equal openai-whisper's, so real ``*.pt`` checkpoints (``{"dims":..., "model_state_dict":...}``)
load unchanged.  Weights are otherwise seeded random-init (no checkpoints offline).
"""
import base64
import gzip
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn


@dataclass
class ModelDimensions:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


# name -> dims of the released checkpoints (shapes only; weights are not available offline)
MODEL_DIMS = {
    "tiny.en": ModelDimensions(80, 1500, 384, 6, 4, 51864, 448, 384, 6, 4),
    "tiny": ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base.en": ModelDimensions(80, 1500, 512, 8, 6, 51864, 448, 512, 8, 6),
    "base": ModelDimensions(80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small.en": ModelDimensions(80, 1500, 768, 12, 12, 51864, 448, 768, 12, 12),
    "small": ModelDimensions(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium.en": ModelDimensions(80, 1500, 1024, 16, 24, 51864, 448, 1024, 16, 24),
    "medium": ModelDimensions(80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v1": ModelDimensions(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v2": ModelDimensions(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v3": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large-v3-turbo": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
    "turbo": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
}


class LayerNorm(nn.LayerNorm):
    def forward(self, x: Tensor) -> Tensor:          # statistics always in fp32
        return super().forward(x.float()).type(x.dtype)


class Linear(nn.Linear):
    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype))


class Conv1d(nn.Conv1d):
    def _conv_forward(self, x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
        return super()._conv_forward(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> Tensor:
    assert channels % 2 == 0
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


@contextmanager
def disable_sdpa():
    prev = MultiHeadAttention.use_sdpa
    try:
        MultiHeadAttention.use_sdpa = False
        yield
    finally:
        MultiHeadAttention.use_sdpa = prev


class MultiHeadAttention(nn.Module):
    use_sdpa = True

    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = Linear(n_state, n_state)
        self.key = Linear(n_state, n_state, bias=False)
        self.value = Linear(n_state, n_state)
        self.out = Linear(n_state, n_state)

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                kv_cache: Optional[dict] = None):
        q = self.query(x)
        if kv_cache is None or xa is None or self.key not in kv_cache:
            # hooks, if installed, prepend the cached keys/values (self-attention)
            k = self.key(x if xa is None else xa)
            v = self.value(x if xa is None else xa)
        else:
            # cross-attention: computed once per window and reused
            k = kv_cache[self.key]
            v = kv_cache[self.value]
        wv, qk = self.qkv_attention(q, k, v, mask)
        return self.out(wv), qk

    def qkv_attention(self, q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None
                      ) -> Tuple[Tensor, Optional[Tensor]]:
        n_batch, n_ctx, n_state = q.shape
        scale = (n_state // self.n_head) ** -0.25
        q = q.view(*q.shape[:2], self.n_head, -1).permute(0, 2, 1, 3)
        k = k.view(*k.shape[:2], self.n_head, -1).permute(0, 2, 1, 3)
        v = v.view(*v.shape[:2], self.n_head, -1).permute(0, 2, 1, 3)
        if MultiHeadAttention.use_sdpa:
            a = F.scaled_dot_product_attention(q, k, v, is_causal=mask is not None and n_ctx > 1)
            out = a.permute(0, 2, 1, 3).flatten(start_dim=2)
            qk = None
        else:
            qk = (q * scale) @ (k * scale).transpose(-1, -2)
            if mask is not None:
                qk = qk + mask[:n_ctx, :n_ctx]
            qk = qk.float()
            w = F.softmax(qk, dim=-1).to(q.dtype)
            out = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
            qk = qk.detach()      # fp32, scaled, PRE-softmax: what stable_whisper/timing.py:53 captures
        return out, qk


class ResidualAttentionBlock(nn.Module):
    def __init__(self, n_state: int, n_head: int, cross_attention: bool = False):
        super().__init__()
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        self.cross_attn = MultiHeadAttention(n_state, n_head) if cross_attention else None
        self.cross_attn_ln = LayerNorm(n_state) if cross_attention else None
        n_mlp = n_state * 4
        self.mlp = nn.Sequential(Linear(n_state, n_mlp), nn.GELU(), Linear(n_mlp, n_state))
        self.mlp_ln = LayerNorm(n_state)

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                kv_cache: Optional[dict] = None):
        x = x + self.attn(self.attn_ln(x), mask=mask, kv_cache=kv_cache)[0]
        if self.cross_attn is not None:
            x = x + self.cross_attn(self.cross_attn_ln(x), xa, kv_cache=kv_cache)[0]
        x = x + self.mlp(self.mlp_ln(x))
        return x


class AudioEncoder(nn.Module):
    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks: Iterable[ResidualAttentionBlock] = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head) for _ in range(n_layer)])
        self.ln_post = LayerNorm(n_state)

    def forward(self, x: Tensor):
        """x: [B, n_mels, 3000] -> [B, 1500, n_state]"""
        x = F.gelu(self.conv1(x))
        x = F.gelu(self.conv2(x))
        x = x.permute(0, 2, 1)
        assert x.shape[1:] == self.positional_embedding.shape, "incorrect audio shape"
        x = (x + self.positional_embedding).to(x.dtype)
        for block in self.blocks:
            x = block(x)
        return self.ln_post(x)


class TextDecoder(nn.Module):
    def __init__(self, n_vocab: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.token_embedding = nn.Embedding(n_vocab, n_state)
        self.positional_embedding = nn.Parameter(torch.empty(n_ctx, n_state))
        self.blocks: Iterable[ResidualAttentionBlock] = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head, cross_attention=True) for _ in range(n_layer)])
        self.ln = LayerNorm(n_state)
        mask = torch.empty(n_ctx, n_ctx).fill_(-np.inf).triu_(1)
        self.register_buffer("mask", mask, persistent=False)

    def forward(self, x: Tensor, xa: Tensor, kv_cache: Optional[dict] = None):
        """x: int [B, <=n_ctx] tokens; xa: [B, 1500, n_state] -> fp32 logits [B, n, n_vocab]"""
        offset = next(iter(kv_cache.values())).shape[1] if kv_cache else 0
        x = self.token_embedding(x) + self.positional_embedding[offset: offset + x.shape[-1]]
        x = x.to(xa.dtype)
        for block in self.blocks:
            x = block(x, xa, mask=self.mask, kv_cache=kv_cache)
        x = self.ln(x)
        return (x @ torch.transpose(self.token_embedding.weight.to(x.dtype), 0, 1)).float()


class Whisper(nn.Module):
    def __init__(self, dims: ModelDimensions):
        super().__init__()
        self.dims = dims
        self.encoder = AudioEncoder(dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head,
                                    dims.n_audio_layer)
        self.decoder = TextDecoder(dims.n_vocab, dims.n_text_ctx, dims.n_text_state, dims.n_text_head,
                                   dims.n_text_layer)
        # default: every head of the last half of the decoder layers
        all_heads = torch.zeros(dims.n_text_layer, dims.n_text_head, dtype=torch.bool)
        all_heads[dims.n_text_layer // 2:] = True
        self.register_buffer("alignment_heads", all_heads.to_sparse(), persistent=False)

    def set_alignment_heads(self, dump):
        """``dump``: the base85+gzip bool table of the released models, or a bool [L, H] array/tensor."""
        if isinstance(dump, (bytes, str)):
            arr = np.frombuffer(gzip.decompress(base64.b85decode(dump)), dtype=bool).copy()
            mask = torch.from_numpy(arr).reshape(self.dims.n_text_layer, self.dims.n_text_head)
        else:
            mask = torch.as_tensor(np.asarray(dump), dtype=torch.bool).reshape(
                self.dims.n_text_layer, self.dims.n_text_head)
        self.register_buffer("alignment_heads", mask.to_sparse(), persistent=False)

    def embed_audio(self, mel: Tensor):
        return self.encoder(mel)

    def logits(self, tokens: Tensor, audio_features: Tensor):
        return self.decoder(tokens, audio_features)

    def forward(self, mel: Tensor, tokens: Tensor) -> Tensor:
        return self.decoder(tokens, self.encoder(mel))

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def is_multilingual(self):
        return self.dims.n_vocab >= 51865

    @property
    def num_languages(self):
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)

    def install_kv_cache_hooks(self, cache: Optional[dict] = None):
        """Forward hooks on every decoder key/value Linear that concatenate outputs along time."""
        cache = {**cache} if cache is not None else {}
        hooks = []

        def save_to_cache(module, _, output):
            if module not in cache or output.shape[1] > self.dims.n_text_ctx:
                cache[module] = output                       # first token or cross-attention
            else:
                cache[module] = torch.cat([cache[module], output], dim=1).detach()
            return cache[module]

        def install_hooks(layer: nn.Module):
            if isinstance(layer, MultiHeadAttention):
                hooks.append(layer.key.register_forward_hook(save_to_cache))
                hooks.append(layer.value.register_forward_hook(save_to_cache))

        self.decoder.apply(install_hooks)
        return cache, hooks

    # bound lazily to avoid an import cycle (decoding imports model types)
    def detect_language(self, mel, tokenizer=None):
        from .decoding import detect_language as _dl
        return _dl(self, mel, tokenizer)

    def decode(self, mel, options=None, **kw):
        from .decoding import decode as _d, DecodingOptions
        return _d(self, mel, options or DecodingOptions(), **kw)


def init_random_(model: Whisper, seed: int = 0, peaky: float = 1.0) -> Whisper:
    """Deterministic random init at the true Whisper shapes (no checkpoints offline).

    Linear/conv: U(-a, a), a = 1/sqrt(fan_in) (PyTorch's default scale); embeddings N(0, 0.05);
    LayerNorm gamma ~ 1 +- 0.1, beta ~ +-0.05.  ``peaky`` scales the cross-attention query/key weights so the
    cross-attention rows become sharply peaked (used by tests that need a DTW landscape with unambiguous minima).
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("token_embedding.weight") or name.endswith("positional_embedding"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "ln" in name.split(".")[-2]:
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * (2 * torch.rand(p.shape, generator=g) - 1))
                else:
                    p.copy_(0.05 * (2 * torch.rand(p.shape, generator=g) - 1))
            else:
                fan_in = p[0].numel() if p.ndim > 1 else p.numel()
                a = 1.0 / np.sqrt(fan_in)
                p.copy_((2 * torch.rand(p.shape, generator=g) - 1) * a)
            if peaky != 1.0 and (".cross_attn.query.weight" in name or ".cross_attn.key.weight" in name):
                p.mul_(peaky)
    return model


def build_model(name_or_dims, seed: int = 0, peaky: float = 1.0) -> Whisper:
    dims = MODEL_DIMS[name_or_dims] if isinstance(name_or_dims, str) else name_or_dims
    model = Whisper(dims)
    init_random_(model, seed, peaky)
    return model.eval()
