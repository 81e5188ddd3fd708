"""Model-object protocol of ``whisper.model.Whisper`` over the B200 kernels (boundary "B2").

The reference's ``timing.py`` / ``alignment.py`` never call a kernel: they call ``model.encoder(mel)``,
``model.decoder(tokens, xa, kv_cache=)``, ``model(mel, tokens)`` and read the cross-attention ``qk`` of every decoder layer
through forward hooks registered on ``model.decoder.blocks[i].cross_attn`` (stable_whisper/timing.py:51-56,
alignment.py:927-938; the hooked module must return a tuple ending in ``qk``).  The classes below give ``B200Whisper``
exactly that surface, so the UNMODIFIED reference functions run over the sm_100a path:

    encoder(mel [B, n_mels, 3000])           -> xa fp32 [B, 1500, d]          (stb_encoder_forward)
    decoder(tokens [B|1, M], xa)             -> logits fp32 [B, M, V]         (stb_cross_kv + stb_decoder_forward); when any
                                                cross_attn module carries a forward hook, the scaled pre-softmax scores of
                                                ALL heads are captured and each ``blocks[l].cross_attn`` is called with its
                                                layer's ``qk`` [B, H, M, 1500] so the hooks fire as they do in whisper
    decoder(tokens, xa, kv_cache=dict)       -> incremental decoding (stb_decode_step per fed token); the per-sequence K/V
                                                state lives in ``kv_cache`` (whisper's dict protocol: ``clear()`` resets)
    install_kv_cache_hooks()                 -> ({}, [])                      (alignment.py:985)
    detect_language(mel)                     -> (language token ids, [probabilities])

These are host-side adapters: ``nn.Module`` is used only because the reference registers hooks through its API.
"""
from typing import List, Optional

import torch
from torch import nn

from . import _lib as L

_STATE = "__stb_state__"


class CrossAttentionTap(nn.Module):
    """Stand-in for ``ResidualAttentionBlock.cross_attn``: forward returns ``(None, qk)`` -- whisper's
    ``MultiHeadAttention.forward`` returns ``(out, qk)`` and the reference's hooks read ``outs[-1]``."""

    def forward(self, qk: torch.Tensor):
        return None, qk


class _KVModule(nn.Module):
    """Placeholder for the ``key`` / ``value`` Linear modules that whisper's PyTorchInference lists as cache keys."""


class _AttnStub(nn.Module):
    def __init__(self):
        super().__init__()
        self.key, self.value = _KVModule(), _KVModule()


class BlockShim(nn.Module):
    def __init__(self):
        super().__init__()
        self.attn = _AttnStub()
        self.cross_attn = CrossAttentionTap()
        self.cross_attn.key, self.cross_attn.value = _KVModule(), _KVModule()


class EncoderShim(nn.Module):
    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)

    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        m = self._owner
        if mel.ndim == 2:
            mel = mel[None]
        enc = m.encode(mel.to(m.device, torch.float32))
        m._remember_encoding(enc)
        return enc["f32"]


class DecoderShim(nn.Module):
    def __init__(self, owner, n_layer: int):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        self.blocks = nn.ModuleList([BlockShim() for _ in range(n_layer)])

    def _hooked(self) -> bool:
        return any(len(b.cross_attn._forward_hooks) for b in self.blocks)

    def forward(self, x: torch.Tensor, xa: torch.Tensor, kv_cache: Optional[dict] = None) -> torch.Tensor:
        m = self._owner
        enc = m._encoding_of(xa)
        B = enc["B"]
        tokens = x.to(m.device, torch.int32)
        if tokens.ndim == 1:
            tokens = tokens[None]
        if tokens.shape[0] == 1 and B > 1:                       # whisper broadcasts a single token row over the audio batch
            tokens = tokens.expand(B, -1)
        if tokens.shape[0] != B:
            raise ValueError(f"decoder: {tokens.shape[0]} token rows for {B} audio windows")
        if kv_cache is not None:
            return self._incremental(tokens.contiguous(), enc, kv_cache)
        ckv = m.cross_kv(enc)
        hooked = self._hooked()
        logits, qk = m.decode_forced(tokens.contiguous(), ckv, heads="all" if hooked else None)
        if hooked:
            H = m.dims.n_text_head
            for l, blk in enumerate(self.blocks):
                blk.cross_attn(qk[:, l * H:(l + 1) * H, :, : L.N_AUDIO_CTX])
        return logits

    def _incremental(self, tokens: torch.Tensor, enc: dict, kv_cache: dict) -> torch.Tensor:
        from .decode import StepEngine
        m = self._owner
        st = kv_cache.get(_STATE)
        if st is None or st["enc_id"] != id(enc):
            eng = StepEngine(m, enc["B"], 1)
            eng.reset()
            st = dict(eng=eng, ckv=m.cross_kv(enc, decode=True), enc_id=id(enc))
            kv_cache[_STATE] = st
        eng, V = st["eng"], m.dims.n_vocab
        out = []
        with torch.cuda.device(m.device):
            for t in range(tokens.shape[1]):
                eng.feed(tokens[:, t].contiguous(), st["ckv"])
                out.append(eng.logits[:, :V].clone())
        return torch.stack(out, dim=1)


class WhisperProtocol:
    """Mixin for ``B200Whisper``: the attributes / methods of the whisper model object that the reference touches."""

    def _init_protocol(self):
        self.encoder = EncoderShim(self)
        self.decoder = DecoderShim(self, self.dims.n_text_layer)
        self._encodings: List[dict] = []

    # xa tensors handed out by ``encoder`` map back to the split planes the decoder kernels consume
    def _remember_encoding(self, enc: dict):
        self._encodings.append(enc)
        del self._encodings[:-4]

    def _encoding_of(self, xa: torch.Tensor) -> dict:
        for enc in reversed(self._encodings):
            f = enc["f32"]
            if f.data_ptr() == xa.data_ptr() and f.shape == xa.shape:
                return enc
        # a tensor that did not come from ``encoder`` (e.g. repeated / indexed by the caller): split it again
        xa = xa.to(self.device, torch.float32).contiguous()
        if xa.ndim == 2:
            xa = xa[None]
        B, T, d = xa.shape
        hi = torch.empty(B * T, d, dtype=torch.float16, device=self.device)
        lo = torch.empty_like(hi) if self._want_lo else None
        with torch.cuda.device(self.device):
            L.check(self._lib.stb_split_f16(L.ptr(xa), B * T, d, d, L.ptr(hi), L.ptr(lo), d, L.stream_ptr()))
        enc = {"f32": xa, "hi": hi, "lo": lo, "B": B}
        self._remember_encoding(enc)
        return enc

    def __call__(self, mel: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
        return self.decoder(tokens, self.encoder(mel))

    forward = __call__

    def embed_audio(self, mel: torch.Tensor) -> torch.Tensor:
        return self.encoder(mel)

    def logits(self, tokens: torch.Tensor, audio_features: torch.Tensor) -> torch.Tensor:
        return self.decoder(tokens, audio_features)

    def install_kv_cache_hooks(self, cache: Optional[dict] = None):
        return ({} if cache is None else dict(cache)), []

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def detect_language(self, mel: torch.Tensor, tokenizer=None):
        """whisper.decoding.detect_language: one decoder position after <|startoftranscript|>, every non-language logit
        masked, -> (language token ids [B], list of {code: probability})."""
        from .tokenizer import LANGUAGE_CODES, get_tokenizer
        if not self.is_multilingual:
            raise ValueError("This model doesn't have language tokens so it can't perform lang id")
        tk = tokenizer or get_tokenizer(self, synthetic=getattr(self, "random_init", False))
        single = mel.ndim == 2
        xa = self.encoder(mel)
        B = xa.shape[0]
        sot = torch.full((B, 1), int(tk.sot), dtype=torch.int32)
        logits = self.decoder(sot, xa)[:, 0]
        codes = LANGUAGE_CODES[: self.num_languages]
        ids = [tk.sot + 1 + i for i in range(len(codes))]
        mask = torch.ones(logits.shape[-1], dtype=torch.bool, device=logits.device)
        mask[ids] = False
        logits = logits.masked_fill(mask, float("-inf"))
        lang_tokens = logits.argmax(dim=-1)
        probs = logits.softmax(dim=-1).cpu()
        out = [{c: probs[b, j].item() for j, c in zip(ids, codes)} for b in range(B)]
        if single:
            return lang_tokens[0], out[0]
        return lang_tokens, out

    def detect_language_of(self, wave: torch.Tensor) -> str:
        """fp32 16 kHz samples (<= 30 s) -> most probable language code (original_whisper.py:318-341)."""
        wave = wave.detach().float().flatten()[:480000]
        mel = self.log_mel(wave.to(self.device)[None])
        _, probs = self.detect_language(mel)
        return max(probs[0], key=probs[0].get)
