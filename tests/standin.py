"""TEST INFRASTRUCTURE: an oracle-backed stand-in for ``B200Whisper``'s method surface.

``B200Whisper`` needs a CUDA device; the build container has none.  The host side of the boundary (plugin closures, bound
model methods, batching, bookkeeping) is nevertheless plain Python over ~12 model methods, so those methods are provided
here on the CPU by the oracle (``oracle.whisper_ref``), with the same argument meaning and output layouts (padded QK rows of
1504 columns, views of padded logits, int32 jumps...).  Tests then drive the UNMODIFIED reference ``Aligner`` / ``Refiner``
over ``stable_ts_b200``'s closures and demand results identical to the reference's own closures over the same oracle model:
that pins the host logic of the boundary; the kernels behind the real methods are pinned separately by the ``-m gpu`` tests.
Never imported by the product."""
from typing import Dict, Optional

import numpy as np
import torch

import oracle.whisper_ref as W
from oracle import stable_path as SP
from oracle.whisper_ref.model import disable_sdpa
from stable_ts_b200.shim import WhisperProtocol

KPAD = 1504


class OracleBackedModel(WhisperProtocol):
    def __init__(self, oracle_model):
        self.om = oracle_model
        self.dims = oracle_model.dims
        self.device = torch.device("cpu")
        self.random_init = True                       # selects the synthetic vocabulary, as for random-weight GPU models
        self.missing_alignment_heads = False
        self.alignment_head_pairs = SP.head_pairs_of(oracle_model)
        self._want_lo = False
        self.calls: Dict[str, int] = {}
        # whisper model-object protocol (shim.py on the real model): here the oracle's own modules
        self.encoder, self.decoder = oracle_model.encoder, oracle_model.decoder
        self.install_kv_cache_hooks = oracle_model.install_kv_cache_hooks

    is_multilingual = property(lambda self: self.om.is_multilingual)
    num_languages = property(lambda self: self.om.num_languages)
    alignment_heads = property(lambda self: self.om.alignment_heads)

    def _count(self, k):
        self.calls[k] = self.calls.get(k, 0) + 1

    def log_mel(self, audio, padded_samples=None, batch_global_max=False):
        self._count("log_mel")
        if audio.ndim == 1:
            audio = audio[None]
        audio = audio.float()[:, :480000]
        n = audio.shape[1]
        pad = (480000 if padded_samples is None else int(padded_samples)) - n
        if batch_global_max:
            mel = W.log_mel_spectrogram(audio, self.dims.n_mels, padding=pad)
        else:
            mel = torch.stack([W.log_mel_spectrogram(a, self.dims.n_mels, padding=pad) for a in audio])
        return W.pad_or_trim(mel, 3000)

    @torch.no_grad()
    def encode(self, mel):
        self._count("encode")
        if mel.ndim == 2:
            mel = mel[None]
        xa = self.om.encoder(mel.float())
        return {"f32": xa, "hi": None, "lo": None, "B": xa.shape[0]}

    def cross_kv(self, enc, decode=False, reuse=False):
        return enc

    @torch.no_grad()
    def decode_forced(self, tokens, ckv, want_logits=True, heads=None, reuse=False):
        self._count("decode_forced")
        xa = ckv["f32"]
        tokens = tokens.long()
        B, M = tokens.shape
        qks = [None] * self.dims.n_text_layer
        hooks = [blk.cross_attn.register_forward_hook(lambda _m, _i, out, i=i: qks.__setitem__(i, out[-1]))
                 for i, blk in enumerate(self.om.decoder.blocks)]
        try:
            with disable_sdpa():
                logits = self.om.decoder(tokens, xa)
        finally:
            for h in hooks:
                h.remove()
        qk = None
        if heads is not None:
            pairs = ([(l, h) for l in range(self.dims.n_text_layer) for h in range(self.dims.n_text_head)]
                     if isinstance(heads, str) else [(int(a), int(b)) for a, b in heads])
            qk = torch.zeros(B, len(pairs), M, KPAD)
            for i, (l, h) in enumerate(pairs):
                qk[:, i, :, :1500] = qks[l][:, h]
        return (logits if want_logits else None), qk

    def token_probs(self, logits_rows, n_classes, targets, want_rank=False):
        p = logits_rows[:, :n_classes].float().softmax(-1)
        t = targets.long()
        prob = p[torch.arange(len(t)), t]
        rank = (p < prob[:, None]).sum(-1).int() if want_rank else None
        return prob, rank

    def softmax_probs(self, logits_rows, n_classes):
        return logits_rows[:, :n_classes].float().softmax(-1)

    def qk_postprocess(self, qk, S, F, R=None, qk_scale=1.0, medfilt_width=7):
        B, A, M, _ = qk.shape
        R = M - 1 - S if R is None else R
        out = []
        for b in range(B):
            w = (qk[b, :, S:S + R, :F] * qk_scale).softmax(dim=-1)
            out.append(SP._znorm_median(w, medfilt_width).mean(dim=0))
        return torch.stack(out)

    def dtw(self, matrix, negate=True, want_path=False):
        assert negate and not want_path
        return torch.from_numpy(np.stack([SP.jumps_from_matrix(m) for m in matrix]).astype(np.int32))
