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


class OracleStepEngine:
    """CPU stand-in for ``stable_ts_b200.decode.StepEngine`` (same attributes and methods), so that the HOST logic of
    ``decode_windows`` / ``decode_with_fallback`` / ``transcribe`` -- right-aligned ragged prompts, per-sequence caps, best_of
    grouping, fallback subsets, prompt carry-over -- runs in the build container and can be compared with the unmodified
    reference.  ``feed`` = the oracle decoder on the sequence's own token history; ``sample`` = the oracle's ApplyTimestampRules
    + the masks the host built + the update rule of stb_sample (argmax, or inverse CDF of the given uniforms)."""

    def __init__(self, model, B, table_rows, reuse_buffers=False, seq_off=None, cache_rows=None):
        self.m, self.B, self.rows = model, B, table_rows
        V = model.dims.n_vocab
        self.ldv = (V + 7) // 8 * 8
        self.off = [0] * B if seq_off is None else [int(v) for v in seq_off]
        self.pos = torch.zeros(1, dtype=torch.int32)
        self.logits = torch.zeros(B, self.ldv)
        self.seq = torch.zeros(B, 6, dtype=torch.int32)
        self.next = torch.zeros(B, dtype=torch.int32)
        self.tok_table = torch.zeros(table_rows, B, dtype=torch.int32)
        self.arg_table = torch.zeros(table_rows, B, dtype=torch.int32)
        self.forced = self.uniform = self.cap = self.graph = None
        self.temperature = 0.0
        self.graph_nodes = 0
        self.hist = [[] for _ in range(B)]
        self.begin = None
        self.sum_lp = [0.0] * B

    def reset(self):
        pass

    def step(self, ckv, *sample_args):
        self.sample(*sample_args)
        self.feed(self.next, ckv)

    @torch.no_grad()
    def feed(self, tokens, ckv):
        om, V, p = self.m.om, self.m.dims.n_vocab, int(self.pos)
        for b in range(self.B):
            if p < self.off[b]:
                continue                                   # idle step of a shorter sequence
            self.hist[b].append(int(tokens[b]))
            if len(self.hist[b]) <= self.m.dims.n_text_ctx:
                self.logits[b, :V] = om.decoder(torch.tensor([self.hist[b]]), ckv["f32"][b:b + 1])[0, -1]
        self.pos += 1

    def sample(self, tk, suppress, first_mask, ts_mask, max_initial_ts, apply_ts_rules):
        from oracle.whisper_ref.decoding import ApplyTimestampRules
        V = self.m.dims.n_vocab
        if self.begin is None:
            self.begin = [len(h) for h in self.hist]
        for b in range(self.B):
            n = len(self.hist[b]) - self.begin[b]
            lg = self.logits[b:b + 1, :V].clone()
            toks = torch.tensor([self.hist[b]])
            if n == 0 and first_mask is not None:
                lg[:, first_mask.bool()] = -np.inf
            if suppress is not None:
                lg[:, suppress.bool()] = -np.inf
            if apply_ts_rules:
                ApplyTimestampRules(tk, self.begin[b], max_initial_ts if max_initial_ts >= 0 else None).apply(lg, toks)
            if ts_mask is not None:
                row = ts_mask if ts_mask.ndim == 1 else ts_mask[b]
                lg[:, tk.timestamp_begin:][:, row.bool()] = -np.inf
            lg.nan_to_num_(-np.inf)
            arg = int(lg[0].argmax())
            pick = arg
            if self.temperature > 0 and n < self.rows:
                c = torch.softmax(lg[0].double() / self.temperature, -1).cumsum(-1)
                hit = (c > float(self.uniform[n, b])).nonzero()
                pick = int(hit[0]) if len(hit) else arg
            done = (n >= 1 and self.hist[b][-1] == tk.eot) or (self.cap is not None and n >= int(self.cap[b]))
            nxt = pick
            if n < self.rows:
                self.arg_table[n, b] = arg
                if self.forced is not None:
                    nxt = int(self.forced[n, b])
            if not done:
                self.sum_lp[b] += float(torch.log_softmax(lg[0].float(), -1)[pick])
            else:
                nxt = int(tk.eot)
            if n < self.rows:
                self.tok_table[n, b] = nxt
            self.next[b] = nxt
            self.seq[b, 4] = int(nxt == tk.eot)
        self.seq[:, 5] = torch.tensor(self.sum_lp, dtype=torch.float32).view(torch.int32)


def _qk_postprocess_new(self, qk_all, S, F, R=None, topk=20, w_colnorm=1.0, w_rownorm=1.0, w_coverage=0.0, qk_scale=1.0,
                        medfilt_width=7):
    """The "new" aligner's matrix (stable_whisper/timing.py:115-163) from the all-head QK block, via the oracle."""
    B, LH, M, _ = qk_all.shape
    Lh, H = self.dims.n_text_layer, self.dims.n_text_head
    out = []
    for b in range(B):
        qks = [qk_all[b, l * H:(l + 1) * H, :, :1500][None] for l in range(Lh)]
        out.append(SP.attention_matrix_new(qks, S, F * 320, medfilt_width, qk_scale, topk=topk, w_colnorm=w_colnorm,
                                           w_rownorm=w_rownorm, w_coverage=w_coverage))
    return torch.stack(out)


def _scale_add(self, y, x, a, b):
    y.mul_(b).add_(x * a) if b != 0.0 else y.copy_(x * a)
    return y


OracleBackedModel.qk_postprocess_new = _qk_postprocess_new
OracleBackedModel.scale_add = _scale_add


def _qk_postprocess_dynamic(self, qk_all, S, F, R=None, count=6, prev_jumps=None, reuse_softmax=False, qk_scale=1.0,
                            medfilt_width=7):
    """Per-token dynamic heads (stable_whisper/timing.py:85-103) from the all-head QK block, via the oracle -> head mean."""
    B, LH, M, _ = qk_all.shape
    Lh, H = self.dims.n_text_layer, self.dims.n_text_head
    R = M - 1 - S if R is None else R
    out = []
    for b in range(B):
        qks = [qk_all[b, l * H:(l + 1) * H, :S + R + 1, :1500][None] for l in range(Lh)]
        pj = None if prev_jumps is None else prev_jumps[b].cpu().numpy()
        out.append(SP.attention_weights_dynamic(qks, S, F * 320, count, pj, medfilt_width, qk_scale).mean(dim=0))
    return torch.stack(out)


OracleBackedModel.qk_postprocess_dynamic = _qk_postprocess_dynamic
