"""Host-side mirror of stable_whisper/decode.py for the B200 path: KV-cached greedy decoding of a BATCH of windows.

``decode_stable`` keeps the reference's call shape for one window; ``decode_windows`` is the batched engine:
    encoder output (cached, as DecodingTaskStable._get_audio_features) -> cross K/V ->
    per step: [fused logit filters + greedy pick] -> [decoder step for B sequences]     (decode.py:33-65)
The two kernels-sequences of a step are captured ONCE into a CUDA graph and replayed; every position-dependent value
lives in device memory (position counter, per-sequence sampling state, token/argmax tables indexed by step), so the
host only polls the `done` flags every few steps instead of the reference's per-step ``.all()`` sync.
"""
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .model import B200Whisper

CHUNK_LENGTH = 30


@dataclass
class DecodingOptions:              # whisper.decoding.DecodingOptions (fields that apply to the greedy path)
    task: str = "transcribe"
    language: Optional[str] = None
    temperature: float = 0.0
    sample_len: Optional[int] = None
    suppress_tokens: Optional[str] = "-1"
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    prompt: Optional[List[int]] = None
    fp16: bool = False


@dataclass
class DecodingResult:               # whisper.decoding.DecodingResult
    audio_features: Optional[torch.Tensor]
    language: str
    language_probs: Optional[Dict[str, float]] = None
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    temperature: float = np.nan
    compression_ratio: float = np.nan


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


def _suppress_list(tokenizer, options: DecodingOptions) -> List[int]:
    st = options.suppress_tokens
    if isinstance(st, str):
        st = [int(t) for t in st.split(",")]
    st = list(st or [])
    if -1 in st:
        st = [t for t in st if t >= 0] + list(tokenizer.non_speech_tokens)
    st += [tokenizer.transcribe, tokenizer.translate, tokenizer.sot, tokenizer.sot_prev, tokenizer.sot_lm]
    if tokenizer.no_speech is not None:
        st.append(tokenizer.no_speech)
    return sorted(set(st))


class StepEngine:
    """Device state of a batch of B decoding sequences + the two launch sequences of one step."""

    def __init__(self, model: B200Whisper, B: int, table_rows: int, reuse_buffers: bool = False):
        self.m, self.B, self.rows = model, B, table_rows
        dev, lib, V = model.device, model._lib, model.dims.n_vocab
        self.ldv = (V + 7) // 8 * 8
        if reuse_buffers:      # model-owned: the self-attention K/V cache (rows beyond `pos` are never read) and the workspace
            self.state = model._buf("decode_state", lib.stb_decode_state_bytes(model._h, B))
            self.ws = model._buf("decode_ws", lib.stb_decode_ws_bytes(model._h, B))
            self.ws.zero_()                                                                          # tickets start at 0
        else:
            self.state = torch.zeros(lib.stb_decode_state_bytes(model._h, B), dtype=torch.uint8, device=dev)
            self.ws = torch.zeros(lib.stb_decode_ws_bytes(model._h, B), dtype=torch.uint8, device=dev)   # tickets start at 0
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits = torch.empty(B, self.ldv, dtype=torch.float32, device=dev)
        self.seq = torch.zeros(B, 6, dtype=torch.int32, device=dev)          # stb_seq_state[B]
        self.next = torch.zeros(B, dtype=torch.int32, device=dev)
        self.tok_table = torch.zeros(table_rows, B, dtype=torch.int32, device=dev)
        self.arg_table = torch.zeros(table_rows, B, dtype=torch.int32, device=dev)
        self.forced = None
        self.graph = None
        self.graph_nodes = 0          # kernels in the captured step graph

    def reset(self):
        self.pos.zero_()
        self.seq.zero_()
        self.seq[:, 3] = -1                                                   # last_ts = -1

    def feed(self, tokens: torch.Tensor, ckv: torch.Tensor):
        """decoder step for tokens [B] int32 (device) -> self.logits; pos += 1"""
        m = self.m
        L.check(m._lib.stb_decode_step(m._h, L.ptr(tokens), self.B, L.ptr(self.pos), L.ptr(ckv), L.ptr(self.state),
                                       L.ptr(self.logits), self.ldv, L.ptr(self.ws), self.ws.numel(), L.stream_ptr()))

    def sample(self, tk, suppress, first_mask, ts_mask, max_initial_ts, apply_ts_rules):
        m = self.m
        L.check(m._lib.stb_sample_greedy(L.ptr(self.logits), self.ldv, self.B, m.dims.n_vocab, int(tk.eot),
                                         int(tk.timestamp_begin), int(tk.no_timestamps), L.ptr(suppress), L.ptr(first_mask),
                                         L.ptr(ts_mask), 0 if ts_mask is None or ts_mask.ndim == 1 else int(ts_mask.stride(0)),
                                         int(max_initial_ts), int(apply_ts_rules), L.ptr(self.forced),
                                         L.ptr(self.seq), L.ptr(self.next), L.ptr(self.tok_table), L.ptr(self.arg_table),
                                         self.rows, L.stream_ptr()))


def decode_windows(model: B200Whisper, *args, **kwargs):
    """See ``_decode_windows``; runs with the model's device current (launches go to that device's current stream)."""
    with torch.cuda.device(model.device):
        return _decode_windows(model, *args, **kwargs)


@torch.no_grad()
def _decode_windows(model: B200Whisper, tokenizer, enc: dict, options: Optional[DecodingOptions] = None, *,
                   ts_token_mask: Optional[torch.Tensor] = None, forced_tokens: Optional[torch.Tensor] = None,
                   use_graph: bool = True, poll_every: int = 16, return_step_logits: bool = False, ckv=None,
                   reuse_buffers: bool = False):
    """Greedy (temperature 0) decode of the B windows whose encoder output is ``enc`` (from ``model.encode``).

    ts_token_mask: silent-timestamp suppression (decode.py:14-16): bool [1501] shared by the batch, bool [B, 1501] with one
                   row per window (what the reference computes, original_whisper.py:504-511), a list of B optional [1501]
                   masks (None = nothing suppressed for that window), or None.
    forced_tokens: int [steps, B] -- the token appended at each step instead of the argmax (fixed-length scripts for
                   random-weight benchmarks); the argmax of every step is still returned.
    reuse_buffers: keep the cross K/V block, the KV cache and the step workspace in model-owned buffers (the returned
                   ``extras["ckv"]`` is then only valid until the next such call).
    -> (list of DecodingResult, extras dict(step_argmax [steps,B], step_tokens, sum_logprob, ckv))
    """
    options = options or DecodingOptions()
    if options.temperature != 0:
        raise NotImplementedError("B200 decode path: only temperature 0 (greedy) is implemented")
    if options.prompt:
        raise NotImplementedError("B200 decode path: prompt conditioning is not implemented (windows are independent)")
    dev, B, V = model.device, enc["B"], model.dims.n_vocab
    n_ctx = model.dims.n_text_ctx
    sample_len = options.sample_len or n_ctx // 2
    init = list(tokenizer.sot_sequence_including_notimestamps if options.without_timestamps else tokenizer.sot_sequence)
    sample_begin = len(init)
    steps = sample_len if forced_tokens is None else min(sample_len, int(forced_tokens.shape[0]))
    eng = StepEngine(model, B, steps, reuse_buffers=reuse_buffers)
    eng.reset()
    if ckv is None:
        ckv = model.cross_kv(enc, decode=True, reuse=reuse_buffers)
    # filter tables (SuppressTokens, SuppressBlank)
    sup = torch.zeros(V, dtype=torch.uint8)
    sup[_suppress_list(tokenizer, options)] = 1 if options.suppress_tokens else 0
    first = torch.zeros(V, dtype=torch.uint8)
    if options.suppress_blank:
        first[tokenizer.encode(" ") + [tokenizer.eot]] = 1
    sup, first = sup.to(dev), first.to(dev)
    tsm = None
    if ts_token_mask is not None:
        if isinstance(ts_token_mask, (list, tuple)):
            if len(ts_token_mask) != B:
                raise ValueError(f"ts_token_mask: {len(ts_token_mask)} masks for {B} windows")
            rows = torch.zeros(B, 1501, dtype=torch.uint8)
            for i, mk in enumerate(ts_token_mask):
                if mk is not None:
                    rows[i] = torch.as_tensor(mk).to(torch.uint8)
            ts_token_mask = rows
        if ts_token_mask.ndim == 2 and ts_token_mask.shape[0] != B:
            raise ValueError(f"ts_token_mask: {ts_token_mask.shape[0]} rows for {B} windows")
        if ts_token_mask.shape[-1] != 1501:
            raise ValueError("ts_token_mask must have 1501 entries per window")
        tsm = ts_token_mask.to(torch.uint8).to(dev).contiguous()
    apply_rules = not options.without_timestamps
    max_init = -1
    if apply_rules and options.max_initial_timestamp:
        max_init = round(options.max_initial_timestamp / (CHUNK_LENGTH / model.dims.n_audio_ctx))
    if forced_tokens is not None:
        eng.forced = forced_tokens[:steps].to(dev, torch.int32).contiguous()
    # ---- initial tokens: one step each; the logits at the SOT position give no_speech_prob (decode.py:42-44)
    no_speech = [float("nan")] * B
    sot_index = init.index(tokenizer.sot)
    for i, t in enumerate(init):
        eng.feed(torch.full((B,), int(t), dtype=torch.int32, device=dev), ckv)
        if i == sot_index and tokenizer.no_speech is not None:
            p, _ = model.token_probs(eng.logits, V, torch.full((B,), int(tokenizer.no_speech)))
            no_speech = p.cpu().tolist()
    step_logits = []

    def one_step():
        eng.sample(tokenizer, sup, first, tsm, max_init, apply_rules)
        eng.feed(eng.next, ckv)

    done_steps = 0
    # step 0 runs eagerly (also warms every kernel), the rest replay a captured graph
    if return_step_logits:
        use_graph = False
    while done_steps < steps:
        if return_step_logits:
            eng.sample(tokenizer, sup, first, tsm, max_init, apply_rules)
            step_logits.append(eng.logits[:, :V].clone())
            eng.feed(eng.next, ckv)
        elif use_graph and done_steps >= 1:
            if eng.graph is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                c0 = model._lib.stb_launch_count()
                with torch.cuda.graph(g):
                    one_step()                          # this capture pass does not execute; replay below does
                eng.graph = g
                eng.graph_nodes = int(model._lib.stb_launch_count() - c0)
                model.graph_kernel_launches -= eng.graph_nodes      # counted by the library at capture, but not executed
            eng.graph.replay()
            model.graph_kernel_launches += eng.graph_nodes          # kernels executed by this replay
        else:
            one_step()
        done_steps += 1
        if sample_begin + done_steps >= n_ctx:          # tokens.shape[-1] > n_ctx stop (decode.py:60)
            break
        if forced_tokens is None and done_steps % poll_every == 0:
            if bool((eng.seq[:, 4] != 0).all()):        # every sequence has emitted EOT
                break
    torch.cuda.synchronize()
    toks = eng.tok_table[:done_steps].cpu().numpy()
    args = eng.arg_table[:done_steps].cpu().numpy()
    sum_lp = eng.seq[:, 5].contiguous().view(torch.float32).cpu().numpy()
    results = []
    for b in range(B):
        seq = toks[:, b].tolist()
        if tokenizer.eot in seq:
            seq = seq[: seq.index(tokenizer.eot)]
        text = tokenizer.decode(seq).strip()
        results.append(DecodingResult(audio_features=enc["f32"][b], language=options.language or "en", tokens=seq, text=text,
                                      avg_logprob=float(sum_lp[b]) / (len(seq) + 1), no_speech_prob=no_speech[b],
                                      temperature=0.0, compression_ratio=compression_ratio(text) if text else float("nan")))
    extras = dict(step_argmax=args, step_tokens=toks, sum_logprob=sum_lp, ckv=ckv, step_logits=step_logits,
                  steps=done_steps)
    return results, extras


def decode_stable(model: B200Whisper, mel: torch.Tensor, options: Optional[DecodingOptions] = None,
                  ts_token_mask: Optional[torch.Tensor] = None, audio_features: Optional[dict] = None, **kwargs):
    """Same contract as stable_whisper/decode.py:70-110: -> (DecodingResult | list, audio_features).
    ``audio_features`` is the dict returned by ``model.encode`` (reused across temperature fallbacks)."""
    from .tokenizer import get_tokenizer
    options = options or DecodingOptions()
    for k, v in kwargs.items():
        setattr(options, k, v)
    single = mel.ndim == 2
    if audio_features is None:
        audio_features = model.encode(mel[None] if single else mel)
    tk = get_tokenizer(model, language=options.language or "en", task=options.task,
                       synthetic=getattr(model, "random_init", True))
    res, _ = decode_windows(model, tk, audio_features, options, ts_token_mask=ts_token_mask)
    return (res[0] if single else res), audio_features
