"""Batched transcription driver for the B200 path (mirror of the per-window body of ``transcribe_stable``,
stable_whisper/whisper_word_level/original_whisper.py:492-710).

The reference walks ONE audio sequentially (its seek depends on the decoded timestamps).  Here independent 30 s windows
(different audios, or static shards of one audio with fixed clip boundaries and no prompt carry-over -- the sharded
setting of SURVEY.md section 8e) are processed B at a time:

    log-mel -> encoder -> KV-cached greedy decode -> segment slicing at timestamp tokens -> batched word timestamps

Out of scope here (reference control plane): silence suppression / VAD, temperature fallback,
prompt conditioning, language detection, regrouping.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from .decode import DecodingOptions, decode_windows
from .model import B200Whisper
from .timing import add_word_timestamps_batch

SAMPLE_RATE = 16000
N_SAMPLES = 480000
N_SAMPLES_PER_TOKEN = 320
TIME_PRECISION = 0.02


def slice_segments(tokens: Sequence[int], tokenizer, time_offset: float, segment_duration: float, result) -> (List[dict], int):
    """Split one window's sampled tokens into segments at consecutive timestamp tokens
    (original_whisper.py:550-602).  -> (segments, end_timestamp_pos)."""
    tb = tokenizer.timestamp_begin
    toks = list(tokens)
    is_ts = [t >= tb for t in toks]
    single_ending = is_ts[-2:] == [False, True]
    consecutive = [i + 1 for i in range(len(toks) - 1) if is_ts[i] and is_ts[i + 1]]
    seek = round(time_offset * SAMPLE_RATE)

    def seg(start, end, ts):
        text_tokens = [t for t in ts if t < tokenizer.eot]
        return dict(seek=round(seek / SAMPLE_RATE, 3), start=start, end=end, tokens=list(ts), text=tokenizer.decode(text_tokens),
                    temperature=result.temperature, avg_logprob=result.avg_logprob,
                    compression_ratio=result.compression_ratio, no_speech_prob=result.no_speech_prob)

    segments, end_pos = [], 0
    if consecutive:
        slices = consecutive + ([len(toks)] if single_ending else [])
        last = 0
        for cur in slices:
            ts = toks[last:cur]
            s_pos, end_pos = ts[0] - tb, ts[-1] - tb
            segments.append(seg(round(time_offset + s_pos * TIME_PRECISION, 3),
                                round(time_offset + min(end_pos * TIME_PRECISION, segment_duration), 3), ts))
            last = cur
    else:
        duration = segment_duration
        stamps = [t for t in toks if t >= tb]
        if stamps and stamps[-1] != tb:
            end_pos = stamps[-1] - tb
            duration = min(end_pos * TIME_PRECISION, segment_duration)
        segments.append(seg(round(time_offset, 3), round(time_offset + duration, 3), toks))
    return segments, end_pos


@torch.no_grad()
def transcribe_windows(model: B200Whisper, tokenizer, audios: Sequence[torch.Tensor], *, time_offsets: Optional[Sequence[float]] = None,
                       word_timestamps: bool = True, options: Optional[DecodingOptions] = None,
                       ts_token_mask: Optional[torch.Tensor] = None, forced_tokens: Optional[torch.Tensor] = None,
                       gap_padding: Optional[str] = " ...", min_word_dur: float = 0.1, punctuations: str = "\"'“¿([{-\"'.。,，!！?？:：”)]}、",
                       enc: Optional[dict] = None, n_samples: Optional[Sequence[int]] = None, use_graph: bool = True):
    """B independent <=30 s windows -> list (per window) of segment dicts with ``words``.
    ``enc`` (+ ``n_samples``) may be passed instead of ``audios`` when the encoder output is already on the device."""
    if enc is None:
        if torch.is_tensor(audios) and audios.ndim == 2 and audios.shape[1] == N_SAMPLES and audios.dtype == torch.float32:
            batch = audios                                   # already a [B, 480000] batch (ideally pinned): no host copy
            if n_samples is None:
                n_samples = [N_SAMPLES] * batch.shape[0]
        else:
            B = len(audios)
            batch = torch.zeros(B, N_SAMPLES, dtype=torch.float32)
            n_samples = []
            for i, a in enumerate(audios):
                a = a.detach().float().flatten()[:N_SAMPLES]
                batch[i, : a.numel()] = a
                n_samples.append(int(a.numel()))
        if not batch.is_pinned():
            batch = batch.pin_memory()
        mel = model.log_mel(batch.to(model.device, non_blocking=True))
        enc = model.encode(mel)
    B = enc["B"]
    n_samples = list(n_samples) if n_samples is not None else [N_SAMPLES] * B
    offs = list(time_offsets) if time_offsets is not None else [0.0] * B
    if options is None:                      # transcribe_stable defaults max_initial_timestamp to None (original_whisper.py:262-263)
        options = DecodingOptions(max_initial_timestamp=None)
    # the cross K/V block and the KV cache live in model-owned buffers: they are consumed inside this call (decode loop, then
    # the alignment pass below) and are by far the largest allocations of a step
    results, extras = decode_windows(model, tokenizer, enc, options, ts_token_mask=ts_token_mask,
                                     forced_tokens=forced_tokens, use_graph=use_graph, reuse_buffers=True)
    windows = []
    for b in range(B):
        dur = n_samples[b] / SAMPLE_RATE
        toks = results[b].tokens if forced_tokens is None else extras["step_tokens"][:, b].tolist()
        segs, end_pos = slice_segments(toks, tokenizer, offs[b], dur, results[b]) if len(toks) else ([], 0)
        # prune punctuation-only and zero-length segments (original_whisper.py:604-627, word_timestamps branch)
        # (`in` on a str is a substring test, so empty-text segments are dropped too, exactly as the reference does)
        segs = [s for s in segs if s["text"].strip() not in punctuations]
        segs = [s for s in segs if not (word_timestamps and s["start"] == s["end"])]
        for s in segs:
            s["seek"] = offs[b]
        num = min(round(end_pos * N_SAMPLES_PER_TOKEN), n_samples[b]) if end_pos > 0 else n_samples[b]
        windows.append(dict(segments=segs, num_samples=num))
    if word_timestamps:
        add_word_timestamps_batch(windows, model, tokenizer, enc=enc, ckv=extras["ckv"], gap_padding=gap_padding,
                                  min_word_dur=min_word_dur)
    return [w["segments"] for w in windows], dict(decode=results, steps=extras["steps"], step_argmax=extras["step_argmax"],
                                                  step_tokens=extras["step_tokens"])


def transcribe(model: B200Whisper, tokenizer, audio: torch.Tensor, *, batch_windows: int = 16, **kw) -> dict:
    """Static 30 s sharding of one long audio (clip boundaries fixed at multiples of 30 s, no prompt carry-over).
    -> dict(text, segments, language) in the shape of WhisperResult.to_dict (result.py:1398-1406)."""
    audio = audio.detach().float().flatten()
    chunks = [audio[i:i + N_SAMPLES] for i in range(0, audio.numel(), N_SAMPLES)]
    segments = []
    for i in range(0, len(chunks), batch_windows):
        part = chunks[i:i + batch_windows]
        segs, _ = transcribe_windows(model, tokenizer, part, time_offsets=[(i + k) * 30.0 for k in range(len(part))], **kw)
        for ws in segs:
            segments.extend(ws)
    for k, s in enumerate(segments):
        s["id"] = k
    return dict(text="".join(s["text"] for s in segments), segments=segments,
                language=getattr(tokenizer, "language", None) or "en")
