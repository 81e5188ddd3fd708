"""CPU oracle for the reference's OWN orchestration of the hot path.  TEST INFRASTRUCTURE ONLY.

Restates, as plain functions over ``oracle.whisper_ref`` (fp32, CPU):

  window_qks                  stable_whisper/timing.py:41-67   (_compute_qks: encoder + teacher-forced decoder,
                                                                cross-attention QK capture, token probabilities)
  attention_weights_legacy    stable_whisper/timing.py:105-110 (alignment heads -> slice -> softmax -> z-norm -> median)
  attention_weights_dynamic   stable_whisper/timing.py:85-103,108-110
  attention_matrix_new        stable_whisper/timing.py:115-163 (arXiv 2509.09987 head scoring)
  jumps_from_matrix           stable_whisper/timing.py:191-198 (head mean, dtw(-matrix), first frame of every row)
  align_window                stable_whisper/timing.py:202-306 + alignment.py:405-429 (word start/end/probability)
  decode_window               stable_whisper/decode.py:33-65,70-110 (KV-cached greedy loop with timestamp mask)
  refine_token_probs          stable_whisper/alignment.py:649-672

It cannot import /root/reference (absent on the GPU box); tests/test_oracle_vs_reference.py checks it against the
unmodified reference in this container, and oracle/make_golden.py stores fixtures produced by the reference itself.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .whisper_ref import audio as _audio
from .whisper_ref import decoding as _decoding
from .whisper_ref import timing as _timing
from .whisper_ref.model import Whisper, disable_sdpa

N_SAMPLES_PER_TOKEN = _audio.N_SAMPLES_PER_TOKEN
TOKENS_PER_SECOND = _audio.TOKENS_PER_SECOND


def alignment_token_row(tokenizer, text_tokens: Sequence[int]) -> List[int]:
    """[*sot_sequence, no_timestamps, *text, eot]  (timing.py:230-237)."""
    return [*tokenizer.sot_sequence, tokenizer.no_timestamps, *text_tokens, tokenizer.eot]


def n_frames_for(num_samples: int) -> int:
    """Usable cross-attention columns: Python round() (banker's) of num_samples/320 (timing.py:88,106,137)."""
    return round(num_samples / N_SAMPLES_PER_TOKEN)


@torch.no_grad()
def window_qks(model: Whisper, tokenizer, text_tokens: Sequence[int], mel: torch.Tensor,
               audio_features: Optional[torch.Tensor] = None):
    """-> (audio_features [1,1500,d], qks: L x fp32 [1,H,M,1500], logits fp32 [M,V], token_probs list[N])."""
    tokens = torch.tensor(alignment_token_row(tokenizer, text_tokens)).to(mel.device)
    qks: List[Optional[torch.Tensor]] = [None] * model.dims.n_text_layer
    hooks = [blk.cross_attn.register_forward_hook(lambda _m, _i, out, i=i: qks.__setitem__(i, out[-1]))
             for i, blk in enumerate(model.decoder.blocks)]
    try:
        with disable_sdpa():
            if audio_features is None:
                audio_features = model.encoder(mel.unsqueeze(0))
            logits = model.decoder(tokens.unsqueeze(0), audio_features)[0]
    finally:
        for h in hooks:
            h.remove()
    S = len(tokenizer.sot_sequence)
    probs = logits[S:, : tokenizer.eot].softmax(dim=-1)
    token_probs = probs[np.arange(len(text_tokens)), list(text_tokens)].tolist()
    return audio_features, qks, logits, token_probs


def _znorm_median(weights: torch.Tensor, medfilt_width: int) -> torch.Tensor:
    std, mean = torch.std_mean(weights, dim=-2, keepdim=True, unbiased=False)
    return _timing.median_filter((weights - mean) / std, medfilt_width)


def attention_weights_legacy(qks, head_pairs: Sequence[Tuple[int, int]], S: int, num_samples: int,
                             medfilt_width: int = 7, qk_scale: float = 1.0) -> torch.Tensor:
    """-> fp32 [A, N+1, F]."""
    F_ = n_frames_for(num_samples)
    w = torch.cat([qks[l][:, h] for l, h in head_pairs], dim=0)
    w = w[:, S:-1, :F_]
    w = (w * qk_scale).softmax(dim=-1)
    return _znorm_median(w, medfilt_width)


def attention_weights_dynamic(qks, S: int, num_samples: int, count: int = 6, prev_jumps: Optional[np.ndarray] = None,
                              medfilt_width: int = 7, qk_scale: float = 1.0) -> torch.Tensor:
    """Per-token top-``count`` heads by distance-weighted mass (timing.py:87-103). -> [count, N+1, F]."""
    F_ = n_frames_for(num_samples)
    allq = torch.cat([qk[0, :, S:-1, :F_] for qk in qks])
    allq = (allq * qk_scale).softmax(dim=-1)
    if prev_jumps is None:
        peaks = allq.topk(1, dim=-1).indices
    else:
        ji = np.pad(prev_jumps, (0, 1), constant_values=F_)
        pk = ji[:-1] + (ji[1:] - ji[:-1]) * 0.5
        peaks = torch.from_numpy(pk).to(allq.device)[None, :, None]
    dist = (peaks.expand_as(allq) - torch.arange(allq.size(-1), device=allq.device)).abs() / 1500
    scores = (dist * allq).sum(dim=-1)
    heads = [s.topk(count, largest=False).indices for s in scores.T]
    w = torch.stack([allq[h, i] for i, h in enumerate(heads)], dim=1)
    return _znorm_median(w, medfilt_width)


def attention_matrix_new(qks, S: int, num_samples: int, medfilt_width: int = 7, qk_scale: float = 1.0, *,
                         topk: int = 20, w_colnorm: float = 1, w_rownorm: float = 1, w_coverage: float = 0
                         ) -> torch.Tensor:
    """-> fp32 [N+1, F] (already head-averaged)."""
    F_ = n_frames_for(num_samples)
    w = torch.cat(qks)[..., :F_]
    w = _timing.median_filter(w, medfilt_width)
    w = (w * qk_scale).softmax(dim=-1)
    L, H = w.size(0), w.size(1)
    score = torch.zeros(L, H, device=w.device)
    if w_colnorm > 0:
        score += w_colnorm * w.norm(dim=-2).sum(-1)
    if w_rownorm > 0:
        score += w_rownorm * w.norm(dim=-1).sum(-1)
    if w_coverage > 0:
        cov = torch.sum(w, dim=2)
        pen = torch.max(cov, cov.clone().fill_(0.5)).sum(-1) - cov.size(-1) * 0.5
        score -= w_coverage * pen
    top = score.flatten().topk(topk).indices
    m = w[top // H, top % H]
    m = torch.mean(m / m.norm(dim=-2, keepdim=True), 0)
    return m[S:-1]


def jumps_from_matrix(matrix: torch.Tensor) -> np.ndarray:
    """matrix fp32 [R, F] (larger = more attention).  -> int [R]: first frame of each decoder row on the DTW path."""
    text_idx, time_idx = _timing.dtw(-matrix)
    jumps = np.pad(np.diff(text_idx), (1, 0), constant_values=1).astype(bool)
    return time_idx[jumps].clip(min=0)


def head_pairs_of(model) -> List[Tuple[int, int]]:
    idx = model.alignment_heads.indices().T.tolist()
    return [(int(l), int(h)) for l, h in idx]


@torch.no_grad()
def align_window(model: Whisper, tokenizer, word_tokens: List[List[int]], mel: torch.Tensor, num_samples: int, *,
                 words: Optional[List[str]] = None, medfilt_width: int = 7, qk_scale: float = 1.0,
                 dynamic_heads=None, aligner="legacy", audio_features=None, return_intermediates: bool = False):
    """One forced-alignment window, the way the reference's ``compute_timestamps`` closure drives it
    (alignment.py:405-429: no gap padding, identity word split, no punctuation merge, seek 0).

    -> list of dict(word, tokens, start, end, probability) (+ intermediates dict).
    """
    text_tokens = [t for wt in word_tokens for t in wt]
    S = len(tokenizer.sot_sequence)
    audio_features, qks, logits, token_probs = window_qks(model, tokenizer, text_tokens, mel, audio_features)
    if dynamic_heads:
        if dynamic_heads is True:
            count, iters = 6, 1
        elif isinstance(dynamic_heads, int):
            count, iters = dynamic_heads, 1
        else:
            c, i = dynamic_heads.split(",")
            count, iters = int(c), int(i)
    else:
        count, iters = None, 1
    if count is None and getattr(model, "missing_alignment_heads", False):
        count = 6
    jumps = None
    weights = matrix = None
    for _ in range(iters or 1):
        if aligner != "legacy":
            kw = aligner if isinstance(aligner, dict) else {}
            matrix = attention_matrix_new(qks, S, num_samples, medfilt_width, qk_scale, **kw)
        else:
            if count:
                weights = attention_weights_dynamic(qks, S, num_samples, count, jumps, medfilt_width, qk_scale)
            else:
                weights = attention_weights_legacy(qks, head_pairs_of(model), S, num_samples, medfilt_width, qk_scale)
            matrix = weights.mean(dim=0)
        jumps = jumps_from_matrix(matrix)
    # word boundaries in token units; the EOT pseudo-word is dropped by truncation (timing.py:249-251,301-306)
    wb = np.pad(np.cumsum([len(t) for t in word_tokens]), (1, 0))
    jump_times = jumps / TOKENS_PER_SECOND
    starts, ends = jump_times[wb[:-1]], jump_times[wb[1:]]
    out = []
    for k, wt in enumerate(word_tokens):
        if len(wt) == 0:
            continue
        out.append(dict(word=None if words is None else words[k],
                        start=round(0 + float(starts[k]), 3), end=round(0 + float(ends[k]), 3),
                        probability=float(np.mean(token_probs[wb[k]:wb[k + 1]])), tokens=list(wt)))
    if return_intermediates:
        return out, dict(audio_features=audio_features, qks=qks, logits=logits, token_probs=token_probs,
                         weights=weights, matrix=matrix, jumps=jumps)
    return out


def align_audio_window(model, tokenizer, word_tokens, audio: torch.Tensor, **kw):
    """audio fp32 [n<=480000] -> as alignment.py:409-413: pad to 30 s in the sample domain, log-mel, trim to 3000."""
    n = int(audio.shape[-1])
    mel = _audio.log_mel_spectrogram(audio, model.dims.n_mels, padding=max(_audio.N_SAMPLES - n, 0))
    mel = _audio.pad_or_trim(mel, _audio.N_FRAMES)
    return align_window(model, tokenizer, word_tokens, mel, n, **kw)


@torch.no_grad()
def decode_window(model: Whisper, mel: torch.Tensor, *, ts_token_mask: Optional[torch.Tensor] = None,
                  audio_features: Optional[torch.Tensor] = None, forced_tokens: Optional[Sequence[int]] = None,
                  return_step_logits: bool = False, **options):
    """Greedy KV-cached decode of one window (decode.py:33-65): logits -> filters -> silent-timestamp mask ->
    nan_to_num(-inf) -> argmax/logprob.  ``forced_tokens`` (bench/test only) overrides the sampled token at each
    step AFTER the argmax has been recorded, so the step count is fixed for random-weight models.

    -> (DecodingResult, audio_features, extras{step_argmax, step_logits})
    """
    options.setdefault("fp16", False)
    task = _decoding.DecodingTask(model, _decoding.DecodingOptions(**options))
    tk = task.tokenizer
    if mel.ndim == 2:
        mel = mel.unsqueeze(0)
    if audio_features is None:
        audio_features = task._get_audio_features(mel)
    tokens = torch.tensor([task.initial_tokens]).to(audio_features.device)
    sum_logprobs = torch.zeros(1)
    step_argmax, step_logits = [], []
    no_speech = float("nan")
    try:
        for i in range(task.sample_len):
            logits = task.inference.logits(tokens, audio_features)
            if i == 0 and tk.no_speech is not None:
                no_speech = logits[:, task.sot_index].float().softmax(dim=-1)[0, tk.no_speech].item()
            logits = logits[:, -1]
            for f in task.logit_filters:
                f.apply(logits, tokens)
            if ts_token_mask is not None:
                logits[:, tk.timestamp_begin:][:, ts_token_mask] = -np.inf
            logits.nan_to_num_(-np.inf)
            if return_step_logits:
                step_logits.append(logits[0].clone())
            step_argmax.append(int(logits[0].argmax()))
            tokens, completed = task.decoder.update(tokens, logits, sum_logprobs)
            if forced_tokens is not None:
                tokens[0, -1] = int(forced_tokens[i])
                completed = i + 1 >= len(forced_tokens)
            if completed or tokens.shape[-1] > task.n_ctx:
                break
    finally:
        task.inference.cleanup_caching()
    toks = tokens[0, task.sample_begin:].tolist()
    if tk.eot in toks:
        toks = toks[: toks.index(tk.eot)]
    res = _decoding.DecodingResult(audio_features=audio_features[0], language=options.get("language") or "en",
                                   tokens=toks, text=tk.decode(toks).strip(),
                                   avg_logprob=float(sum_logprobs[0]) / (len(toks) + 1), no_speech_prob=no_speech,
                                   temperature=options.get("temperature", 0.0),
                                   compression_ratio=_decoding.compression_ratio(tk.decode(toks).strip()))
    return res, audio_features, dict(step_argmax=step_argmax, step_logits=step_logits,
                                     sum_logprob=float(sum_logprobs[0]))


@torch.no_grad()
def refine_token_probs(model: Whisper, tokenizer, audio2: torch.Tensor, tokens: Sequence[int]) -> torch.Tensor:
    """audio2 fp32 [2, n] -> softmax probs fp32 [2, N, eot] (alignment.py:649-672: batch-global log-mel max,
    frame-domain zero padding, token row broadcast over the 2 audio rows)."""
    row = torch.tensor(alignment_token_row(tokenizer, tokens))
    mel = _audio.pad_or_trim(_audio.log_mel_spectrogram(audio2, model.dims.n_mels), _audio.N_FRAMES)
    logits = model(mel, row.unsqueeze(0))
    S = len(tokenizer.sot_sequence)
    return logits[:, S:S + len(tokens), : tokenizer.eot].softmax(dim=-1)


def prob_and_rank(probs3: torch.Tensor, tokens: Sequence[int]):
    """What refinement.py:305-325 extracts: p[target] and the position of target in the ASCENDING sort of the row."""
    idx = torch.arange(len(tokens))
    t = torch.tensor(list(tokens))
    p = probs3[:, idx, t]
    order = probs3.sort(dim=-1).indices                       # [2, N, V]
    rank = (order == t[None, :, None]).nonzero()[:, -1].reshape(probs3.shape[0], len(tokens))
    return p, rank


def synth_audio(n_samples: int, seed: int = 1234) -> torch.Tensor:
    """Speech-like synthetic audio: AM-modulated sinusoids + noise, peak 0.3."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / _audio.SAMPLE_RATE
    k = int(torch.randint(3, 6, (1,), generator=g))
    x = torch.zeros(n_samples, dtype=torch.float64)
    for _ in range(k):
        f = 100 + 3900 * float(torch.rand(1, generator=g))
        fm = 2 + 6 * float(torch.rand(1, generator=g))
        ph = 2 * np.pi * float(torch.rand(1, generator=g))
        x += torch.sin(2 * np.pi * f * t + ph) * (0.5 + 0.5 * torch.sin(2 * np.pi * fm * t))
    x += 0.01 * torch.randn(n_samples, generator=g, dtype=torch.float64)
    x = 0.3 * x / x.abs().max()
    return x.float()


def synth_token_script(n: int, eot: int, seed: int = 4321, lo: int = 256) -> List[int]:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, eot, (n,), generator=g).tolist()


def words_from_script(tokens: Sequence[int], seed: int = 7) -> List[List[int]]:
    """Group a token script into synthetic 'words' of 1-3 tokens."""
    g = torch.Generator().manual_seed(seed)
    out, i = [], 0
    while i < len(tokens):
        k = int(torch.randint(1, 4, (1,), generator=g))
        out.append(list(tokens[i:i + k]))
        i += k
    return out


def synth_gapped_audio(n_samples: int, seed: int = 77, floor: float = 0.0) -> torch.Tensor:
    """``synth_audio`` with silent / near-silent stretches (random 0.05-2.5 s gaps, ``floor`` = residual noise amplitude in
    the gaps): exercises the silence detector."""
    x = synth_audio(n_samples, seed=seed).double()
    g = torch.Generator().manual_seed(seed + 1000)
    env = torch.ones(n_samples, dtype=torch.float64)
    pos = int(torch.randint(0, 16000, (1,), generator=g))
    while pos < n_samples:
        gap = int((0.05 + 2.45 * float(torch.rand(1, generator=g))) * 16000)
        env[pos:pos + gap] = 0.0
        pos += gap + int((0.2 + 3.0 * float(torch.rand(1, generator=g))) * 16000)
    y = x * env
    if floor > 0:
        y = y + floor * torch.randn(n_samples, generator=g, dtype=torch.float64) * (1 - env)
    return y.float()


# ---------------------------------------------------------------------------------------------------------------------
# One transcribe window end to end (the per-window body of transcribe_stable): decode -> segments at timestamp tokens ->
# gap-padded word timestamps.  Restates stable_whisper/whisper_word_level/original_whisper.py:537-665 and
# stable_whisper/timing.py:309-500; pinned against the unmodified reference in tests/test_oracle_vs_reference.py.
# ---------------------------------------------------------------------------------------------------------------------
PREPEND_PUNCTUATIONS = "\"'“¿([{-"
APPEND_PUNCTUATIONS = "\"'.。,，!！?？:：”)]}、"
TIME_PRECISION = 0.02


def group_tokens_into_words(tokens: Sequence[int], tokenizer) -> Tuple[List[str], List[List[int]]]:
    """Incremental-decode word grouping (timing.py:309-341): a token group closes when its decoded text is a prefix
    of what is left of the full decode; in space-delimited languages a closed group that neither starts with a space
    nor is pure punctuation is glued to the previous word."""
    import string as _string
    lang = getattr(tokenizer, "language_code", tokenizer.language)
    spaced = lang not in {"zh", "ja", "th", "lo", "my"}
    rest = tokenizer.decode_with_timestamps(list(tokens))
    words: List[str] = []
    groups: List[List[int]] = []
    open_group: List[int] = []
    piece = ""
    glue = False
    for tok in tokens:
        open_group.append(tok)
        piece = tokenizer.decode(open_group)
        closed = tok >= tokenizer.eot
        if not closed and rest.startswith(piece):
            closed = True
            if spaced:
                glue = not (piece.startswith(" ") or piece.strip() in _string.punctuation)
        if not closed:
            continue
        if glue and words:
            words[-1] += piece
            groups[-1] += open_group
        else:
            words.append(piece)
            groups.append(open_group)
        rest = rest[len(piece):]
        open_group = []
    if open_group:                                   # undecodable tail: keeps whatever text is left
        words.append(rest if rest else piece)
        groups.append(open_group)
    elif rest:
        words[-1] += rest
    return words, groups


def window_word_script(segments: List[dict], tokenizer, gap_padding: Optional[str] = " ...", pad_first_seg: bool = True):
    """timing.py:344-392 without char_split: -> (text_tokens, words, word_tokens, seg_of_word); gap-padding pseudo-words
    have word None and are NOT counted in seg_of_word."""
    pad = None if gap_padding is None else (tokenizer.encode(gap_padding) if isinstance(gap_padding, str) else [gap_padding])
    text_tokens: List[int] = []
    words: List[Optional[str]] = []
    word_tokens: List[List[int]] = []
    seg_of_word: List[int] = []
    for si, seg in enumerate(segments):
        toks = [t for t in seg["tokens"] if t < tokenizer.eot]
        w, g = group_tokens_into_words(toks, tokenizer)
        # NB the reference compares a token id / list against the padding LIST, so these two tests are always True
        # for a str padding (timing.py:367-369); kept verbatim in meaning
        if pad is not None and g[0][0] != pad and (not text_tokens or text_tokens[-1] != pad) and (pad_first_seg or si):
            text_tokens += pad
            words.append(None)
            word_tokens.append(list(pad))
        seg_of_word += [si] * len(w)
        for grp in g:
            text_tokens += grp
        words += w
        word_tokens += g
    return text_tokens, words, word_tokens, seg_of_word


@torch.no_grad()
def word_timestamps_window(model: Whisper, tokenizer, segments: List[dict], mel: torch.Tensor, num_samples: int, *,
                           audio_features=None, gap_padding: Optional[str] = " ...", pad_first_seg: bool = True,
                           min_word_dur: float = 0.1, prepend_punctuations: str = PREPEND_PUNCTUATIONS,
                           append_punctuations: str = APPEND_PUNCTUATIONS, medfilt_width: int = 7, qk_scale: float = 1.0):
    """add_word_timestamps_stable for one window, legacy aligner (timing.py:411-500): fills seg['words'] and moves
    seg['start'|'end'] onto the first / last word."""
    if not segments:
        return
    for seg in segments:
        seg["words"] = []
    text_tokens, words, word_tokens, seg_of_word = window_word_script(segments, tokenizer, gap_padding, pad_first_seg)
    S = len(tokenizer.sot_sequence)
    _, qks, _, token_probs = window_qks(model, tokenizer, text_tokens, mel, audio_features)
    weights = attention_weights_legacy(qks, head_pairs_of(model), S, num_samples, medfilt_width, qk_scale)
    jumps = jumps_from_matrix(weights.mean(dim=0))
    bounds = np.pad(np.cumsum([len(g) for g in word_tokens]), (1, 0))
    times = jumps / TOKENS_PER_SECOND
    timed = [dict(word=w, tokens=list(g), start=float(times[a]), end=float(times[b]),
                  probability=float(np.mean(token_probs[a:b]))) for w, g, a, b in zip(words, word_tokens, bounds[:-1], bounds[1:])]
    # pull the gap-padding pseudo-words out; remember, per segment, the pseudo-word that precedes it (timing.py:395-407)
    pad_before: Dict[int, dict] = {}
    real: List[dict] = []
    for t in timed:
        if t["word"] is None:
            pad_before[seg_of_word[len(real)]] = t
        else:
            real.append(t)
    from types import SimpleNamespace
    real = [SimpleNamespace(**t) for t in real]
    _timing.merge_punctuations(real, prepend_punctuations, append_punctuations)
    offset = segments[0]["seek"]
    min_word_dur = min_word_dur or 0
    for si, t in zip(seg_of_word, real):
        if not t.tokens:
            continue
        start = t.start
        if not segments[si]["words"] and (t.end - t.start) < min_word_dur and si in pad_before:
            start = pad_before[si]["start"]
        segments[si]["words"].append(dict(word=t.word, start=round(offset + start, 3), end=round(offset + t.end, 3),
                                          probability=t.probability, tokens=t.tokens))
    for seg in segments:
        if seg["words"]:
            seg["start"], seg["end"] = seg["words"][0]["start"], seg["words"][-1]["end"]


def slice_window_segments(tokens: Sequence[int], tokenizer, time_offset: float, segment_duration: float, result,
                          word_timestamps: bool = True, punctuations: str = PREPEND_PUNCTUATIONS + APPEND_PUNCTUATIONS):
    """original_whisper.py:550-627: cut the sampled tokens at consecutive timestamp pairs, then drop punctuation-only
    and (with word timestamps) zero-length segments.  -> (segments, end_timestamp_pos, single_timestamp_ending)."""
    toks = list(tokens)
    tb = tokenizer.timestamp_begin
    stamp = [t >= tb for t in toks]
    single_ending = stamp[-2:] == [False, True]
    cuts = [i + 1 for i in range(len(toks) - 1) if stamp[i] and stamp[i + 1]]

    def make(start, end, part):
        return dict(seek=round(time_offset, 3), start=start, end=end, tokens=list(part),
                    text=tokenizer.decode([t for t in part if t < tokenizer.eot]), temperature=result.temperature,
                    avg_logprob=result.avg_logprob, compression_ratio=result.compression_ratio,
                    no_speech_prob=result.no_speech_prob)

    segs, end_pos = [], 0
    if cuts:
        if single_ending:
            cuts.append(len(toks))
        prev = 0
        for cut in cuts:
            part = toks[prev:cut]
            end_pos = part[-1] - tb
            segs.append(make(round(time_offset + (part[0] - tb) * TIME_PRECISION, 3),
                             round(time_offset + min(end_pos * TIME_PRECISION, segment_duration), 3), part))
            prev = cut
    else:
        dur = segment_duration
        stamps = [t for t in toks if t >= tb]
        if stamps and stamps[-1] != tb:
            end_pos = stamps[-1] - tb
            dur = min(end_pos * TIME_PRECISION, segment_duration)
        segs.append(make(round(time_offset, 3), round(time_offset + dur, 3), toks))
    segs = [s for s in segs if s["text"].strip() not in punctuations]       # substring test, as the reference's `in`
    if word_timestamps:
        segs = [s for s in segs if s["start"] != s["end"]]
    return segs, end_pos, single_ending


@torch.no_grad()
def transcribe_window(model: Whisper, tokenizer, audio: torch.Tensor, *, time_offset: float = 0.0,
                      forced_tokens: Optional[Sequence[int]] = None, ts_token_mask=None, gap_padding: Optional[str] = " ...",
                      min_word_dur: float = 0.1, max_instant_words: Optional[float] = None, **decode_options):
    """One <=30 s window the way transcribe_stable processes it at temperature 0 (original_whisper.py:500-665), without
    the cross-window state (prompt, seek): log-mel -> decode_stable -> slice -> add_word_timestamps_stable.
    ``forced_tokens``: fixed script appended instead of the argmax (random-weight benchmarks).
    -> (segments, extras{decode, step_argmax, tokens})"""
    n = int(audio.shape[-1])
    mel = _audio.pad_or_trim(_audio.log_mel_spectrogram(audio, model.dims.n_mels, padding=max(_audio.N_SAMPLES - n, 0)),
                             _audio.N_FRAMES)
    decode_options.setdefault("max_initial_timestamp", None)                # transcribe_stable's default (:262-263)
    res, af, ex = decode_window(model, mel, ts_token_mask=ts_token_mask, forced_tokens=forced_tokens, **decode_options)
    toks = list(forced_tokens)[: len(ex["step_argmax"])] if forced_tokens is not None else res.tokens
    extras = dict(decode=res, step_argmax=ex["step_argmax"], tokens=toks)
    if not toks:
        return [], extras
    segs, end_pos, _ = slice_window_segments(toks, tokenizer, time_offset, n / _audio.SAMPLE_RATE, res)
    num = min(round(end_pos * N_SAMPLES_PER_TOKEN), n) if end_pos > 0 else n
    word_timestamps_window(model, tokenizer, segs, mel, num, audio_features=af, gap_padding=gap_padding,
                           min_word_dur=min_word_dur)
    if max_instant_words is not None:                                        # original_whisper.py:655-663
        segs = [s for s in segs if not s["words"] or
                float(np.mean(np.array([w["start"] == w["end"] for w in s["words"]]).astype(np.float16))) <= max_instant_words]
    return segs, extras
