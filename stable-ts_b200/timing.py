"""Host-side mirror of stable_whisper/timing.py for the B200 path.

Same names, argument meaning and error behaviour as the reference (``WordTiming``, ``find_alignment_stable``,
``split_word_tokens``, ``pop_empty_alignment``, ``add_word_timestamps_stable``), but the per-window math is the
batched kernel pipeline of libstablets_b200.so:

    log-mel -> encoder -> cross K/V -> teacher-forced decoder (+QK capture of the alignment heads) -> token probs
    -> QK post-processing -> DTW (+jump extraction)                       [stable_whisper/timing.py:41-198]

Only the bookkeeping that the reference keeps in Python (word/token grouping, gap padding, punctuation merge,
rounding) runs on the host.  Windows are independent here, so ``align_windows``
takes a LIST of windows and runs them as one batch (the reference is batch 1, timing.py:60-61).
"""
import math
import string
from dataclasses import dataclass
from itertools import chain
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from .model import B200Whisper

N_SAMPLES = 480000
N_SAMPLES_PER_TOKEN = 320
TOKENS_PER_SECOND = 50


@dataclass
class WordTiming:          # stable_whisper/timing.py:22-28
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


@dataclass
class WindowJob:
    """One <=30 s window: audio (fp32, 16 kHz, <=480000 samples) or a mel / encoder output, plus its token script."""
    text_tokens: List[int]
    num_samples: int
    audio: Optional[torch.Tensor] = None


def n_frames_for(num_samples: int) -> int:
    return round(num_samples / N_SAMPLES_PER_TOKEN)     # Python banker's rounding, as timing.py:88,106


def token_row(tokenizer, text_tokens: Sequence[int]) -> List[int]:
    return [*tokenizer.sot_sequence, tokenizer.no_timestamps, *text_tokens, tokenizer.eot]


def window_batch_forward(model: B200Whisper, tokenizer, jobs: List[WindowJob], *, enc=None, ckv=None, heads=None,
                         want_logits: bool = True, reuse_buffers: bool = False):
    """Device side of ``_compute_qks`` for a batch of windows.  Returns dict(enc, ckv, logits, qk, M, S, rows)."""
    B = len(jobs)
    if enc is None:
        first = jobs[0].audio
        base = first._base if (first is not None and first._base is not None) else None
        if (base is not None and base.ndim == 2 and base.shape == (B, N_SAMPLES) and base.dtype == torch.float32
                and all(j.audio is not None and j.audio._base is base and j.audio.shape[-1] == N_SAMPLES for j in jobs)
                and all(j.audio.data_ptr() == base[i].data_ptr() for i, j in enumerate(jobs))):
            audio = base                                   # the windows are the rows of one [B, 480000] batch: no host copy
        else:
            audio = torch.zeros(B, N_SAMPLES, dtype=torch.float32)
            for i, j in enumerate(jobs):
                a = j.audio.detach().float().flatten()[:N_SAMPLES]
                audio[i, : a.numel()] = a
        if audio.device.type == "cpu" and model.device.type == "cuda" and not audio.is_pinned():
            audio = audio.pin_memory()
        audio = audio.to(model.device, non_blocking=True)
        mel = model.log_mel(audio)                         # == log_mel_spectrogram(audio, padding=N_SAMPLES-n)
        enc = model.encode(mel)
    if ckv is None:                                        # cross K/V of the window batch (reused from the decode pass)
        ckv = model.cross_kv(enc, reuse=reuse_buffers)
    S = len(tokenizer.sot_sequence)
    rows = [token_row(tokenizer, j.text_tokens) for j in jobs]
    M = max(len(r) for r in rows)
    tok = torch.full((B, M), int(tokenizer.eot), dtype=torch.int32)
    for i, r in enumerate(rows):
        tok[i, : len(r)] = torch.tensor(r, dtype=torch.int32)
    logits, qk = model.decode_forced(tok, ckv, want_logits=want_logits,
                                     heads=model.alignment_head_pairs if heads is None else heads, reuse=reuse_buffers)
    return dict(enc=enc, ckv=ckv, logits=logits, qk=qk, M=M, S=S, rows=rows)


def _parse_dynamic_heads(dynamic_heads):
    """(count, iterations) as stable_whisper/timing.py:254-267."""
    if not dynamic_heads:
        return None, 1
    if dynamic_heads is True:
        return 6, 1
    if isinstance(dynamic_heads, int):
        return int(dynamic_heads), 1
    assert "," in dynamic_heads
    c, i = dynamic_heads.split(",")
    return int(c), int(i)


def align_windows(model: B200Whisper, tokenizer, jobs: List[WindowJob], *, medfilt_width: int = 7, qk_scale: float = 1.0,
                  enc=None, ckv=None, dynamic_heads=None, aligner: Union[str, dict] = "legacy",
                  return_intermediates: bool = False, extra_models: Optional[Sequence[B200Whisper]] = None):
    """Batched equivalent of ``_compute_jump_indices``: legacy alignment heads, per-token dynamic heads
    (``dynamic_heads``: True | count | "count,iterations") or the "new" aligner (``aligner="new"`` or a dict of its
    options), stable_whisper/timing.py:70-198.

    ``extra_models`` (legacy / dynamic heads only, as in the reference, timing.py:177-189): every extra model runs its own
    encoder + teacher-forced pass on the windows' audio; the per-model attention weights are averaged over ALL heads of all
    models before the DTW (each model's head-mean weighted by its head count, combined on the device by stb_axpby) and the
    token probabilities are averaged over the models.

    -> list (per window) of (jump_indices int array [N+1], text_token_probs list[N]) (+ intermediates).
    """
    count, iters = _parse_dynamic_heads(dynamic_heads)
    new = aligner != "legacy"
    extra_models = list(extra_models or [])
    if extra_models and not new:
        bad = {type(m) for m in extra_models} - {type(model)}
        if bad:
            raise NotImplementedError(f"Got unsupported model type(s): {bad}")          # timing.py:219-220
        if any(j.audio is None for j in jobs):
            raise ValueError("extra_models: every window needs its audio (each model runs its own encoder)")
    if new:
        extra_models = []                                  # the "new" aligner ignores them (timing.py:174-175)
    if count is None and not new and getattr(model, "missing_alignment_heads", False):
        count = 6
    all_heads = bool(count) or new
    # logits / QK / cross K/V are consumed inside this function (only host arrays leave it), so they live in model-owned
    # buffers unless the caller asked for the intermediates
    fw = window_batch_forward(model, tokenizer, jobs, enc=enc, ckv=ckv, heads="all" if all_heads else None,
                              reuse_buffers=not return_intermediates)
    S, logits, qk, M = fw["S"], fw["logits"], fw["qk"], fw["M"]
    if extra_models:
        return _align_windows_multi(model, extra_models, tokenizer, jobs, fw, count, iters, medfilt_width, qk_scale)
    inter = []
    # windows with the same (N, F) share one post-processing / DTW launch
    groups = {}
    for i, j in enumerate(jobs):
        groups.setdefault((len(j.text_tokens), n_frames_for(j.num_samples)), []).append(i)
    results = [None] * len(jobs)
    for (N, F), idx in groups.items():
        sel = torch.tensor(idx, device=model.device)
        qk_g = qk if len(idx) == len(jobs) else qk.index_select(0, sel).contiguous()
        if new:
            kw = dict(aligner) if isinstance(aligner, dict) else {}
            kw.pop("char_split", None)                     # only changes the token script (add_word_timestamps_stable)
            # the reference slices [S:-1] of the decoder rows it ran (M_i = S + N + 2); padded rows are excluded by
            # running the scoring on a view of exactly those rows
            Mi = S + N + 2
            qk_i = qk_g if Mi == M else qk_g[:, :, :Mi].contiguous()
            matrix = model.qk_postprocess_new(qk_i, S, F, R=N + 1, qk_scale=qk_scale, medfilt_width=medfilt_width, **kw)
            jumps = model.dtw(matrix, negate=True)
        elif count:
            jumps = None
            for it in range(iters or 1):
                matrix = model.qk_postprocess_dynamic(qk_g, S, F, R=N + 1, count=count, prev_jumps=jumps,
                                                      reuse_softmax=it > 0, qk_scale=qk_scale, medfilt_width=medfilt_width)
                jumps = model.dtw(matrix, negate=True)
        else:
            matrix = model.qk_postprocess(qk_g, S, F, R=N + 1, qk_scale=qk_scale, medfilt_width=medfilt_width)
            jumps = model.dtw(matrix, negate=True)
        tgt = torch.tensor([jobs[i].text_tokens for i in idx], dtype=torch.int32).reshape(-1)
        if N > 0 and len(idx) == len(jobs):
            # the whole batch is one group: run the probability kernel over every decoder row in place (M / N ~ 2 % extra
            # rows with a dummy target) instead of gathering the text rows into a second multi-GB tensor
            tgt_full = torch.zeros(len(jobs), M, dtype=torch.int32)
            tgt_full[:, S:S + N] = tgt.view(len(jobs), N)
            p_full, _ = model.token_probs(logits.flatten(0, 1), tokenizer.eot, tgt_full.reshape(-1))
            probs = p_full.view(len(jobs), M)[:, S:S + N].reshape(-1)
        else:
            rows = torch.cat([logits[i, S:S + N] for i in idx]) if N > 0 else logits[:0, 0]
            probs, _ = model.token_probs(rows, tokenizer.eot, tgt) if N > 0 else (torch.empty(0), None)
        jumps_h = jumps.cpu().numpy()
        probs_h = probs.cpu().numpy().astype(np.float64).reshape(len(idx), N)
        for k, i in enumerate(idx):
            results[i] = (jumps_h[k].astype(np.int64), probs_h[k].tolist())
            if return_intermediates:
                inter.append((i, matrix[k].cpu()))
    if return_intermediates:
        inter = [m for _, m in sorted(inter, key=lambda t: t[0])]
        return results, dict(forward=fw, matrices=inter)
    return results


def _align_windows_multi(model, extra_models, tokenizer, jobs, fw_main, count, iters, medfilt_width, qk_scale):
    """``align_windows`` with ``extra_models`` (timing.py:177-189).  One (N, F) group at a time, every model's forward done
    once; per dynamic-heads iteration the matrices of all models are combined and one DTW gives the shared jumps."""
    models = [model] + list(extra_models)
    fws = [fw_main] + [window_batch_forward(m, tokenizer, jobs, heads="all" if count else None, reuse_buffers=False)
                       for m in extra_models]
    S, M = fw_main["S"], fw_main["M"]
    groups = {}
    for i, j in enumerate(jobs):
        groups.setdefault((len(j.text_tokens), n_frames_for(j.num_samples)), []).append(i)
    results = [None] * len(jobs)
    for (N, F), idx in groups.items():
        tgt = torch.tensor([jobs[i].text_tokens for i in idx], dtype=torch.int32).reshape(-1)
        probs = []
        for m, fw in zip(models, fws):
            rows = torch.cat([fw["logits"][i, S:S + N] for i in idx]) if N > 0 else fw["logits"][:0, 0]
            p = m.token_probs(rows, tokenizer.eot, tgt)[0].cpu().numpy().astype(np.float64) if N > 0 else np.zeros(0)
            probs.append(p.reshape(len(idx), N))
        heads = [count if count else len(m.alignment_head_pairs) for m in models]
        total = float(sum(heads))
        jumps = None
        # the reference replaces the main cache's probabilities by the mean over [extras..., main] on EVERY iteration
        # (timing.py:183-189), so from the second iteration on the previous mean takes the main model's place
        main_probs = probs[0]
        for it in range(iters or 1):
            combined = None
            for m, fw, h in zip(models, fws, heads):
                sel = torch.tensor(idx, device=m.device)
                qk_g = fw["qk"] if len(idx) == len(jobs) else fw["qk"].index_select(0, sel).contiguous()
                if count:
                    # only the MAIN model's cache ever receives jump_indices (timing.py:198): the extra models pick their
                    # heads from the attention peaks on every iteration
                    mat = m.qk_postprocess_dynamic(qk_g, S, F, R=N + 1, count=count, prev_jumps=jumps if m is model else None,
                                                   reuse_softmax=it > 0, qk_scale=qk_scale, medfilt_width=medfilt_width)
                else:
                    mat = m.qk_postprocess(qk_g, S, F, R=N + 1, qk_scale=qk_scale, medfilt_width=medfilt_width)
                if combined is None:
                    combined = model.scale_add(mat, mat, h / total, 0.0)
                else:
                    model.scale_add(combined, mat, h / total, 1.0)
            jumps = model.dtw(combined, negate=True)
            main_probs = np.mean(np.stack(probs[1:] + [main_probs]), axis=0)
        jumps_h = jumps.cpu().numpy()
        for k, i in enumerate(idx):
            results[i] = (jumps_h[k].astype(np.int64), main_probs[k].tolist())
    return results


def word_timings_from_jumps(jumps: np.ndarray, token_probs: List[float], words, word_tokens, ignore_tokens=None,
                            word_tokens_out=None) -> List[WordTiming]:
    """stable_whisper/timing.py:251-253,289-306; ``word_tokens`` already ends with the [eot] pseudo-word.
    ``ignore_tokens`` / ``word_tokens_out``: the char_split form (timing.py:240-253,299-301): boundaries are counted in the
    character tokens, a word that starts with the space token starts one token later, and the returned WordTiming carries
    the word's ORIGINAL tokens."""
    wb = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    if ignore_tokens:
        itk = list(ignore_tokens)
        wb = wb + np.array([list(t[:len(itk)]) == itk for t in word_tokens], dtype=wb.dtype)
    jump_times = jumps / TOKENS_PER_SECOND
    # plain Python floats from here on: same float64 values as the reference's numpy scalars, without ~3 us of numpy
    # scalar overhead per word and per round() on the host path (120 windows x ~170 words per step)
    starts, ends = jump_times[wb[:-1]].tolist(), jump_times[wb[1:]].tolist()
    bounds = wb.tolist()
    tp = token_probs if isinstance(token_probs, list) else list(token_probs)
    probs = [(math.fsum(tp[i:j]) / (j - i)) if j > i else float("nan") for i, j in zip(bounds[:-1], bounds[1:])]
    if word_tokens_out is not None:
        assert len(word_tokens_out) == len(word_tokens)
        word_tokens = word_tokens_out
    return [WordTiming(w, t, s, e, p) for w, t, s, e, p in zip(words, word_tokens, starts, ends, probs)]


def find_alignment_stable(model: B200Whisper, tokenizer, text_tokens: List[int], audio: Optional[torch.Tensor],
                          num_samples: int, *, medfilt_width: int = 7, qk_scale: float = 1.0, token_split=None,
                          enc=None, dynamic_heads=None, aligner: Union[str, dict] = "legacy", extra_models=None,
                          ts_num: int = 0, ts_noise=None) -> List[WordTiming]:
    """One window (stable_whisper/timing.py:202-306).  ``audio`` replaces ``mel`` (the log-mel runs on the device).
    ``token_split`` = (words, word_tokens) or, for the "new" aligner's char_split, (words, dict(tokens, tokens_orig,
    ignore_tokens)) as ``split_word_tokens`` returns it (timing.py:240-246)."""
    assert isinstance(aligner, dict) or aligner in ("new", "legacy"), f'aligner must be "new"/"legacy", got "{aligner}"'
    orig = itk = None
    if token_split is None:
        words, word_tokens = tokenizer.split_to_word_tokens(list(text_tokens) + [tokenizer.eot])
    else:
        words, word_tokens = token_split
        if isinstance(word_tokens, dict):
            orig = list(word_tokens["tokens_orig"]) + [[tokenizer.eot]]
            itk = word_tokens["ignore_tokens"]
            word_tokens = word_tokens["tokens"]
        words = list(words) + [tokenizer.decode([tokenizer.eot])]
        word_tokens = list(word_tokens) + [[tokenizer.eot]]
    job = WindowJob(list(text_tokens), num_samples, audio)
    (jumps, probs), = align_windows(model, tokenizer, [job], medfilt_width=medfilt_width, qk_scale=qk_scale, enc=enc,
                                    dynamic_heads=dynamic_heads, aligner=aligner, extra_models=extra_models)
    return word_timings_from_jumps(jumps, probs, words, word_tokens, ignore_tokens=itk, word_tokens_out=orig)


# ---------------------------------------------------------------------------------------------------------
# bookkeeping the reference keeps on the host
# ---------------------------------------------------------------------------------------------------------
def _split_tokens(tokens: List[int], tokenizer):
    """Group tokens into words by incremental decoding (stable_whisper/timing.py:309-341)."""
    lang = getattr(tokenizer, "language_code", getattr(tokenizer, "language", None))
    by_space = lang not in {"zh", "ja", "th", "lo", "my"}
    remaining = tokenizer.decode_with_timestamps(tokens)
    words, groups, cur = [], [], []
    append_to_prev = False
    cur_text = ""
    for tok in tokens:
        cur.append(tok)
        cur_text = tokenizer.decode(cur)
        whole = tok >= tokenizer.eot
        if not whole:
            whole = remaining[: len(cur_text)] == cur_text
            if whole and by_space:
                append_to_prev = not (cur_text.startswith(" ") or cur_text.strip() in string.punctuation)
        if whole:
            if append_to_prev and words:
                words[-1] += cur_text
                groups[-1].extend(cur)
            else:
                words.append(cur_text)
                groups.append(cur)
            remaining = remaining[len(cur_text):]
            cur = []
    if cur:
        words.append(cur_text if len(remaining) == 0 else remaining)
        groups.append(cur)
    elif remaining:
        words[-1] += remaining
    return words, groups


def split_word_tokens(segments: List[dict], tokenizer, *, padding: Union[str, int, None] = None,
                      split_callback: Optional[Callable] = None, pad_first_seg: bool = True, char_split: bool = False):
    """stable_whisper/timing.py:344-392.  char_split (the "new" aligner's character-level script): the token script is the
    encoding of every character of every word; ``word_tokens`` becomes dict(tokens=per-word character tokens,
    tokens_orig=the words' own tokens, ignore_tokens=encode(' '))."""
    if padding is not None:
        padding = tokenizer.encode(padding) if isinstance(padding, str) else [padding]
    tokens, seg_indices, words, word_tokens, word_char_tokens = [], [], [], [], []
    for i, seg in enumerate(segments):
        text_toks = [t for t in seg["tokens"] if not isinstance(t, int) or t < tokenizer.eot]
        cw, cwt = _split_tokens(text_toks, tokenizer) if split_callback is None else split_callback(text_toks, tokenizer)
        assert len(cw) == len(cwt), f"word count and token group count do not match, {len(cw)} and {len(cwt)}"
        if (padding is not None and cwt[0][0] != padding and (len(tokens) == 0 or tokens[-1] != padding)
                and (pad_first_seg or i != 0)):
            tokens.extend(padding)
            words.append(None)
            word_tokens.append(padding)
        seg_indices.extend([i] * len(cw))
        if char_split:
            cct = [[ct for ch in word for ct in tokenizer.encode(ch)] for word in cw]
            word_char_tokens.extend(cct)
            tokens.extend(chain.from_iterable(cct))
        else:
            tokens.extend(chain.from_iterable(cwt))
        words.extend(cw)
        word_tokens.extend(cwt)
    if char_split:
        word_tokens = dict(tokens=word_char_tokens, tokens_orig=word_tokens, ignore_tokens=tokenizer.encode(" "))
    return tokens, (words, word_tokens), seg_indices


def pop_empty_alignment(alignment: List[WordTiming], seg_indices: Optional[List[int]] = None):
    """Remove the gap-padding pseudo-words (word is None); stable_whisper/timing.py:395-407."""
    if seg_indices is None:
        kept = [a for a in alignment if a.word is None]
        alignment[:] = [a for a in alignment if a.word is not None]
        return kept
    pos = len(seg_indices)
    popped = {}
    for i in reversed(range(len(alignment))):
        assert pos != -1
        if alignment[i].word is None:
            popped[seg_indices[pos]] = alignment.pop(i)
        else:
            pos -= 1
    return popped


def merge_punctuations(alignment: List[WordTiming], prepended: str, appended: str):
    """whisper.timing.merge_punctuations semantics."""
    i, j = len(alignment) - 2, len(alignment) - 1
    while i >= 0:
        prev, nxt = alignment[i], alignment[j]
        if prev.word.startswith(" ") and prev.word.strip() in prepended:
            nxt.word, nxt.tokens = prev.word + nxt.word, prev.tokens + nxt.tokens
            prev.word, prev.tokens = "", []
        else:
            j = i
        i -= 1
    i, j = 0, 1
    while j < len(alignment):
        prev, nxt = alignment[i], alignment[j]
        if not prev.word.endswith(" ") and nxt.word in appended:
            prev.word, prev.tokens = prev.word + nxt.word, prev.tokens + nxt.tokens
            nxt.word, nxt.tokens = "", []
        else:
            i = j
        j += 1


PREPEND_PUNCT = "\"'“¿([{-"
APPEND_PUNCT = "\"'.。,，!！?？:：”)]}、"


def _prepare_word_timestamps(segments, tokenizer, split_callback, gap_padding, pad_first_seg, char_split: bool = False):
    """-> (text_tokens, words, word_tokens, seg_indices, ignore_tokens, word_tokens_out); the last two are None unless
    ``char_split`` (the "new" aligner's character-level script: no gap padding, timing.py:380-390,442-444)."""
    for seg in segments:
        seg["words"] = []
    text_tokens, (words, word_tokens), seg_indices = split_word_tokens(segments, tokenizer,
                                                                       padding=None if char_split else gap_padding,
                                                                       split_callback=split_callback,
                                                                       pad_first_seg=pad_first_seg, char_split=char_split)
    words = list(words) + [tokenizer.decode([tokenizer.eot])]
    itk = orig = None
    if isinstance(word_tokens, dict):
        orig = list(word_tokens["tokens_orig"]) + [[tokenizer.eot]]
        itk = word_tokens["ignore_tokens"]
        word_tokens = word_tokens["tokens"]
    word_tokens = list(word_tokens) + [[tokenizer.eot]]
    return text_tokens, words, word_tokens, seg_indices, itk, orig


def _finish_word_timestamps(segments, alignment, seg_indices, prepend_punctuations, append_punctuations, min_word_dur,
                            gap_padding, pad_first_seg):
    alt_begin = pop_empty_alignment(alignment, seg_indices)
    merge_punctuations(alignment, prepend_punctuations, append_punctuations)
    offset = segments[0]["seek"]
    assert len(alignment) == len(seg_indices)
    assert gap_padding is None or len(segments) == len(alt_begin) + (1, 0)[pad_first_seg]
    for i, timing in zip(seg_indices, alignment):
        if len(timing.tokens) == 0:
            continue
        start, end = timing.start, timing.end
        if len(segments[i]["words"]) == 0 and (end - start) < min_word_dur and i in alt_begin:
            start = alt_begin[i].start
        segments[i]["words"].append(dict(word=timing.word, start=round(offset + start, 3), end=round(offset + end, 3),
                                         probability=timing.probability, tokens=timing.tokens))
    for seg in segments:
        if seg["words"]:
            seg["start"] = seg["words"][0]["start"]
            seg["end"] = seg["words"][-1]["end"]


def add_word_timestamps_batch(windows: List[dict], model: B200Whisper, tokenizer, *, enc=None, ckv=None,
                              prepend_punctuations: Optional[str] = PREPEND_PUNCT,
                              append_punctuations: Optional[str] = APPEND_PUNCT, min_word_dur: float = 0.1,
                              split_callback: Optional[Callable] = None, gap_padding: Optional[str] = " ...",
                              pad_first_seg: bool = True, medfilt_width: int = 7, qk_scale: float = 1.0, dynamic_heads=None,
                              aligner: Union[str, dict] = "legacy", extra_models=None):
    """Batched ``add_word_timestamps_stable``: ``windows`` = list of dict(segments, num_samples[, audio]); all windows
    share ONE encoder/decoder/DTW batch (``enc`` = the dict from ``model.encode`` for the same windows, in order).
    Windows without segments are skipped (timing.py:430-431).  ``dynamic_heads`` / ``aligner`` / ``extra_models`` as in
    ``add_word_timestamps_stable``; ``extra_models`` need ``audio`` in every window dict.  ``char_split`` is popped from the
    caller's ``aligner`` dict exactly as the reference pops it (timing.py:442), so -- as there -- only the FIRST window
    processed with that dict is character-split."""
    min_word_dur = min_word_dur or 0
    prepend_punctuations = PREPEND_PUNCT if prepend_punctuations is None else prepend_punctuations
    append_punctuations = APPEND_PUNCT if append_punctuations is None else append_punctuations
    live = [i for i, w in enumerate(windows) if len(w["segments"]) > 0]
    if not live:
        return
    char_split = bool(isinstance(aligner, dict) and aligner.pop("char_split", False))
    pads = {i: (None if (char_split and i == live[0]) else gap_padding) for i in live}
    prep = {i: _prepare_word_timestamps(windows[i]["segments"], tokenizer, split_callback, pads[i], pad_first_seg,
                                        char_split=char_split and i == live[0]) for i in live}
    jobs = [WindowJob(list(prep[i][0]), int(windows[i]["num_samples"]), windows[i].get("audio")) for i in live]
    sub = enc
    if enc is None or len(live) != enc["B"]:
        ckv = None                                          # cached cross K/V only matches the full batch
    if enc is not None and len(live) != enc["B"]:
        from .decode import enc_select
        sub = enc_select(enc, live)
    res = align_windows(model, tokenizer, jobs, medfilt_width=medfilt_width, qk_scale=qk_scale, enc=sub, ckv=ckv,
                        dynamic_heads=dynamic_heads, aligner=aligner, extra_models=extra_models)
    for (jumps, probs), i in zip(res, live):
        _, words, word_tokens, seg_indices, itk, orig = prep[i]
        alignment = word_timings_from_jumps(jumps, probs, words, word_tokens, ignore_tokens=itk, word_tokens_out=orig)
        _finish_word_timestamps(windows[i]["segments"], alignment, seg_indices, prepend_punctuations, append_punctuations,
                                min_word_dur, pads[i], pad_first_seg)


def add_word_timestamps_stable(*, segments: List[dict], model: B200Whisper, tokenizer, audio: Optional[torch.Tensor] = None,
                               num_samples: int, prepend_punctuations: Optional[str] = PREPEND_PUNCT,
                               append_punctuations: Optional[str] = APPEND_PUNCT, enc=None, min_word_dur: float = 0.1,
                               split_callback: Optional[Callable] = None, gap_padding: Optional[str] = " ...",
                               pad_first_seg: bool = True, aligner="legacy", mel: Optional[torch.Tensor] = None,
                               audio_features: Optional[torch.Tensor] = None, ts_num: int = 0, ts_noise=None, **kwargs):
    """Mutates ``segments[i]['words'|'start'|'end']`` in place (stable_whisper/timing.py:411-500).

    The window is given as ``audio`` (fp32 samples; the log-mel then runs on the device), as ``mel`` [n_mels, 3000] (the
    reference's argument, timing.py:416) or as ``enc`` / ``audio_features`` (encoder output, timing.py:421)."""
    if len(segments) == 0:
        return
    if ts_num or ts_noise:
        import warnings
        warnings.warn("ts_num and ts_noise are deprecated and will be removed in future versions.", stacklevel=2)
    min_word_dur = min_word_dur or 0
    prepend_punctuations = PREPEND_PUNCT if prepend_punctuations is None else prepend_punctuations
    append_punctuations = APPEND_PUNCT if append_punctuations is None else append_punctuations
    if enc is None and audio_features is not None:
        enc = model._encoding_of(audio_features)
    if enc is None and audio is None and mel is not None:
        enc = model.encode(mel.to(model.device, torch.float32))
    # timing.py:442-444: char_split belongs to the "new" aligner's options and switches the gap padding off
    # (the key is popped from the CALLER's dict, exactly as the reference does: with one options object shared by all windows
    # of an align() call only the first window is character-split there, and therefore here)
    char_split = bool(isinstance(aligner, dict) and aligner.pop("char_split", False))
    if char_split:
        gap_padding = None
    for seg in segments:
        seg["words"] = []
    text_tokens, token_split, seg_indices = split_word_tokens(segments, tokenizer, padding=gap_padding, split_callback=split_callback,
                                                              pad_first_seg=pad_first_seg, char_split=char_split)
    alignment = find_alignment_stable(model, tokenizer, text_tokens, audio, num_samples, token_split=token_split,
                                      enc=enc, aligner=aligner, **kwargs)
    _finish_word_timestamps(segments, alignment, seg_indices, prepend_punctuations, append_punctuations, min_word_dur,
                            gap_padding, pad_first_seg)
