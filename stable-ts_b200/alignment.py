"""Plugin closures and batched drivers for the B200 path (mirror of stable_whisper/alignment.py:396-509, 640-753).

``get_b200_alignment_func`` / ``get_b200_refinement_func`` return callables with EXACTLY the contract the reference's
model-agnostic ``Aligner`` / ``Refiner`` document (stable_whisper/non_whisper/alignment.py:85-88,
refinement.py:44-47), so a reference install can drive them unchanged.  The batched entry points
(``align_words_batch``, ``refine_probs``) expose what the GPU is good at: many independent 30 s windows per call.
"""
from typing import List, Optional, Sequence

import torch

from .model import B200Whisper
from .timing import WindowJob, add_word_timestamps_stable, align_windows, token_row, word_timings_from_jumps

N_SAMPLES = 480000


def get_b200_alignment_func(model: B200Whisper, tokenizer, options=None):
    """-> compute_timestamps(audio_segment fp32 [n<=480000], word_tokens) -> list of word dicts
    (same closure as stable_whisper/alignment.py:405-429: no gap padding, identity split, no punctuation merge)."""
    al = getattr(options, "align", None) if options is not None else None
    extra = getattr(al, "extra_models", None) if al is not None else None
    dyn = getattr(al, "dynamic_heads", None) if al is not None else None
    aligner = getattr(al, "aligner", "legacy") if al is not None else "legacy"

    def compute_timestamps(audio_segment: torch.Tensor, word_tokens) -> List[dict]:
        words = [wt.word for wt in word_tokens]
        toks = [list(wt.tokens) for wt in word_tokens]
        seg = [dict(seek=0, tokens=(words, toks))]
        add_word_timestamps_stable(segments=seg, model=model, tokenizer=tokenizer, audio=audio_segment,
                                   num_samples=int(audio_segment.size(-1)), split_callback=(lambda x, _: x),
                                   prepend_punctuations="", append_punctuations="", gap_padding=None, dynamic_heads=dyn,
                                   aligner=aligner, extra_models=extra)
        return [w for s in seg for w in s["words"]]

    return compute_timestamps


def align_words_batch(model: B200Whisper, tokenizer, audios: Sequence[torch.Tensor],
                      word_tokens: Sequence[List[List[int]]], words: Optional[Sequence[List[str]]] = None,
                      *, medfilt_width: int = 7, qk_scale: float = 1.0, dynamic_heads=None, aligner="legacy",
                      return_intermediates: bool = False):
    """Many independent windows in ONE batch (the natural GPU form of ``align_words``: one window per pre-timed
    segment, stable_whisper/non_whisper/alignment.py:443-465).  -> per window list of word dicts."""
    jobs = [WindowJob([t for w in wt for t in w], int(a.shape[-1]), a) for a, wt in zip(audios, word_tokens)]
    res = align_windows(model, tokenizer, jobs, medfilt_width=medfilt_width, qk_scale=qk_scale, dynamic_heads=dynamic_heads,
                        aligner=aligner, return_intermediates=return_intermediates)
    inter = None
    if return_intermediates:
        res, inter = res
    out = []
    for k, ((jumps, probs), wt) in enumerate(zip(res, word_tokens)):
        ws = [None] * len(wt) if words is None else list(words[k])
        tim = word_timings_from_jumps(jumps, probs, ws + ["<eot>"], [list(w) for w in wt] + [[tokenizer.eot]])
        out.append([dict(word=t.word, start=round(float(t.start), 3), end=round(float(t.end), 3),
                         probability=float(t.probability), tokens=t.tokens) for t in tim if len(t.tokens)])
    return (out, inter) if return_intermediates else out


def _refine_logit_rows(model: B200Whisper, tokenizer, audio_segment: torch.Tensor, tokens: Sequence[int]):
    """alignment.py:649-668: log-mel of [2, n] with the batch-global max and no sample padding, one teacher-forced pass with
    the token row broadcast over the audio rows -> logits rows [2 * N, V_pad] of the script positions (device view)."""
    a = audio_segment.to(model.device, torch.float32)
    n = int(a.shape[-1])
    mel = model.log_mel(a, padded_samples=n, batch_global_max=True)
    enc = model.encode(mel)
    ckv = model.cross_kv(enc)
    row = torch.tensor([token_row(tokenizer, tokens)] * a.shape[0], dtype=torch.int32)
    logits, _ = model.decode_forced(row, ckv)
    S, N = len(tokenizer.sot_sequence), len(tokens)
    return torch.cat([logits[b, S:S + N] for b in range(a.shape[0])]), a.shape[0], N


def refine_probs(model: B200Whisper, tokenizer, audio_segment: torch.Tensor, tokens: Sequence[int], want_rank: bool = True):
    """audio fp32 [2, n] -> (probs fp32 [2, N], rank int32 [2, N]) of the script tokens
    (stable_whisper/alignment.py:649-672 + the gather/rank of refinement.py:305-325, without materialising [2,N,V]).
    rank = number of classes whose logit is strictly below the target's = the target's index in the ascending sort."""
    rows, A, N = _refine_logit_rows(model, tokenizer, audio_segment, tokens)
    tgt = torch.tensor(list(tokens) * A, dtype=torch.int32)
    p, r = model.token_probs(rows, tokenizer.eot, tgt, want_rank=want_rank)
    return p.view(A, N), (r.view(A, N) if r is not None else None)


def get_b200_refinement_func(model: B200Whisper, tokenizer, form: str = "3d"):
    """-> inference_func(audio [2, n], tokens) for the reference's ``Refiner`` (non_whisper/refinement.py:44-47).

    form="3d" (default): Tensor [2, N, eot] of softmax probabilities on the model's device -- the same tensor the reference's
    own closure returns (alignment.py:669-672), so the Refiner's token-rank test (refinement.py:305-325, ``best_tks_changed``
    at :427) sees exactly what it sees with a PyTorch model.  form="2d": Tensor [2, N] of the script tokens' probabilities
    only (the other form refinement.py:291-304 accepts; the rank test is then off, as the reference defines it)."""
    if form not in ("3d", "2d"):
        raise ValueError("form must be '3d' or '2d'")

    def inference_func(audio_segment: torch.Tensor, tokens: List[int]) -> torch.Tensor:
        if form == "2d":
            p, _ = refine_probs(model, tokenizer, audio_segment, tokens, want_rank=False)
            return p.cpu()
        rows, A, N = _refine_logit_rows(model, tokenizer, audio_segment, tokens)
        return model.softmax_probs(rows, tokenizer.eot).view(A, N, int(tokenizer.eot))
    return inference_func
