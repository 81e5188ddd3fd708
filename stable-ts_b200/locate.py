"""``locate()`` on the B200 path (mirror of stable_whisper/alignment.py:756-1116).

Per 30 s chunk the reference (a) runs the encoder and a teacher-forced decoder pass over ``initial_tokens + text_tokens``
with the cross-attention QK of the alignment heads captured, softmax / z-norm / median / head-mean, and takes the argmax of
the LAST row as the time the text ends (:920-949); (b) unless ``mode == 2``, re-encodes a short window around that time and
greedily decodes it token by token, forcing the target tokens when they are probable enough, to confirm the match
(:974-1060); (c) on a confirmed match, word-aligns the decoded tokens (:1088-1095).

(a) and (c) are the same kernels as ``align`` (``stb_logmel`` / ``stb_encoder_forward`` / ``stb_decoder_forward`` /
``stb_qk_postprocess`` with S = 0 and all rows / ``stb_dtw``); (b) is a data-dependent, one-token-at-a-time host loop in the
reference as well, and runs here over the KV-cached ``stb_decode_step`` through the whisper-protocol decoder of shim.py.
"""
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from .decode import DecodingOptions, _suppress_list
from .timing import add_word_timestamps_stable, split_word_tokens

SAMPLE_RATE = 16000
CHUNK_SAMPLES = 480000
N_FRAMES = 3000
FRAMES_PER_SECOND = 100
N_FFT = 400


def _segment_class():
    try:
        from stable_whisper.result import Segment
        return Segment
    except Exception:
        from .result import Segment
        return Segment


def window_target_end(model, tokens: Sequence[int], audio_segment: torch.Tensor):
    """(a): -> (target_end seconds within the chunk, mel [1, n_mels, 3000], enc) for one chunk (alignment.py:920-949)."""
    n = int(audio_segment.shape[-1])
    padded = min(n + N_FFT // 2 + 1, CHUNK_SAMPLES + 160)      # padding=201 samples; frames >= 3000 are trimmed either way
    mel = model.log_mel(audio_segment.to(model.device)[None], padded_samples=max(padded, 201))
    enc = model.encode(mel)
    ckv = model.cross_kv(enc)
    row = torch.tensor([list(tokens)], dtype=torch.int32)
    _, qk = model.decode_forced(row, ckv, want_logits=False, heads=model.alignment_head_pairs)
    # every decoder row, every one of the 1500 columns: softmax -> z-norm over rows -> median(7) -> mean over heads
    matrix = model.qk_postprocess(qk, S=0, F=L.N_AUDIO_CTX, R=len(tokens))
    sec_per_emb = model.dims.n_audio_ctx / 30
    target_end = round((matrix[0, -1].argmax() / sec_per_emb).item(), 3)
    return target_end, mel, enc


@torch.no_grad()
def locate(model, audio: torch.Tensor, text: Union[str, List[int]], language: str, count: int = 1,
           duration_window: Union[float, Tuple[float, float]] = 3.0, *, mode: int = 0, start: float = None, end: float = None,
           probability_threshold: float = 0.5, eots: int = 1, max_token_per_seg: int = 20, exact_token: bool = False,
           case_sensitive: bool = False, verbose: bool = False, initial_prompt: str = None,
           suppress_tokens: Union[str, List[int]] = "-1", tokenizer=None, **unsupported):
    """Same arguments and return values as the reference's ``locate`` (list of Segment / dict); ``audio`` is a 16 kHz fp32
    waveform.  Denoiser / demucs / only_voice_freq pre-processing is control plane and not available here."""
    for k, v in unsupported.items():
        if v not in (None, False):
            raise NotImplementedError(f"B200 locate: option {k} is not implemented")
    from .tokenizer import get_tokenizer
    tk = tokenizer or get_tokenizer(model, language=language, task="transcribe", synthetic=getattr(model, "random_init", False))
    if isinstance(duration_window, (float, int)):
        duration_window = [duration_window] * 2
    window_sum = sum(duration_window)
    assert CHUNK_SAMPLES > window_sum, f"Sum of [duration_window] must be less than {CHUNK_SAMPLES}, got {window_sum}"
    adjusted_chunk_size = CHUNK_SAMPLES - round(duration_window[0] * SAMPLE_RATE)
    # DecodingTask(..., without_timestamps=True).initial_tokens (whisper decoding.py _get_initial_tokens)
    initial_tokens = list(tk.sot_sequence_including_notimestamps)
    if initial_prompt:
        ptoks = tk.encode(" " + initial_prompt.strip())
        initial_tokens = [tk.sot_prev] + ptoks[-(model.dims.n_text_ctx // 2 - 1):] + initial_tokens
    text_tokens, text = (tk.encode(text), text) if isinstance(text, str) else (list(text), tk.decode(text))
    if not exact_token and not case_sensitive:
        text = text.lower()
    suppress = [i for i in _suppress_list(tk, DecodingOptions(suppress_tokens=suppress_tokens)) if i < tk.eot] \
        if suppress_tokens else []
    suppress_t = torch.tensor(suppress, dtype=torch.long, device=model.device)
    audio = audio.detach().float().flatten()
    if end:
        audio = audio[: round(end * SAMPLE_RATE)]
    seek_sample = round(start * SAMPLE_RATE) if start else 0
    total_samples = int(audio.shape[-1])
    Segment = _segment_class()
    state = dict(found=0, prev_target_end=None)

    def one_chunk():
        nonlocal seek_sample
        seek = round(seek_sample / SAMPLE_RATE, 3)
        audio_segment = audio[seek_sample: seek_sample + CHUNK_SAMPLES]
        target_end, mel_segment, enc = window_target_end(model, initial_tokens + text_tokens, audio_segment)
        if verbose:
            print(f'"{text}" ending at ~{target_end + seek:.2f}s')
        if mode == 2:
            state["found"] += 1
            if (seek_sample + CHUNK_SAMPLES >= total_samples) or (count and state["found"] >= count) or \
                    (state["prev_target_end"] == target_end):
                seek_sample = total_samples
            else:
                seek_sample += round(target_end * SAMPLE_RATE)
            state["prev_target_end"] = target_end
            return dict(tokens=[], target_end=target_end + seek)

        curr_start = round(max(target_end - duration_window[0], 0.0), 3)
        curr_end = round(target_end + duration_window[1], 3)
        start_frame, end_frame = round(curr_start * FRAMES_PER_SECOND), round(curr_end * FRAMES_PER_SECOND)
        section = torch.zeros_like(mel_segment)
        part = mel_segment[..., start_frame:end_frame][..., :N_FRAMES]
        section[..., : part.shape[-1]] = part                                  # pad_or_trim(mel[..., a:b], 3000)
        xa = model.encoder(section)
        temp_tokens = torch.tensor([initial_tokens], dtype=torch.int32)
        infer_tokens: List[int] = list(initial_tokens)
        predictions, tokens_to_decode, replace_found = [], [], []
        target_idx, curr_eots, found_target, not_end = 0, 0, False, True
        kv_cache, hooks = model.install_kv_cache_hooks()
        while not_end:
            logits = model.decoder(temp_tokens, xa, kv_cache=kv_cache)[0, -1, : tk.eot + 1].clone()
            if len(suppress):
                logits[suppress_t] = -np.inf
            top2 = logits.sort(dim=-1).indices[-2:].tolist()
            best_token = top2[-1]
            best_non_eot = top2[-2] if best_token == tk.eot else best_token
            # In the reference both names are views of ONE tensor element when the best token is not EOT (:990-991), so its
            # in-place write `best_token[None] = target` (:1013) is also seen through `best_non_eot_token` and through the
            # entry just appended to `tokens_to_decode`; the observable behaviour is reproduced explicitly below.
            aliased = best_token != tk.eot
            probs = logits[: tk.eot].softmax(dim=-1)
            if found_target:
                target_prob = is_match = None
            else:
                if exact_token:
                    is_match = False
                else:
                    tokens_to_decode.append(best_non_eot)
                    temp_text = tk.decode(tokens_to_decode)
                    if not case_sensitive:
                        temp_text = temp_text.lower()
                    is_match = temp_text.endswith(text)
                    if is_match:
                        tokens_to_decode = []
                target_prob = probs[text_tokens[target_idx]].item()
            if target_prob is not None and (target_prob >= probability_threshold or best_non_eot == text_tokens[target_idx]
                                            or is_match):
                if is_match:
                    best_token = best_non_eot
                    token_prob = probs[best_token].item()
                    found_target = True
                else:
                    best_token = text_tokens[target_idx]
                    if aliased:
                        best_non_eot = best_token
                        if not exact_token and tokens_to_decode:
                            tokens_to_decode[-1] = best_token
                    if len(replace_found) or best_non_eot != text_tokens[target_idx]:
                        replace_found.append(best_non_eot)
                    target_idx += 1
                    if target_idx == len(text_tokens):
                        found_target = True
                    token_prob = target_prob
                if found_target:
                    state["found"] += 1
                curr_eots = 0
            else:
                if not found_target:
                    if len(replace_found):
                        # un-force (alignment.py:1027-1035).  The reference rebuilds a context from its own picks here, but
                        # the rebuilt `temp_tokens` is overwritten a few lines later (:1053-1054), so what it actually does
                        # is: drop the KV cache and continue from the next token alone.  (Its `torch.cat` of the two parts
                        # also runs along dim 0 and raises unless they happen to have equal lengths; that crash is not
                        # reproduced.)
                        replace_found = []
                        kv_cache.clear()
                    target_idx = 0
                if best_token == tk.eot:
                    if curr_eots >= eots or found_target:
                        not_end = False
                    else:
                        curr_eots += 1
                        best_token = best_non_eot
                else:
                    curr_eots = 0
                token_prob = None if best_token == tk.eot else probs[best_token].item()
            predictions.append(dict(token=int(best_token), prob=token_prob))
            if len(predictions) > max_token_per_seg:
                not_end = False
            if not_end:
                infer_tokens.append(int(best_token))
                temp_tokens = torch.tensor([[int(best_token)]], dtype=torch.int32)
        kv_cache.clear()
        for hk in hooks:
            hk.remove()

        if not found_target:
            seek_sample += adjusted_chunk_size if audio_segment.shape[-1] == CHUNK_SAMPLES else int(audio_segment.shape[-1])
            return None
        final_tokens = [p["token"] for p in predictions]
        if mode == 1:
            _, (ws, wts), _ = split_word_tokens([dict(tokens=final_tokens)], tk)
            final_probs = [p["prob"] for p in predictions]
            wps = [float(np.mean([final_probs.pop(0) for _ in wt])) for wt in wts]
            words = [dict(word=w, tokens=wt, probability=wp) for w, wt, wp in zip(ws, wts, wps)]
            seek_sample += round(curr_end * SAMPLE_RATE)
            return dict(end=target_end + seek, text=text, duration_window_text="".join(ws), duration_window_word=words)
        segment = dict(seek=0, tokens=final_tokens)
        add_word_timestamps_stable(segments=[segment], model=model, tokenizer=tk, enc=enc,
                                   num_samples=round(curr_end * SAMPLE_RATE), gap_padding=None)
        words = [dict(w, start=round(w["start"] + seek, 3), end=round(w["end"] + seek, 3)) for w in segment["words"]]
        seek_sample += round(segment["words"][-1]["end"] * SAMPLE_RATE)
        seg = Segment(words=words)
        seg.seek = curr_start
        return seg

    matches = []
    while seek_sample < total_samples and (not count or state["found"] < count):
        before = seek_sample
        m = one_chunk()
        if m is not None:
            matches.append(m)
        if seek_sample <= before:                             # a match ending at 0.0 s would never advance
            seek_sample = before + 1 if m is None else max(seek_sample, before + 1)
    if verbose and not matches:
        print(f'Failed to locate "{text}".')
    return matches
