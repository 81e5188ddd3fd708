"""Tokenizer for the B200 path: Whisper's special-token layout + a pluggable text vocabulary.

The hot path only needs the special ids (``sot_sequence``, ``no_timestamps``, ``eot``, ``timestamp_begin``...;
stable_whisper/timing.py:230-237, decode.py:42-53).  Text <-> ids comes from, in order of preference:
  1. the real ``whisper.tokenizer`` if openai-whisper is importable (what the reference itself uses,
     stable_whisper/whisper_compatibility.py:310-335);
  2. a tiktoken rank file (``multilingual.tiktoken`` / ``gpt2.tiktoken`` from openai-whisper's assets) given by path;
  3. ``synthetic=True``: a deterministic invertible stand-in vocabulary over the same id range, for random-weight
     benchmarks and tests (no vocabulary files exist offline).
"""
import base64
import os
import re
import string
from functools import cached_property
from typing import Dict, List, Optional, Tuple

LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy "
    "sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo "
    "uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue").split()

_PIECE = re.compile(r" t(\d+)|s(\d+)")


class _SyntheticCodec:
    """id < 256 -> that byte; otherwise ' t<id>' (starts a word) or 's<id>' when id % 4 == 0 (continues one)."""

    def __init__(self, n_base: int):
        self.n_base = n_base

    def piece(self, i: int) -> bytes:
        return bytes([i]) if i < 256 else (f"s{i}" if i % 4 == 0 else f" t{i}").encode()

    def encode(self, text: str) -> List[int]:
        out, pos = [], 0
        for m in _PIECE.finditer(text):
            i = int(m.group(1) or m.group(2))
            if not (256 <= i < self.n_base and self.piece(i).decode() == m.group(0)):
                continue
            out.extend(text[pos:m.start()].encode("utf-8"))
            out.append(i)
            pos = m.end()
        out.extend(text[pos:].encode("utf-8"))
        return out


class _TiktokenCodec:
    def __init__(self, path: str, specials: Dict[str, int]):
        import tiktoken
        ranks = {base64.b64decode(tok): int(rank) for tok, rank in (line.split() for line in open(path) if line)}
        self.n_base = len(ranks)
        self.enc = tiktoken.Encoding(
            name=os.path.basename(path), explicit_n_vocab=self.n_base + len(specials),
            pat_str=r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
            mergeable_ranks=ranks, special_tokens=specials)

    def piece(self, i: int) -> bytes:
        return self.enc.decode_single_token_bytes(i)

    def encode(self, text: str) -> List[int]:
        return self.enc.encode(text)


class Tokenizer:
    def __init__(self, multilingual: bool, num_languages: int = 99, language: Optional[str] = None,
                 task: Optional[str] = None, vocab_path: Optional[str] = None):
        self.multilingual, self.num_languages = multilingual, num_languages
        self.language, self.task = language, task
        self.language_code = language
        n_base = 50257 if multilingual else 50256
        names = ["<|endoftext|>", "<|startoftranscript|>", *[f"<|{c}|>" for c in LANGUAGE_CODES[:num_languages]],
                 "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
                 "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)]]
        self.special_tokens = {n: n_base + k for k, n in enumerate(names)}
        self._special_by_id = {v: k for k, v in self.special_tokens.items()}
        self.codec = _TiktokenCodec(vocab_path, self.special_tokens) if vocab_path else _SyntheticCodec(n_base)
        assert self.codec.n_base == n_base, f"vocabulary has {self.codec.n_base} entries, expected {n_base}"
        seq = [self.sot]
        if language is not None:
            seq.append(self.sot + 1 + LANGUAGE_CODES.index(language))
        if task is not None:
            seq.append(self.transcribe if task == "transcribe" else self.translate)
        self.sot_sequence: Tuple[int, ...] = tuple(seq)

    def _tok(self, name):
        return self.special_tokens[name]

    eot = cached_property(lambda self: self._tok("<|endoftext|>"))
    sot = cached_property(lambda self: self._tok("<|startoftranscript|>"))
    transcribe = cached_property(lambda self: self._tok("<|transcribe|>"))
    translate = cached_property(lambda self: self._tok("<|translate|>"))
    sot_lm = cached_property(lambda self: self._tok("<|startoflm|>"))
    sot_prev = cached_property(lambda self: self._tok("<|startofprev|>"))
    no_speech = cached_property(lambda self: self._tok("<|nospeech|>"))
    no_timestamps = cached_property(lambda self: self._tok("<|notimestamps|>"))
    timestamp_begin = cached_property(lambda self: self._tok("<|0.00|>"))

    @cached_property
    def sot_sequence_including_notimestamps(self):
        return tuple(list(self.sot_sequence) + [self.no_timestamps])

    @cached_property
    def all_language_tokens(self):
        return tuple(range(self.sot + 1, self.sot + 1 + self.num_languages))

    @cached_property
    def all_language_codes(self):
        return tuple(LANGUAGE_CODES[: self.num_languages])

    @cached_property
    def non_speech_tokens(self):
        if isinstance(self.codec, _SyntheticCodec):
            # the stand-in vocabulary has no multi-character pieces: the set is the single-byte ids of whisper's ASCII symbol
            # list plus "-" and "'" (the same definition as the oracle's stand-in tokenizer, oracle/whisper_ref/tokenizer.py)
            return tuple(sorted({ord(c) for c in '"#()*+/:;<=>@[\\]^_`{|}~'} | {ord("-"), ord("'")}))
        symbols = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』') + "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        misc = set("♩♪♫♬♭♮♯")
        result = {self.encode(" -")[0], self.encode(" '")[0]}
        for s in symbols + list(misc):
            for toks in (self.encode(s), self.encode(" " + s)):
                if len(toks) == 1 or s in misc:
                    result.add(toks[0])
        return tuple(sorted(result))

    def encode(self, text: str, **kw) -> List[int]:
        return self.codec.encode(text)

    _piece_cache = None     # id -> bytes, per instance (the host bookkeeping decodes every token several times per window)

    def _decode_all(self, ids) -> str:
        cache = self._piece_cache
        if cache is None:
            cache = self._piece_cache = {}
        out = []
        for t in ids:
            b = cache.get(t)
            if b is None:
                t = int(t)
                b = self.codec.piece(t) if t < self.codec.n_base else self._special_by_id[t].encode()
                cache[t] = b
            out.append(b)
        return b"".join(out).decode("utf-8", errors="replace")

    def decode(self, ids, **kw) -> str:
        return self._decode_all([t for t in ids if int(t) < self.timestamp_begin])

    def decode_with_timestamps(self, ids, **kw) -> str:
        return self._decode_all(ids)

    def split_to_word_tokens(self, tokens: List[int]):
        if self.language in {"zh", "ja", "th", "lo", "my", "yue"}:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]):
        full = self.decode_with_timestamps(tokens)
        bad = "�"
        words, groups, cur, off = [], [], [], 0
        for t in tokens:
            cur.append(t)
            dec = self.decode_with_timestamps(cur)
            if bad not in dec or full[off + dec.index(bad)] == bad:
                words.append(dec)
                groups.append(cur)
                cur = []
                off += len(dec)
        return words, groups

    def split_tokens_on_spaces(self, tokens: List[int]):
        subwords, subgroups = self.split_tokens_on_unicode(tokens)
        words, groups = [], []
        for sw, sg in zip(subwords, subgroups):
            if sg[0] >= self.eot or sw.startswith(" ") or sw.strip() in string.punctuation or not words:
                words.append(sw)
                groups.append(sg)
            else:
                words[-1] += sw
                groups[-1].extend(sg)
        return words, groups


def get_tokenizer(model=None, *, multilingual: Optional[bool] = None, num_languages: Optional[int] = None,
                  language: Optional[str] = None, task: Optional[str] = None, vocab_path: Optional[str] = None,
                  synthetic: bool = False):
    """Same call shape as stable_whisper.whisper_compatibility.get_tokenizer (model first)."""
    if multilingual is None:
        multilingual = bool(model.is_multilingual)
    if num_languages is None:
        num_languages = int(model.num_languages) if model is not None else 99
    if multilingual:
        language, task = (language or "en").lower(), task or "transcribe"
        if language not in LANGUAGE_CODES:
            raise ValueError(f"Unsupported language: {language}")
    else:
        language = task = None
    if vocab_path is None and not synthetic:
        try:                                            # real openai-whisper tokenizer when it is installed
            from whisper.tokenizer import get_tokenizer as _real
            return _real(multilingual, num_languages=num_languages, language=language, task=task)
        except ImportError:
            vocab_path = os.environ.get("STB_TIKTOKEN_MULTILINGUAL" if multilingual else "STB_TIKTOKEN_GPT2")
            if not vocab_path:
                raise RuntimeError("no Whisper vocabulary available: install openai-whisper, pass vocab_path=<*.tiktoken>, "
                                   "or request the synthetic stand-in vocabulary with synthetic=True")
    return Tokenizer(multilingual, num_languages, language, task, vocab_path=vocab_path)
