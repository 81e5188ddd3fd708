"""CPU oracle: ``whisper.tokenizer``-shaped tokenizer (openai-whisper 20250625).  TEST INFRASTRUCTURE.

The special-token LAYOUT (ids of sot/eot/language/task/no_speech/no_timestamps/timestamps, ``sot_sequence``) is
restated exactly; that is all the hot path depends on (stable_whisper/timing.py:230-237, decode.py:42-53,
alignment.py:650-657).  The BPE text vocabulary (tiktoken ranks shipped as ``assets/*.tiktoken``) is not available
offline, so text<->id uses a synthetic, deterministic, invertible vocabulary over the same id range:

  id < 256          -> the single byte ``id``
  256 <= id < eot   -> " t<id>" (starts a word)  or  "s<id>" when id % 4 == 0 (continues a word)

Real vocabularies plug in through ``ranks=`` (a tiktoken rank dict) when the asset files are present.
"""
import re
import string
from dataclasses import dataclass, field
from functools import cached_property
from typing import Dict, List, Optional, Tuple

LANGUAGE_CODES = [
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi",
    "vi", "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml",
    "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs",
    "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am",
    "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln",
    "ha", "ba", "jw", "su", "yue",
]

# dict forms the reference imports (stable_whisper/whisper_compatibility.py:64); names are not needed on the hot path
LANGUAGES = {c: ("english" if c == "en" else c) for c in LANGUAGE_CODES}
TO_LANGUAGE_CODE = {v: k for k, v in LANGUAGES.items()}

_PIECE = re.compile(r" t(\d+)|s(\d+)")


@dataclass
class Tokenizer:
    multilingual: bool
    num_languages: int = 99
    language: Optional[str] = None
    task: Optional[str] = None
    special_tokens: Dict[str, int] = field(default_factory=dict)
    sot_sequence: Tuple[int, ...] = ()

    def __post_init__(self):
        self.n_base = 50257 if self.multilingual else 50256
        specials = [
            "<|endoftext|>", "<|startoftranscript|>",
            *[f"<|{lang}|>" for lang in LANGUAGE_CODES[: self.num_languages]],
            "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
            "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)],
        ]
        for k, name in enumerate(specials):
            self.special_tokens[name] = self.n_base + k
        self._special_by_id = {v: k for k, v in self.special_tokens.items()}
        sot = self.special_tokens["<|startoftranscript|>"]
        seq = [sot]
        if self.language is not None:
            seq.append(sot + 1 + LANGUAGE_CODES.index(self.language))
        if self.task is not None:
            seq.append(self.special_tokens["<|transcribe|>" if self.task == "transcribe" else "<|translate|>"])
        self.sot_sequence = tuple(seq)
        self.language_code = self.language

    # ---- text <-> ids (synthetic vocabulary) ----
    def _piece(self, i: int) -> bytes:
        if i < 256:
            return bytes([i])
        return (f"s{i}" if i % 4 == 0 else f" t{i}").encode()

    def encode(self, text: str, **kwargs) -> List[int]:
        out, pos = [], 0
        for m in _PIECE.finditer(text):
            i = int(m.group(1) or m.group(2))
            ok = 256 <= i < self.n_base and self._piece(i).decode() == m.group(0)
            if not ok:
                continue
            out.extend(text[pos:m.start()].encode("utf-8"))
            out.append(i)
            pos = m.end()
        out.extend(text[pos:].encode("utf-8"))
        return out

    def decode(self, token_ids, **kwargs) -> str:
        token_ids = [int(t) for t in token_ids if int(t) < self.timestamp_begin]
        return self._decode_all(token_ids)

    def _decode_all(self, token_ids) -> str:
        buf = b""
        for t in token_ids:
            t = int(t)
            buf += self._piece(t) if t < self.n_base else self._special_by_id[t].encode()
        return buf.decode("utf-8", errors="replace")

    def decode_with_timestamps(self, token_ids, **kwargs) -> str:
        return self._decode_all(token_ids)

    # ---- special ids ----
    @cached_property
    def eot(self): return self.special_tokens["<|endoftext|>"]
    @cached_property
    def transcribe(self): return self.special_tokens["<|transcribe|>"]
    @cached_property
    def translate(self): return self.special_tokens["<|translate|>"]
    @cached_property
    def sot(self): return self.special_tokens["<|startoftranscript|>"]
    @cached_property
    def sot_lm(self): return self.special_tokens["<|startoflm|>"]
    @cached_property
    def sot_prev(self): return self.special_tokens["<|startofprev|>"]
    @cached_property
    def no_speech(self): return self.special_tokens["<|nospeech|>"]
    @cached_property
    def no_timestamps(self): return self.special_tokens["<|notimestamps|>"]
    @cached_property
    def timestamp_begin(self): return self.special_tokens["<|0.00|>"]

    @cached_property
    def language_token(self) -> int:
        if self.language is None:
            raise ValueError("This tokenizer does not have language token configured")
        return self.to_language_token(self.language)

    def to_language_token(self, language):
        if (tok := self.special_tokens.get(f"<|{language}|>")) is not None:
            return tok
        raise KeyError(f"Language {language} not found in tokenizer.")

    @cached_property
    def all_language_tokens(self) -> Tuple[int, ...]:
        sot = self.sot
        return tuple(range(sot + 1, sot + 1 + self.num_languages))

    @cached_property
    def all_language_codes(self) -> Tuple[str, ...]:
        return tuple(LANGUAGE_CODES[: self.num_languages])

    @cached_property
    def sot_sequence_including_notimestamps(self) -> Tuple[int, ...]:
        return tuple(list(self.sot_sequence) + [self.no_timestamps])

    @cached_property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        symbols = '"#()*+/:;<=>@[\\]^_`{|}~'
        return tuple(sorted({ord(c) for c in symbols} | {ord("-"), ord("'")}))

    # ---- word splitting (whisper.tokenizer.Tokenizer.split_to_word_tokens) ----
    def split_to_word_tokens(self, tokens: List[int]):
        if self.language in {"zh", "ja", "th", "lo", "my", "yue"}:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]):
        decoded_full = self.decode_with_timestamps(tokens)
        replacement_char = "�"
        words, word_tokens, current_tokens, unicode_offset = [], [], [], 0
        for token in tokens:
            current_tokens.append(token)
            decoded = self.decode_with_timestamps(current_tokens)
            if (replacement_char not in decoded
                    or decoded_full[unicode_offset + decoded.index(replacement_char)] == replacement_char):
                words.append(decoded)
                word_tokens.append(current_tokens)
                current_tokens = []
                unicode_offset += len(decoded)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: List[int]):
        subwords, subword_tokens_list = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for subword, subword_tokens in zip(subwords, subword_tokens_list):
            special = subword_tokens[0] >= self.eot
            with_space = subword.startswith(" ")
            punctuation = subword.strip() in string.punctuation
            if special or with_space or punctuation or len(words) == 0:
                words.append(subword)
                word_tokens.append(subword_tokens)
            else:
                words[-1] = words[-1] + subword
                word_tokens[-1].extend(subword_tokens)
        return words, word_tokens


def get_tokenizer(multilingual: bool, *, num_languages: int = 99, language: Optional[str] = None,
                  task: Optional[str] = None) -> Tokenizer:
    if language is not None:
        language = language.lower()
        if language not in LANGUAGE_CODES:
            raise ValueError(f"Unsupported language: {language}")
    if multilingual:
        language = language or "en"
        task = task or "transcribe"
    else:
        language = None
        task = None
    return Tokenizer(multilingual=multilingual, num_languages=num_languages, language=language, task=task)
