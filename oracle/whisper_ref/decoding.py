"""CPU oracle: restatement of ``whisper.decoding`` (openai-whisper 20250625), greedy / sampling paths.

TEST INFRASTRUCTURE.  The reference subclasses ``DecodingTask`` (stable_whisper/decode.py:20-65) and overrides
``_main_loop``; everything that loop touches is restated here: ``inference.logits`` with KV-cache hooks,
``logit_filters`` (SuppressBlank / SuppressTokens / ApplyTimestampRules), ``decoder.update`` (argmax at T=0,
log-prob accumulation, EOT latching), ``run`` (initial tokens, no-speech prob, ranking, DecodingResult).
Beam search is not on the configured path (temperature=0, beam_size=None) and is not restated.
"""
import zlib
from dataclasses import dataclass, field, replace
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor
from torch.distributions import Categorical

from .audio import CHUNK_LENGTH
from .tokenizer import Tokenizer, get_tokenizer


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


@torch.no_grad()
def detect_language(model, mel: Tensor, tokenizer: Tokenizer = None):
    if tokenizer is None:
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages)
    if tokenizer.language is None or tokenizer.language_token not in tokenizer.sot_sequence:
        raise ValueError("This model doesn't have language tokens so it can't perform lang id")
    single = mel.ndim == 2
    if single:
        mel = mel.unsqueeze(0)
    if mel.shape[-2:] != (model.dims.n_audio_ctx, model.dims.n_audio_state):
        mel = model.encoder(mel)
    n_audio = mel.shape[0]
    x = torch.tensor([[tokenizer.sot]] * n_audio).to(mel.device)
    logits = model.logits(x, mel)[:, 0]
    mask = torch.ones(logits.shape[-1], dtype=torch.bool)
    mask[list(tokenizer.all_language_tokens)] = False
    logits[:, mask] = -np.inf
    language_tokens = logits.argmax(dim=-1)
    language_token_probs = logits.softmax(dim=-1).cpu()
    language_probs = [
        {c: language_token_probs[i, j].item() for j, c in zip(tokenizer.all_language_tokens,
                                                              tokenizer.all_language_codes)}
        for i in range(n_audio)
    ]
    if single:
        language_tokens = language_tokens[0]
        language_probs = language_probs[0]
    return language_tokens, language_probs


@dataclass(frozen=True)
class DecodingOptions:
    task: str = "transcribe"
    language: Optional[str] = None
    temperature: float = 0.0
    sample_len: Optional[int] = None
    best_of: Optional[int] = None
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[Union[str, List[int]]] = None
    prefix: Optional[Union[str, List[int]]] = None
    suppress_tokens: Optional[Union[str, Iterable[int]]] = "-1"
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    fp16: bool = True


@dataclass(frozen=True)
class DecodingResult:
    audio_features: Tensor
    language: str
    language_probs: Optional[Dict[str, float]] = None
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    temperature: float = np.nan
    compression_ratio: float = np.nan


class PyTorchInference:
    def __init__(self, model, initial_token_length: int):
        self.model = model
        self.initial_token_length = initial_token_length
        self.kv_cache = {}
        self.hooks = []
        self.kv_modules = ([b.attn.key for b in model.decoder.blocks] +
                           [b.attn.value for b in model.decoder.blocks])

    def logits(self, tokens: Tensor, audio_features: Tensor) -> Tensor:
        if not self.kv_cache:
            self.kv_cache, self.hooks = self.model.install_kv_cache_hooks()
        if tokens.shape[-1] > self.initial_token_length:
            tokens = tokens[:, -1:]                    # only the newest token after the first call
        return self.model.decoder(tokens, audio_features, kv_cache=self.kv_cache)

    def cleanup_caching(self):
        for h in self.hooks:
            h.remove()
        self.kv_cache = {}
        self.hooks = []

    def rearrange_kv_cache(self, source_indices):
        if source_indices != list(range(len(source_indices))):
            for m in self.kv_modules:
                self.kv_cache[m] = self.kv_cache[m][source_indices].detach()


class MaximumLikelihoodRanker:
    def __init__(self, length_penalty: Optional[float]):
        self.length_penalty = length_penalty

    def rank(self, tokens: List[List[Tensor]], sum_logprobs: List[List[float]]):
        def scores(logprobs, lengths):
            out = []
            for lp, n in zip(logprobs, lengths):
                pen = n if self.length_penalty is None else ((5 + n) / 6) ** self.length_penalty
                out.append(lp / pen)
            return out
        lengths = [[len(t) for t in s] for s in tokens]
        return [int(np.argmax(scores(p, l))) for p, l in zip(sum_logprobs, lengths)]


class GreedyDecoder:
    def __init__(self, temperature: float, eot: int):
        self.temperature = temperature
        self.eot = eot

    def reset(self):
        pass

    def update(self, tokens: Tensor, logits: Tensor, sum_logprobs: Tensor) -> Tuple[Tensor, bool]:
        if self.temperature == 0:
            next_tokens = logits.argmax(dim=-1)
        else:
            next_tokens = Categorical(logits=logits / self.temperature).sample()
        logprobs = F.log_softmax(logits.float(), dim=-1)
        current = logprobs[torch.arange(logprobs.shape[0]), next_tokens]
        sum_logprobs += current * (tokens[:, -1] != self.eot)
        next_tokens[tokens[:, -1] == self.eot] = self.eot
        tokens = torch.cat([tokens, next_tokens[:, None]], dim=-1)
        completed = (tokens[:, -1] == self.eot).all()
        return tokens, completed

    def finalize(self, tokens: Tensor, sum_logprobs: Tensor):
        tokens = F.pad(tokens, (0, 1), value=self.eot)       # every sequence has at least one EOT
        return tokens, sum_logprobs.tolist()


class LogitFilter:
    def apply(self, logits: Tensor, tokens: Tensor) -> None:
        raise NotImplementedError


class SuppressBlank(LogitFilter):
    def __init__(self, tokenizer: Tokenizer, sample_begin: int):
        self.tokenizer = tokenizer
        self.sample_begin = sample_begin

    def apply(self, logits, tokens):
        if tokens.shape[1] == self.sample_begin:
            logits[:, self.tokenizer.encode(" ") + [self.tokenizer.eot]] = -np.inf


class SuppressTokens(LogitFilter):
    def __init__(self, suppress_tokens: Sequence[int]):
        self.suppress_tokens = list(suppress_tokens)

    def apply(self, logits, tokens):
        logits[:, self.suppress_tokens] = -np.inf


class ApplyTimestampRules(LogitFilter):
    def __init__(self, tokenizer: Tokenizer, sample_begin: int, max_initial_timestamp_index: Optional[int]):
        self.tokenizer = tokenizer
        self.sample_begin = sample_begin
        self.max_initial_timestamp_index = max_initial_timestamp_index

    def apply(self, logits, tokens):
        tk = self.tokenizer
        if tk.no_timestamps is not None:
            logits[:, tk.no_timestamps] = -np.inf
        for k in range(tokens.shape[0]):
            sampled = tokens[k, self.sample_begin:]
            seq = sampled.tolist()
            last_was_ts = len(seq) >= 1 and seq[-1] >= tk.timestamp_begin
            penult_was_ts = len(seq) < 2 or seq[-2] >= tk.timestamp_begin
            if last_was_ts:
                if penult_was_ts:                       # pair complete: next must be text
                    logits[k, tk.timestamp_begin:] = -np.inf
                else:                                   # open pair: next must be a timestamp/EOT
                    logits[k, : tk.eot] = -np.inf
            timestamps = sampled[sampled.ge(tk.timestamp_begin)]
            if timestamps.numel() > 0:
                if last_was_ts and not penult_was_ts:
                    ts_last = timestamps[-1]
                else:
                    ts_last = timestamps[-1] + 1        # force non-zero segment length
                logits[k, tk.timestamp_begin: ts_last] = -np.inf
        if tokens.shape[1] == self.sample_begin:
            logits[:, : tk.timestamp_begin] = -np.inf
            if self.max_initial_timestamp_index is not None:
                last_allowed = tk.timestamp_begin + self.max_initial_timestamp_index
                logits[:, last_allowed + 1:] = -np.inf
        logprobs = F.log_softmax(logits.float(), dim=-1)
        for k in range(tokens.shape[0]):
            ts_lp = logprobs[k, tk.timestamp_begin:].logsumexp(dim=-1)
            max_text_lp = logprobs[k, : tk.timestamp_begin].max()
            if ts_lp > max_text_lp:
                logits[k, : tk.timestamp_begin] = -np.inf


class DecodingTask:
    def __init__(self, model, options: DecodingOptions):
        self.model = model
        language = options.language or "en"
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language,
                                  task=options.task)
        self.tokenizer: Tokenizer = tokenizer
        self.options: DecodingOptions = self._verify_options(options)
        self.n_group: int = options.beam_size or options.best_of or 1
        self.n_ctx: int = model.dims.n_text_ctx
        self.sample_len: int = options.sample_len or model.dims.n_text_ctx // 2
        self.sot_sequence = tokenizer.sot_sequence
        if self.options.without_timestamps:
            self.sot_sequence = tokenizer.sot_sequence_including_notimestamps
        self.initial_tokens: Tuple[int, ...] = self._get_initial_tokens()
        self.sample_begin: int = len(self.initial_tokens)
        self.sot_index: int = self.initial_tokens.index(tokenizer.sot)
        self.inference = PyTorchInference(model, len(self.initial_tokens))
        self.sequence_ranker = MaximumLikelihoodRanker(options.length_penalty)
        if options.beam_size is not None:
            raise NotImplementedError("beam search is outside the restated path")
        self.decoder = GreedyDecoder(options.temperature, tokenizer.eot)
        self.logit_filters: List[LogitFilter] = []
        if self.options.suppress_blank:
            self.logit_filters.append(SuppressBlank(self.tokenizer, self.sample_begin))
        if self.options.suppress_tokens:
            self.logit_filters.append(SuppressTokens(self._get_suppress_tokens()))
        if not options.without_timestamps:
            precision = CHUNK_LENGTH / model.dims.n_audio_ctx
            max_initial_timestamp_index = None
            if options.max_initial_timestamp:
                max_initial_timestamp_index = round(self.options.max_initial_timestamp / precision)
            self.logit_filters.append(ApplyTimestampRules(tokenizer, self.sample_begin, max_initial_timestamp_index))

    def _verify_options(self, options: DecodingOptions) -> DecodingOptions:
        if options.beam_size is not None and options.best_of is not None:
            raise ValueError("beam_size and best_of can't be given together")
        if options.temperature == 0 and options.best_of is not None:
            raise ValueError("best_of with greedy sampling (T=0) is not compatible")
        if options.patience is not None and options.beam_size is None:
            raise ValueError("patience requires beam_size to be given")
        if options.length_penalty is not None and not (0 <= options.length_penalty <= 1):
            raise ValueError("length_penalty (alpha) should be a value between 0 and 1")
        return options

    def _get_initial_tokens(self) -> Tuple[int, ...]:
        tokens = list(self.sot_sequence)
        if prefix := self.options.prefix:
            ptoks = self.tokenizer.encode(" " + prefix.strip()) if isinstance(prefix, str) else prefix
            if self.sample_len is not None:
                ptoks = ptoks[-(self.n_ctx // 2 - self.sample_len):]
            tokens = tokens + ptoks
        if prompt := self.options.prompt:
            ptoks = self.tokenizer.encode(" " + prompt.strip()) if isinstance(prompt, str) else prompt
            tokens = [self.tokenizer.sot_prev] + ptoks[-(self.n_ctx // 2 - 1):] + tokens
        return tuple(tokens)

    def _get_suppress_tokens(self) -> Tuple[int, ...]:
        st = self.options.suppress_tokens
        if isinstance(st, str):
            st = [int(t) for t in st.split(",")]
        if -1 in st:
            st = [t for t in st if t >= 0]
            st.extend(self.tokenizer.non_speech_tokens)
        elif st is None or len(st) == 0:
            st = []
        else:
            assert isinstance(st, list), "suppress_tokens must be a list"
        tk = self.tokenizer
        st.extend([tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm])
        if tk.no_speech is not None:
            st.append(tk.no_speech)
        return tuple(sorted(set(st)))

    def _get_audio_features(self, mel: Tensor):
        if self.options.fp16:
            mel = mel.half()
        if mel.shape[-2:] == (self.model.dims.n_audio_ctx, self.model.dims.n_audio_state):
            audio_features = mel                          # encoded features passed in
        else:
            audio_features = self.model.encoder(mel)
        if audio_features.dtype != (torch.float16 if self.options.fp16 else torch.float32):
            raise TypeError(f"audio_features has an incorrect dtype: {audio_features.dtype}")
        return audio_features

    def _detect_language(self, audio_features: Tensor, tokens: Tensor):
        languages = [self.options.language] * audio_features.shape[0]
        lang_probs = None
        if self.options.language is None or self.options.task == "lang_id":
            lang_tokens, lang_probs = self.model.detect_language(audio_features, self.tokenizer)
            languages = [max(p, key=p.get) for p in lang_probs]
            if self.options.language is None:
                tokens[:, self.sot_index + 1] = lang_tokens
        return languages, lang_probs

    def _main_loop(self, audio_features: Tensor, tokens: Tensor):
        n_batch = tokens.shape[0]
        sum_logprobs = torch.zeros(n_batch, device=audio_features.device)
        no_speech_probs = [np.nan] * n_batch
        try:
            for i in range(self.sample_len):
                logits = self.inference.logits(tokens, audio_features)
                if i == 0 and self.tokenizer.no_speech is not None:
                    probs_at_sot = logits[:, self.sot_index].float().softmax(dim=-1)
                    no_speech_probs = probs_at_sot[:, self.tokenizer.no_speech].tolist()
                logits = logits[:, -1]
                for f in self.logit_filters:
                    f.apply(logits, tokens)
                tokens, completed = self.decoder.update(tokens, logits, sum_logprobs)
                if completed or tokens.shape[-1] > self.n_ctx:
                    break
        finally:
            self.inference.cleanup_caching()
        return tokens, sum_logprobs, no_speech_probs

    @torch.no_grad()
    def run(self, mel: Tensor) -> List[DecodingResult]:
        self.decoder.reset()
        tk = self.tokenizer
        n_audio = mel.shape[0]
        audio_features = self._get_audio_features(mel)
        tokens = torch.tensor([self.initial_tokens]).repeat(n_audio, 1)
        languages, language_probs = self._detect_language(audio_features, tokens)
        if self.options.task == "lang_id":
            return [DecodingResult(audio_features=f, language=l, language_probs=p)
                    for f, l, p in zip(audio_features, languages, language_probs)]
        tokens = tokens.repeat_interleave(self.n_group, dim=0).to(audio_features.device)
        tokens, sum_logprobs, no_speech_probs = self._main_loop(audio_features, tokens)
        audio_features = audio_features[:: self.n_group]
        no_speech_probs = no_speech_probs[:: self.n_group]
        assert audio_features.shape[0] == len(no_speech_probs) == n_audio
        tokens = tokens.reshape(n_audio, self.n_group, -1)
        sum_logprobs = sum_logprobs.reshape(n_audio, self.n_group)
        tokens, sum_logprobs = self.decoder.finalize(tokens, sum_logprobs)
        tokens = [[t[self.sample_begin: (t == tk.eot).nonzero()[0, 0]] for t in s] for s in tokens]
        selected = self.sequence_ranker.rank(tokens, sum_logprobs)
        tokens = [t[i].tolist() for i, t in zip(selected, tokens)]
        texts = [tk.decode(t).strip() for t in tokens]
        sum_logprobs = [lp[i] for i, lp in zip(selected, sum_logprobs)]
        avg_logprobs = [lp / (len(t) + 1) for t, lp in zip(tokens, sum_logprobs)]
        fields = (texts, languages, tokens, audio_features, avg_logprobs, no_speech_probs)
        if len(set(map(len, fields))) != 1:
            raise RuntimeError(f"inconsistent result lengths: {list(map(len, fields))}")
        return [
            DecodingResult(audio_features=f, language=l, tokens=t, text=x, avg_logprob=a, no_speech_prob=n,
                           temperature=self.options.temperature, compression_ratio=compression_ratio(x))
            for x, l, t, f, a, n in zip(*fields)
        ]


@torch.no_grad()
def decode(model, mel: Tensor, options: DecodingOptions = DecodingOptions(), **kwargs):
    if single := mel.ndim == 2:
        mel = mel.unsqueeze(0)
    if kwargs:
        options = replace(options, **kwargs)
    result = DecodingTask(model, options).run(mel)
    return result[0] if single else result
