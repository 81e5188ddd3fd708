"""Pins the oracle's restatement of openai-whisper against INDEPENDENT ports of the same algorithms that ship in
HF transformers, and the plain-C restatement against the Python one.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle.whisper_ref import audio as A
from oracle.whisper_ref import timing as T
from oracle.whisper_ref.model import ModelDimensions, Whisper, disable_sdpa, init_random_


def test_mel_filterbank_matches_hf_slaney():
    from transformers.audio_utils import mel_filter_bank
    for n_mels in (80, 128):
        ref = mel_filter_bank(201, n_mels, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T
        mine = A.mel_filterbank_np(n_mels)
        assert mine.shape == (n_mels, 201)
        np.testing.assert_allclose(mine, ref.astype(np.float32), rtol=0, atol=2e-7)


def test_log_mel_matches_hf_feature_extractor():
    from transformers import WhisperFeatureExtractor
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(480000, generator=g) * 2 - 1) * 0.1
    fe = WhisperFeatureExtractor(feature_size=80)
    ref = fe(x.numpy(), sampling_rate=16000, return_tensors="np")["input_features"][0]
    mine = A.log_mel_spectrogram(x, 80).numpy()
    assert mine.shape == (80, 3000)
    np.testing.assert_allclose(mine, ref, atol=2e-4, rtol=0)


@pytest.mark.parametrize("shape", [(3, 17, 40), (2, 5, 4), (1, 9, 3)])
def test_median_filter_matches_hf_port_and_c(shape):
    from transformers.models.whisper.generation_whisper import _median_filter
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g)
    mine = T.median_filter(x, 7)
    if shape[-1] > 3:
        ref = _median_filter(x, 7)
        assert torch.equal(mine, ref)
    else:
        assert torch.equal(mine, x)
    c = c_oracle.median_filter(x.numpy(), 7)
    assert np.array_equal(c, mine.numpy())


@pytest.mark.parametrize("N,M,quant", [(13, 50, False), (41, 333, False), (30, 200, True), (1, 7, False),
                                       (9, 1, False), (101, 1500, False)])
def test_dtw_matches_hf_port_and_c(N, M, quant):
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping
    rng = np.random.default_rng(N * 1000 + M)
    x = rng.standard_normal((N, M)).astype(np.float32)
    if quant:                                   # many exact ties
        x = np.round(x * 2) / 2
    mine = T.dtw(torch.from_numpy(x))
    ti, tj = _dynamic_time_warping(x.astype(np.float64))
    assert np.array_equal(mine[0], ti) and np.array_equal(mine[1], tj)
    path, jumps = c_oracle.dtw(x)
    assert np.array_equal(path, mine)
    jm = np.pad(np.diff(mine[0]), (1, 0), constant_values=1).astype(bool)
    assert np.array_equal(jumps, mine[1][jm].clip(min=0))


def test_dtw_nan_rows_follow_ieee_comparisons():
    x = np.random.default_rng(3).standard_normal((6, 20)).astype(np.float32)
    x[2] = np.nan
    mine = T.dtw(torch.from_numpy(x))
    path, _ = c_oracle.dtw(x)
    assert np.array_equal(path, mine)


def _hf_from_oracle(model: Whisper):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    d = model.dims
    cfg = WhisperConfig(vocab_size=d.n_vocab, num_mel_bins=d.n_mels, encoder_layers=d.n_audio_layer,
                        encoder_attention_heads=d.n_audio_head, decoder_layers=d.n_text_layer,
                        decoder_attention_heads=d.n_text_head, decoder_ffn_dim=4 * d.n_text_state,
                        encoder_ffn_dim=4 * d.n_audio_state, d_model=d.n_audio_state,
                        max_source_positions=d.n_audio_ctx, max_target_positions=d.n_text_ctx,
                        attn_implementation="eager", activation_function="gelu", pad_token_id=0, bos_token_id=1,
                        eos_token_id=2, decoder_start_token_id=1, suppress_tokens=None, begin_suppress_tokens=None)
    hf = WhisperForConditionalGeneration(cfg).eval()
    ren = [("blocks", "layers"), ("mlp.0", "fc1"), ("mlp.2", "fc2"), ("mlp_ln", "final_layer_norm"),
           (".attn.query", ".self_attn.q_proj"), (".attn.key", ".self_attn.k_proj"),
           (".attn.value", ".self_attn.v_proj"), (".attn_ln", ".self_attn_layer_norm"),
           (".attn.out", ".self_attn.out_proj"), (".cross_attn.query", ".encoder_attn.q_proj"),
           (".cross_attn.key", ".encoder_attn.k_proj"), (".cross_attn.value", ".encoder_attn.v_proj"),
           (".cross_attn_ln", ".encoder_attn_layer_norm"), (".cross_attn.out", ".encoder_attn.out_proj"),
           ("decoder.ln.", "decoder.layer_norm."), ("encoder.ln_post.", "encoder.layer_norm."),
           ("token_embedding", "embed_tokens"),
           ("encoder.positional_embedding", "encoder.embed_positions.weight"),
           ("decoder.positional_embedding", "decoder.embed_positions.weight")]
    sd = {}
    for k, v in model.state_dict().items():
        for a, b in ren:
            k = k.replace(a, b)
        sd["model." + k] = v
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in m for m in missing) or not missing, missing
    return hf


def test_model_forward_matches_hf_whisper():
    """Encoder output, logits and (softmaxed) cross-attention of the restated model == HF's independent port."""
    dims = ModelDimensions(80, 1500, 64, 2, 2, 51864, 448, 64, 2, 2)
    model = init_random_(Whisper(dims), seed=3).eval()
    hf = _hf_from_oracle(model)
    g = torch.Generator().manual_seed(0)
    mel = torch.randn(1, 80, 3000, generator=g) * 0.5
    toks = torch.randint(0, 50000, (1, 23), generator=g)
    with torch.no_grad(), disable_sdpa():
        qks = []
        hooks = [b.cross_attn.register_forward_hook(lambda m, i, o: qks.append(o[-1])) for b in model.decoder.blocks]
        xa = model.encoder(mel)
        logits = model.decoder(toks, xa)
        [h.remove() for h in hooks]
        out = hf(input_features=mel, decoder_input_ids=toks, output_attentions=True)
    np.testing.assert_allclose(xa.numpy(), out.encoder_last_hidden_state.numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(logits.numpy(), out.logits.numpy(), atol=2e-5, rtol=1e-5)
    for l in range(2):
        np.testing.assert_allclose(qks[l].softmax(-1).numpy(), out.cross_attentions[l].numpy(), atol=1e-6, rtol=1e-4)


def test_sdpa_and_explicit_attention_agree():
    dims = ModelDimensions(80, 1500, 64, 2, 1, 51864, 448, 64, 2, 1)
    model = init_random_(Whisper(dims), seed=4).eval()
    g = torch.Generator().manual_seed(1)
    mel = torch.randn(1, 80, 3000, generator=g)
    toks = torch.randint(0, 50000, (1, 9), generator=g)
    with torch.no_grad():
        a = model(mel, toks)
        with disable_sdpa():
            b = model(mel, toks)
    np.testing.assert_allclose(a.numpy(), b.numpy(), atol=1e-5, rtol=1e-5)
