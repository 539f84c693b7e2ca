"""Pin the oracle: block-by-block agreement with the HF PyTorch twins that import here
(transformers 5.5: modeling_clip.py, modeling_mbart.py, modeling_bart.py, modeling_vit.py) — SURVEY.md §8c."""
import math

import numpy as np
import pytest
import torch

import mic_b200
from mic_b200 import synthetic
from oracle import reference_model as rm
from oracle import reference_generate as rg

torch.manual_seed(0)


def _load_linear(lin, p):
    lin.weight.data = torch.from_numpy(p["kernel"]).t().contiguous()
    lin.bias.data = torch.from_numpy(p["bias"]).clone()


def _load_ln(ln, p):
    ln.weight.data = torch.from_numpy(p["scale"]).clone()
    ln.bias.data = torch.from_numpy(p["bias"]).clone()


def _hf_clip(cfg, params):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    c = cfg.clip_vision_config
    hc = CLIPVisionConfig(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                          num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                          image_size=c.image_size, patch_size=c.patch_size, hidden_act=c.hidden_act,
                          layer_norm_eps=c.layer_norm_eps)
    m = CLIPVisionModel(hc).eval()
    vp = params["model"]["encoder"]["vision_model"]
    vm = m.vision_model
    vm.embeddings.class_embedding.data = torch.from_numpy(vp["embeddings"]["class_embedding"]).clone()
    # Flax HWIO (kh,kw,cin,cout) -> torch OIHW
    vm.embeddings.patch_embedding.weight.data = torch.from_numpy(
        vp["embeddings"]["patch_embedding"]["kernel"]).permute(3, 2, 0, 1).contiguous()
    vm.embeddings.position_embedding.weight.data = torch.from_numpy(
        vp["embeddings"]["position_embedding"]["embedding"]).clone()
    _load_ln(vm.pre_layrnorm, vp["pre_layrnorm"])
    _load_ln(vm.post_layernorm, vp["post_layernorm"])
    for i, layer in enumerate(vm.encoder.layers):
        lp = vp["encoder"]["layers"][str(i)]
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _load_linear(getattr(layer.self_attn, n), lp["self_attn"][n])
        _load_ln(layer.layer_norm1, lp["layer_norm1"])
        _load_ln(layer.layer_norm2, lp["layer_norm2"])
        _load_linear(layer.mlp.fc1, lp["mlp"]["fc1"])
        _load_linear(layer.mlp.fc2, lp["mlp"]["fc2"])
    return m


def _hf_mbart_decoder(cfg, params):
    from transformers import MBartConfig
    from transformers.models.mbart.modeling_mbart import MBartDecoder
    t = cfg.mbart_config
    hc = MBartConfig(vocab_size=t.vocab_size, d_model=t.d_model, decoder_layers=t.decoder_layers,
                     decoder_attention_heads=t.decoder_attention_heads, decoder_ffn_dim=t.decoder_ffn_dim,
                     activation_function=t.activation_function, scale_embedding=t.scale_embedding,
                     max_position_embeddings=t.max_position_embeddings, dropout=0.0, pad_token_id=t.pad_token_id)
    d = MBartDecoder(hc).eval()
    dp = params["model"]["decoder"]
    d.embed_tokens.weight.data = torch.from_numpy(params["model"]["shared"]["embedding"]).clone()
    d.embed_positions.weight.data = torch.from_numpy(dp["embed_positions"]["embedding"]).clone()
    _load_ln(d.layernorm_embedding, dp["layernorm_embedding"])
    _load_ln(d.layer_norm, dp["layer_norm"])
    for i, layer in enumerate(d.layers):
        lp = dp["layers"][str(i)]
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _load_linear(getattr(layer.self_attn, n), lp["self_attn"][n])
            _load_linear(getattr(layer.encoder_attn, n), lp["encoder_attn"][n])
        _load_ln(layer.self_attn_layer_norm, lp["self_attn_layer_norm"])
        _load_ln(layer.encoder_attn_layer_norm, lp["encoder_attn_layer_norm"])
        _load_ln(layer.final_layer_norm, lp["final_layer_norm"])
        _load_linear(layer.fc1, lp["fc1"])
        _load_linear(layer.fc2, lp["fc2"])
    return d


@pytest.fixture(scope="module")
def setup():
    cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
    cfg.mbart_config.layer_norm_eps = 1e-5          # the PyTorch twin hard-codes nn.LayerNorm default
    params = synthetic.make_params(cfg, seed=1, perturbed=True, std=0.1)
    batch = synthetic.make_batch(cfg, batch_size=3, seq_len=16, seed=0, min_len=4)
    return cfg, params, batch


def test_clip_encoder_matches_hf(setup):
    cfg, params, batch = setup
    m = _hf_clip(cfg, params)
    with torch.no_grad():
        ref = m(pixel_values=torch.from_numpy(batch["pixel_values"]).permute(0, 3, 1, 2)).last_hidden_state
        got = rm.vision_encoder(rm.to_torch_tree(params), batch["pixel_values"], cfg)
    assert got.shape == ref.shape == (3, 5, 128)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), float((got - ref).abs().max())


def test_mbart_decoder_matches_hf(setup):
    cfg, params, batch = setup
    d = _hf_mbart_decoder(cfg, params)
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        enc = rm.encode(p, batch["pixel_values"], cfg)
        ids = torch.from_numpy(batch["decoder_input_ids"])
        # full (all ones) mask: the twin derives positions itself (arange), same as the reference default
        ref = d(input_ids=ids, attention_mask=torch.ones_like(ids), encoder_hidden_states=enc).last_hidden_state
        got = rm.decoder(p, ids, None, None, enc, cfg)
    assert torch.allclose(got, ref, atol=5e-5, rtol=1e-4), float((got - ref).abs().max())


def test_mbart_decoder_padding_mask_matches_hf(setup):
    """Key-padding semantics (the label mask of main.py:692): compare on rows whose query is not padded
    (HF uses finfo.min, the reference -inf; identical unless a row is fully masked, which cannot
    happen because key 0 is always kept)."""
    cfg, params, batch = setup
    d = _hf_mbart_decoder(cfg, params)
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        enc = rm.encode(p, batch["pixel_values"], cfg)
        ids = torch.from_numpy(batch["decoder_input_ids"])
        mask = torch.from_numpy(batch["attention_mask"])
        ref = d(input_ids=ids, attention_mask=mask, encoder_hidden_states=enc).last_hidden_state
        got = rm.decoder(p, ids, mask, None, enc, cfg)
    assert torch.allclose(got, ref, atol=5e-5, rtol=1e-4), float((got - ref).abs().max())


def test_cached_decode_equals_full_prefix(setup):
    """the cached 1-token path must equal the full pass on the prefix, last row."""
    cfg, params, batch = setup
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        enc = rm.encode(p, batch["pixel_values"], cfg)
        ids = batch["decoder_input_ids"][:, :7]
        full = rm.lm_head(p, rm.decoder(p, ids, None, None, enc, cfg)).numpy()
        cache = rg.DecodeCache(cfg.mbart_config.decoder_layers)
        for t in range(7):
            step = rg.decode_step(p, ids[:, t], t, enc, cache, cfg)
            np.testing.assert_allclose(step, full[:, t], atol=3e-5, rtol=1e-4)


def test_loss_matches_torch_cross_entropy(setup):
    cfg, params, batch = setup
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        logits = rm.forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"],
                                   batch["attention_mask"], None, cfg)
    labels = torch.from_numpy(batch["input_ids"])
    m = torch.from_numpy(batch["attention_mask"]).float()
    for eps in (0.0, 0.1):
        got = rm.loss_fn(logits, labels, m, eps)
        V = logits.shape[-1]
        lsm = torch.log_softmax(logits, -1)
        nll = -lsm.gather(-1, labels[..., None])[..., 0]
        if eps == 0.0:
            want = (nll * m).sum() / m.sum()
        else:
            low = eps / (V - 1)
            conf = 1 - eps
            const = -(conf * math.log(conf) + (V - 1) * low * math.log(low + 1e-20))
            per = conf * nll + low * (-(lsm.sum(-1)) - nll) - const
            want = (per * m).sum() / m.sum()
        assert abs(float(got) - float(want)) < 1e-5
