"""Seeded synthetic weights and batches shared by tests, smoke() and bench.py.

Nothing here is model code: it only fabricates inputs of the shapes the reference feeds the hot
path.  Input contract = what `main.py:165-245,493-543` hands to `train_step`:
NHWC float32 pixels, int labels `[lang_code, tokens..., eos, pad...]`, `attention_mask = labels != pad`,
`decoder_input_ids = shift_tokens_right(labels, pad)` (`main.py:362-369`).

The parameter tree uses the Flax names the reference's checkpoints use.
"""
from __future__ import annotations

import numpy as np

LANG_CODES = (250004, 250008, 250005, 250003)  # en_XX fr_XX es_XX de_DE (main.py:201-206)


def shift_tokens_right(input_ids: np.ndarray, pad_token_id: int) -> np.ndarray:
    """`main.py:362-369`: shift one to the right, position 0 becomes the pad id."""
    out = np.zeros(input_ids.shape, dtype=np.int64)
    out[:, 1:] = input_ids[:, :-1]
    out[:, 0] = pad_token_id
    return out


def make_batch(config, batch_size: int, seq_len: int = 64, seed: int = 0, min_len: int = 8):
    """SURVEY.md §8d config 1 recipe."""
    rng = np.random.default_rng(seed)
    v, t = config.clip_vision_config, config.mbart_config
    if v.channel_first_input:
        pixel_values = rng.standard_normal((batch_size, 3, v.image_size, v.image_size), dtype=np.float32)
    else:
        pixel_values = rng.standard_normal((batch_size, v.image_size, v.image_size, 3), dtype=np.float32)
    labels = np.full((batch_size, seq_len), t.pad_token_id, dtype=np.int64)
    lo_tok = 4
    hi_tok = max(lo_tok + 1, min(t.vocab_size - 60, 250000))
    codes = [c for c in LANG_CODES if c < t.vocab_size] or [t.bos_token_id]
    for i in range(batch_size):
        n = int(rng.integers(min(min_len, seq_len), seq_len + 1))
        labels[i, 0] = codes[i % len(codes)]
        if n > 2:
            labels[i, 1:n - 1] = rng.integers(lo_tok, hi_tok, size=n - 2)
        labels[i, n - 1] = t.eos_token_id
    attention_mask = (labels != t.pad_token_id).astype(np.int64)
    decoder_input_ids = shift_tokens_right(labels, t.pad_token_id)
    return {"pixel_values": pixel_values, "input_ids": labels, "attention_mask": attention_mask,
            "decoder_input_ids": decoder_input_ids}


def _dense(rng, din, dout, std, perturbed):
    p = {"kernel": (rng.standard_normal((din, dout), dtype=np.float32) * std)}
    p["bias"] = (rng.standard_normal((dout,), dtype=np.float32) * std) if perturbed else np.zeros((dout,), np.float32)
    return p


def _ln(rng, d, perturbed):
    if perturbed:
        return {"scale": 1.0 + rng.standard_normal((d,), dtype=np.float32) * 0.02,
                "bias": rng.standard_normal((d,), dtype=np.float32) * 0.02}
    return {"scale": np.ones((d,), np.float32), "bias": np.zeros((d,), np.float32)}


def make_params(config, seed: int = 1, perturbed: bool = False, std: float | None = None):
    """Random-init parameter tree with the reference's names and shapes.

    perturbed=False: "init-like" (N(0, std^2) kernels/embeddings, zero biases, unit LayerNorm).
    perturbed=True : non-zero biases / LN scale+bias / final_logits_bias so bias bugs cannot hide.
    """
    rng = np.random.default_rng(seed)
    v, t = config.clip_vision_config, config.mbart_config
    std = t.init_std if std is None else std
    dv, d, p = v.hidden_size, t.d_model, v.patch_size

    def attn(dm):
        return {k: _dense(rng, dm, dm, std, perturbed) for k in ("q_proj", "k_proj", "v_proj", "out_proj")}

    vlayers = {}
    for i in range(v.num_hidden_layers):
        vlayers[str(i)] = {
            "self_attn": attn(dv), "layer_norm1": _ln(rng, dv, perturbed),
            "mlp": {"fc1": _dense(rng, dv, v.intermediate_size, std, perturbed),
                    "fc2": _dense(rng, v.intermediate_size, dv, std, perturbed)},
            "layer_norm2": _ln(rng, dv, perturbed)}
    patch = {"kernel": rng.standard_normal((p, p, 3, dv), dtype=np.float32) * std}
    if v.patch_bias:
        patch["bias"] = (rng.standard_normal((dv,), dtype=np.float32) * std) if perturbed else np.zeros((dv,), np.float32)
    vision = {
        "embeddings": {"class_embedding": rng.standard_normal((dv,), dtype=np.float32) * std,
                       "patch_embedding": patch,
                       "position_embedding": {"embedding": rng.standard_normal((v.num_tokens, dv), dtype=np.float32) * std}},
        "pre_layrnorm": _ln(rng, dv, perturbed),
        "encoder": {"layers": vlayers},
        "post_layernorm": _ln(rng, dv, perturbed)}
    dlayers = {}
    for i in range(t.decoder_layers):
        dlayers[str(i)] = {
            "self_attn": attn(d), "self_attn_layer_norm": _ln(rng, d, perturbed),
            "encoder_attn": attn(d), "encoder_attn_layer_norm": _ln(rng, d, perturbed),
            "fc1": _dense(rng, d, t.decoder_ffn_dim, std, perturbed),
            "fc2": _dense(rng, t.decoder_ffn_dim, d, std, perturbed),
            "final_layer_norm": _ln(rng, d, perturbed)}
    decoder = {
        "embed_positions": {"embedding": rng.standard_normal(
            (t.max_position_embeddings + t.position_offset, d), dtype=np.float32) * std},
        "layernorm_embedding": _ln(rng, d, perturbed),
        "layers": dlayers,
        "layer_norm": _ln(rng, d, perturbed)}
    shared = rng.standard_normal((t.vocab_size, d), dtype=np.float32) * std
    flb = (rng.standard_normal((1, t.vocab_size), dtype=np.float32) * 0.05) if perturbed \
        else np.zeros((1, t.vocab_size), np.float32)
    return {"model": {"encoder": {"vision_model": vision}, "decoder": decoder,
                      "shared": {"embedding": shared},
                      "visual_projection": _dense(rng, dv, d, std, perturbed)},
            "final_logits_bias": flb}


def tree_flatten(tree, prefix=()):
    """Yield (path_tuple, leaf) in deterministic (insertion) order."""
    for k, val in tree.items():
        if isinstance(val, dict):
            yield from tree_flatten(val, prefix + (k,))
        else:
            yield prefix + (k,), val


def tree_map(fn, tree):
    return {k: (tree_map(fn, v) if isinstance(v, dict) else fn(v)) for k, v in tree.items()}


def make_peaked_params(config, seed: int = 11, std: float = 0.03, out_scale: float = 5.3, emb_std: float = 0.04,
                       ln_emb_scale: float = 0.08, bias_std: float = 1.0, eos_bias: float = 7.6):
    """"Trained-weights-shaped" synthetic set for the full-size generation goldens.  Plain random init with tied
    embeddings just repeats its input token (the embedding of the fed token dominates the residual stream), so its
    goldens say little about attention / FFN / cache handling.  Here the decoder sublayers matter: their output
    matrices (out_proj, fc2) are scaled up, `layernorm_embedding` is scaled down, `final_logits_bias` is heavy
    (N(0, bias_std^2), EOS at `eos_bias`) — giving varied token sequences whose decisions mostly carry margins well
    above the bf16 noise, and EOS firing naturally at different lengths (finished-set merge, early stopping).
    Deterministic in `seed`; rebuilt on both sides of a golden test, never stored."""
    p = make_params(config, seed=seed, perturbed=True, std=std)
    t = config.mbart_config
    for lp in p["model"]["decoder"]["layers"].values():
        lp["self_attn"]["out_proj"]["kernel"] *= np.float32(out_scale)
        lp["encoder_attn"]["out_proj"]["kernel"] *= np.float32(out_scale)
        lp["fc2"]["kernel"] *= np.float32(out_scale)
    p["model"]["shared"]["embedding"] *= np.float32(emb_std / std)
    p["model"]["decoder"]["layernorm_embedding"]["scale"] *= np.float32(ln_emb_scale)
    p["model"]["decoder"]["layernorm_embedding"]["bias"] *= np.float32(ln_emb_scale)
    rng = np.random.default_rng(seed + 1000)
    flb = rng.standard_normal((1, t.vocab_size), dtype=np.float32) * np.float32(bias_std)
    flb[0, t.eos_token_id] = np.float32(eos_bias)
    p["final_logits_bias"] = flb
    return p
