"""ORACLE — test infrastructure only.  PARITY UNPINNED BY THE REFERENCE (it ships no tests/goldens).

CPU fp32 restatement (torch-CPU, so autograd supplies the gradient oracle) of the arithmetic behind
`FlaxCLIPVisionMBartForConditionalGeneration` and the `flax_vit_bart` variant.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it;
the product package never does.

The reference wires upstream modules and holds little arithmetic of its own, so each
function cites (a) the reference site that instantiates / calls it and (b) the upstream module whose
published algorithm is restated.  Upstream pins (requirements.txt:10,13,14,21,39):
transformers@0085e712ddf80fa5cd5f355498fe7f13b839eafa, flax==0.3.4, jax==0.2.16, optax==0.0.9 — none
installable here, so the restatement is cross-checked block by block against the HF *PyTorch* twins
(`tests/test_oracle_vs_hf.py`) and against analytic known-answer tests (`tests/test_oracle_kat.py`).
Switches for the unverifiable-offline items (SURVEY.md §8c U1,U2,U7) are config fields / kwargs.

Parameter trees are nested dicts with Flax names; leaves may be numpy arrays or
torch tensors.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(np.ascontiguousarray(x))


def to_torch_tree(tree, requires_grad=False):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out[k] = to_torch_tree(v, requires_grad)
        else:
            t = _t(v).detach().clone().float()
            t.requires_grad_(requires_grad)
            out[k] = t
    return out


# ----------------------------------------------------------------------------------------------
# primitives (flax.linen 0.3.4 semantics)
# ----------------------------------------------------------------------------------------------
def layer_norm(x, p, eps, flax_variance=True):
    """flax.linen.LayerNorm (0.3.4): mean, mean-of-squares, var = E[x^2]-E[x]^2 (risk U7), biased."""
    mean = x.mean(-1, keepdim=True)
    if flax_variance:
        var = (x * x).mean(-1, keepdim=True) - mean * mean
        var = torch.clamp(var, min=0.0)
    else:
        var = ((x - mean) ** 2).mean(-1, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    return y * _t(p["scale"]) + _t(p["bias"])


def dense(x, p):
    """flax.linen.Dense: x @ kernel(in,out) + bias."""
    y = x @ _t(p["kernel"])
    if "bias" in p:
        y = y + _t(p["bias"])
    return y


def activation(x, name, gelu_approximate=False):
    """transformers ACT2FN (Flax): quick_gelu = x*sigmoid(1.702x); gelu = exact erf (risk U2)."""
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu(x, approximate="tanh" if gelu_approximate else "none")
    if name == "relu":
        return F.relu(x)
    raise ValueError(name)


def attention(q_in, kv_in, p, num_heads, bias=None):
    """FlaxCLIPAttention / FlaxMBartAttention: q,k,v,out Dense with bias; query scaled by
    1/sqrt(head_dim) BEFORE q.k^T (flax dot_product_attention_weights); softmax over keys; additive
    bias (0 / -inf).  HF-PT twins: modeling_clip.py:282-336, modeling_mbart.py:156-271."""
    B, Tq, D = q_in.shape
    Tk = kv_in.shape[1]
    hd = D // num_heads
    q = dense(q_in, p["q_proj"]).view(B, Tq, num_heads, hd)
    k = dense(kv_in, p["k_proj"]).view(B, Tk, num_heads, hd)
    v = dense(kv_in, p["v_proj"]).view(B, Tk, num_heads, hd)
    q = q / math.sqrt(hd)
    w = torch.einsum("bqhd,bkhd->bhqk", q, k)
    if bias is not None:
        w = w + bias
    w = torch.softmax(w, dim=-1)
    o = torch.einsum("bhqk,bkhd->bqhd", w, v).reshape(B, Tq, D)
    return dense(o, p["out_proj"])


# ----------------------------------------------------------------------------------------------
# vision encoder  (reference: modeling_clip_vision_mbart.py:46-48,79-90; modeling_vit_bart.py:58-81)
# ----------------------------------------------------------------------------------------------
def vision_encoder(params, pixel_values, config):
    """FlaxCLIPVisionModule (HF-PT twin modeling_clip.py:138-218,354-386,647-697) or, with the
    ViT switches of the config, FlaxViTModule (modeling_vit.py:43-128,199-347,402-458).
    Returns last_hidden_state WITHOUT post_layernorm for CLIP; WITH `layernorm` for ViT."""
    c = config.clip_vision_config
    vp = params["model"]["encoder"]["vision_model"]
    x = _t(pixel_values).float()
    if c.channel_first_input:            # modeling_vit_bart.py:445
        x = x.permute(0, 2, 3, 1)
    B = x.shape[0]
    g, ps = c.image_size // c.patch_size, c.patch_size
    # conv stride=kernel=patch, VALID, HWIO kernel == GEMM over (kh,kw,c)-flattened patches
    patches = x.reshape(B, g, ps, g, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, g * g, ps * ps * 3)
    w = _t(vp["embeddings"]["patch_embedding"]["kernel"]).reshape(ps * ps * 3, c.hidden_size)
    pe = patches @ w
    if "bias" in vp["embeddings"]["patch_embedding"]:
        pe = pe + _t(vp["embeddings"]["patch_embedding"]["bias"])
    cls = _t(vp["embeddings"]["class_embedding"]).reshape(1, 1, -1).expand(B, 1, -1)
    h = torch.cat([cls, pe], dim=1) + _t(vp["embeddings"]["position_embedding"]["embedding"])[None]
    if c.pre_layernorm:
        h = layer_norm(h, vp["pre_layrnorm"], c.layer_norm_eps)
    for i in range(c.num_hidden_layers):
        lp = vp["encoder"]["layers"][str(i)]
        a = layer_norm(h, lp["layer_norm1"], c.layer_norm_eps)
        h = h + attention(a, a, lp["self_attn"], c.num_attention_heads)
        m = layer_norm(h, lp["layer_norm2"], c.layer_norm_eps)
        h = h + dense(activation(dense(m, lp["mlp"]["fc1"]), c.hidden_act), lp["mlp"]["fc2"])
    if c.final_layernorm:
        h = layer_norm(h, vp["post_layernorm"], c.layer_norm_eps)
    return h


def encode(params, pixel_values, config, int32_cast=False):
    """`encode()` modeling_clip_vision_mbart.py:284-337: ViT then visual_projection (:322).
    int32_cast=True reproduces the `jnp.array(pixel_values, dtype="i4")` truncation at :330 that
    generate() sees; `__call__` (:501) keeps float32."""
    x = _t(pixel_values).float()
    if int32_cast:
        x = torch.trunc(x)          # float -> int32 conversion truncates toward zero
    h = vision_encoder(params, x, config)
    return dense(h, params["model"]["visual_projection"])      # :53-59,90


# ----------------------------------------------------------------------------------------------
# text decoder (reference: modeling_clip_vision_mbart.py:49-51,92-102)
# ----------------------------------------------------------------------------------------------
def _embed(params, ids, pos, config):
    t = config.mbart_config
    dp = params["model"]["decoder"]
    scale = math.sqrt(t.d_model) if t.scale_embedding else 1.0
    h = _t(params["model"]["shared"]["embedding"])[ids] * scale
    h = h + _t(dp["embed_positions"]["embedding"])[pos + t.position_offset]
    return layer_norm(h, dp["layernorm_embedding"], t.layer_norm_eps)


def _decoder_layer(h, lp, enc, self_bias, t, self_kv=None):
    """FlaxMBartDecoderLayer (pre-LN) / FlaxBartDecoderLayer (post-LN).
    HF-PT twins modeling_mbart.py:386-422, modeling_bart.py:354-389."""
    eps, H, act = t.layer_norm_eps, t.decoder_attention_heads, t.activation_function
    kv = (lambda a: a) if self_kv is None else self_kv
    if t.pre_layernorm:
        a = layer_norm(h, lp["self_attn_layer_norm"], eps)
        h = h + attention(a, kv(a), lp["self_attn"], H, self_bias)
        c = layer_norm(h, lp["encoder_attn_layer_norm"], eps)
        h = h + attention(c, enc, lp["encoder_attn"], H)
        f = layer_norm(h, lp["final_layer_norm"], eps)
        h = h + dense(activation(dense(f, lp["fc1"]), act), lp["fc2"])
    else:
        h = layer_norm(h + attention(h, kv(h), lp["self_attn"], H, self_bias), lp["self_attn_layer_norm"], eps)
        h = layer_norm(h + attention(h, enc, lp["encoder_attn"], H), lp["encoder_attn_layer_norm"], eps)
        h = layer_norm(h + dense(activation(dense(h, lp["fc1"]), act), lp["fc2"]), lp["final_layer_norm"], eps)
    return h


def decoder(params, decoder_input_ids, decoder_attention_mask, position_ids, enc, config):
    """FlaxMBartDecoder full-sequence pass.  Mask = causal AND key padding
    (`decoder_attention_mask`), additive 0 / -inf.  Dropout is identity (deterministic)."""
    t = config.mbart_config
    ids = _t(decoder_input_ids).long()
    B, T = ids.shape
    mask = torch.ones(B, T, dtype=torch.long) if decoder_attention_mask is None else _t(decoder_attention_mask).long()
    pos = torch.arange(T)[None].expand(B, T) if position_ids is None else _t(position_ids).long()
    h = _embed(params, ids, pos, config)
    causal = torch.tril(torch.ones(T, T, dtype=torch.bool))
    allow = causal[None, None] & (mask[:, None, None, :] > 0)
    bias = torch.zeros(B, 1, T, T).masked_fill(~allow, float("-inf"))
    dp = params["model"]["decoder"]
    for i in range(t.decoder_layers):
        h = _decoder_layer(h, dp["layers"][str(i)], enc, bias, t)
    if t.final_layer_norm:
        h = layer_norm(h, dp["layer_norm"], t.layer_norm_eps)
    return h


def lm_head(params, h):
    """modeling_clip_vision_mbart.py:170-178: tied kernel = shared.embedding.T, then += final_logits_bias."""
    return h @ _t(params["model"]["shared"]["embedding"]).t() + _t(params["final_logits_bias"])


def forward_logits(params, pixel_values, decoder_input_ids, decoder_attention_mask=None,
                   decoder_position_ids=None, config=None):
    """`__call__` modeling_clip_vision_mbart.py:447-510 -> logits (B,T,V)."""
    enc = encode(params, pixel_values, config, int32_cast=False)
    h = decoder(params, decoder_input_ids, decoder_attention_mask, decoder_position_ids, enc, config)
    return lm_head(params, h)


# ----------------------------------------------------------------------------------------------
# loss / grads / optimiser  (main.py:658-680, 684-707, 281-292, 629-635)
# ----------------------------------------------------------------------------------------------
def loss_fn(logits, labels, padding_mask, label_smoothing_factor=0.0):
    """main.py:658-680, line by line."""
    V = logits.shape[-1]
    confidence = 1.0 - label_smoothing_factor
    low = (1.0 - confidence) / (V - 1)
    const = -(confidence * math.log(confidence) + (V - 1) * low * math.log(low + 1e-20)) \
        if confidence > 0 else -((V - 1) * low * math.log(low + 1e-20))
    labels = _t(labels).long()
    soft = torch.full(logits.shape, low, dtype=logits.dtype)
    soft.scatter_(-1, labels[..., None], confidence)
    loss = -(soft * torch.log_softmax(logits, dim=-1)).sum(-1)      # optax.softmax_cross_entropy
    loss = loss - const
    m = _t(padding_mask).to(logits.dtype)
    return (loss * m).sum() / m.sum()


def loss_and_grads(params, batch, config, label_smoothing_factor=0.0):
    """compute_loss + jax.value_and_grad (main.py:688-697) on one device's shard.
    Note main.py:692 passes batch["attention_mask"] (the LABEL mask) as decoder_attention_mask."""
    p = to_torch_tree(params, requires_grad=True)
    logits = forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"],
                            None, config)
    loss = loss_fn(logits, batch["input_ids"], batch["attention_mask"], label_smoothing_factor)
    loss.backward()

    def grads(tree):
        return {k: (grads(v) if isinstance(v, dict) else
                    (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()) for k, v in tree.items()}
    return float(loss.detach()), grads(p), logits.detach()


def linear_warmup_decay_lr(step, lr, warmup_steps, total_steps):
    """create_learning_rate_fn main.py:281-292 (optax.linear_schedule joined at warmup_steps)."""
    if step < warmup_steps:
        return lr * step / max(warmup_steps, 1)
    frac = min(max((step - warmup_steps) / max(total_steps - warmup_steps, 1), 0.0), 1.0)
    return lr * (1.0 - frac)


def adamw_update(p, g, m, v, count, lr_at_count, b1=0.9, b2=0.999, eps=1e-8, weight_decay=0.0):
    """optax 0.0.9 `adamw` = scale_by_adam -> add_decayed_weights -> scale_by_schedule(-lr) (risk U8).
    `count` is the pre-increment step (0 for the first update); bias correction uses count+1; the
    schedule is evaluated at `count` (main.py:629-635, TrainState.apply_gradients main.py:701)."""
    t = count + 1
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** t)
    vhat = v / (1 - b2 ** t)
    upd = mhat / (np.sqrt(vhat) + eps) + weight_decay * p
    return (p - lr_at_count * upd).astype(np.float32), m.astype(np.float32), v.astype(np.float32)
