"""fp32 verification path: forward + loss with fp32 storage and fp32 SIMT kernels (csrc/fp32_path.cu).

Mirrors `engine.CaptionEngine.encode / decoder_forward / logits / loss` op for op (same parameter store: the fp32
MASTER copy of the weights instead of the bf16 shadow), for the parity bar of BASELINE configs[0]: batch 8, fp32,
logits within 1e-3 relative and loss within 1e-4 of the reference restatement (`oracle/`).  Not a training path and
not a performance path - the product path is the bf16 tcgen05 one in engine.py.
Reference: modeling_clip_vision_mbart.py:447-510 (__call__), main.py:658-680 (loss)."""
from __future__ import annotations

import math

import torch

from ._lib import lib, check
from .ops import ACT

F32, I32 = torch.float32, torch.int32


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, (tuple(t.shape), t.stride())
    return t.stride(0)


def gemm(a, b, *, b_nk=False, bias=None, act="none", residual=None, out=None):
    """out[M,N] = act(a[M,K] @ B + bias) + residual; B = b[K,N] (Flax kernel) or b[N,K]^T (b_nk, tied lm_head)."""
    M, K = a.shape
    N = b.shape[0] if b_nk else b.shape[1]
    assert (b.shape[1] if b_nk else b.shape[0]) == K
    if out is None:
        out = torch.empty((M, N), dtype=F32, device=a.device)
    check(lib().mic_f32_gemm(_s(), _p(a), _ld(a), _p(b), _ld(b), int(b_nk), M, N, K, _p(bias), ACT[act], _p(residual),
                             _ld(residual) if residual is not None else 0, _p(out), _ld(out)), "mic_f32_gemm")
    return out


def layernorm(x, gamma, beta, eps):
    y = torch.empty_like(x)
    check(lib().mic_f32_layernorm(_s(), _p(x), _p(gamma), _p(beta), float(eps), _p(y), x.shape[0], x.shape[1]),
          "mic_f32_layernorm")
    return y


def attention(q, k, v, key_mask, causal, B, H, Tq, Tk, scale):
    o = torch.empty((B * Tq, H * 64), dtype=F32, device=q.device)
    check(lib().mic_f32_attention(_s(), _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(o), _ld(o), _p(key_mask),
                                  int(causal), B, H, Tq, Tk, 64, float(scale)), "mic_f32_attention")
    return o


class Fp32Forward:
    def __init__(self, engine):
        self.e = engine

    # ---- vision tower + visual projection (FlaxCLIPVisionModel / ViT variant; engine.encode) ----
    def encode(self, pixel_values, trunc_int=False):
        e = self.e
        c, ps = e.c, e.ps
        B = pixel_values.shape[0]
        S, npatch, dv = c.num_tokens, c.num_patches, c.hidden_size
        px = pixel_values.to(e.dev, F32).contiguous()
        patches = torch.empty((B * npatch, c.patch_size ** 2 * 3), dtype=F32, device=e.dev)
        check(lib().mic_f32_patchify(_s(), _p(px), _p(patches), B, c.image_size, c.patch_size,
                                     int(c.channel_first_input), int(trunc_int)), "mic_f32_patchify")
        patch_out = gemm(patches, ps.f("v.patch.w"))
        x = torch.empty((B * S, dv), dtype=F32, device=e.dev)
        check(lib().mic_f32_vit_embed(_s(), _p(patch_out), _p(ps.f("v.patch.b")) if c.patch_bias else None,
                                      _p(ps.f("v.cls")), _p(ps.f("v.pos")), _p(x), B, S, dv), "mic_f32_vit_embed")
        if c.pre_layernorm:
            x = layernorm(x, ps.f("v.pre_ln.scale"), ps.f("v.pre_ln.bias"), c.layer_norm_eps)
        H = c.num_attention_heads
        scale = 1.0 / math.sqrt(c.head_dim)
        for l in range(c.num_hidden_layers):
            n = f"v.{l}"
            a = layernorm(x, ps.f(n + ".ln1.scale"), ps.f(n + ".ln1.bias"), c.layer_norm_eps)
            qkv = gemm(a, ps.f(n + ".qkv.w"), bias=ps.f(n + ".qkv.b"))
            att = attention(qkv[:, :dv], qkv[:, dv:2 * dv], qkv[:, 2 * dv:], None, False, B, H, S, S, scale)
            xm = gemm(att, ps.f(n + ".o.w"), bias=ps.f(n + ".o.b"), residual=x)
            m = layernorm(xm, ps.f(n + ".ln2.scale"), ps.f(n + ".ln2.bias"), c.layer_norm_eps)
            g = gemm(m, ps.f(n + ".fc1.w"), bias=ps.f(n + ".fc1.b"), act=c.hidden_act)
            x = gemm(g, ps.f(n + ".fc2.w"), bias=ps.f(n + ".fc2.b"), residual=xm)
        if c.final_layernorm:
            x = layernorm(x, ps.f("v.post_ln.scale"), ps.f("v.post_ln.bias"), c.layer_norm_eps)
        return gemm(x, ps.f("proj.w"), bias=ps.f("proj.b"))

    # ---- decoder (FlaxMBartDecoder pre-LN / FlaxBartDecoder post-LN; engine.decoder_forward) ----
    def decoder(self, ids, key_mask, enc, B, T):
        e = self.e
        t, ps, c = e.t, e.ps, e.c
        d, M, H, S = t.d_model, B * T, t.decoder_attention_heads, c.num_tokens
        eps = t.layer_norm_eps
        scale = 1.0 / math.sqrt(t.head_dim)
        emb = torch.empty((M, d), dtype=F32, device=e.dev)
        check(lib().mic_f32_embed(_s(), _p(ids), _p(ps.f("shared")), float(e.emb_scale), _p(ps.f("d.pos")),
                                  t.position_offset, T, _p(emb), M, d), "mic_f32_embed")
        x = layernorm(emb, ps.f("d.ln_emb.scale"), ps.f("d.ln_emb.bias"), eps)
        enc_kv = gemm(enc, ps.f("d.ca_kv.w"), bias=ps.f("d.ca_kv.b"))
        ln = lambda v, name: layernorm(v, ps.f(name + ".scale"), ps.f(name + ".bias"), eps)
        for l in range(t.decoder_layers):
            n = f"d.{l}"
            kl = enc_kv[:, l * 2 * d: l * 2 * d + d]
            vl = enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]
            if t.pre_layernorm:
                a = ln(x, n + ".ln_sa")
                qkv = gemm(a, ps.f(n + ".sa_qkv.w"), bias=ps.f(n + ".sa_qkv.b"))
                sa = attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], key_mask, True, B, H, T, T, scale)
                x1 = gemm(sa, ps.f(n + ".sa_o.w"), bias=ps.f(n + ".sa_o.b"), residual=x)
                qc = gemm(ln(x1, n + ".ln_ca"), ps.f(n + ".ca_q.w"), bias=ps.f(n + ".ca_q.b"))
                ca = attention(qc, kl, vl, None, False, B, H, T, S, scale)
                x2 = gemm(ca, ps.f(n + ".ca_o.w"), bias=ps.f(n + ".ca_o.b"), residual=x1)
                g = gemm(ln(x2, n + ".ln_f"), ps.f(n + ".fc1.w"), bias=ps.f(n + ".fc1.b"), act=t.activation_function)
                x = gemm(g, ps.f(n + ".fc2.w"), bias=ps.f(n + ".fc2.b"), residual=x2)
            else:       # BART: h = LN(h + sublayer(h))
                qkv = gemm(x, ps.f(n + ".sa_qkv.w"), bias=ps.f(n + ".sa_qkv.b"))
                sa = attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], key_mask, True, B, H, T, T, scale)
                x1 = ln(gemm(sa, ps.f(n + ".sa_o.w"), bias=ps.f(n + ".sa_o.b"), residual=x), n + ".ln_sa")
                qc = gemm(x1, ps.f(n + ".ca_q.w"), bias=ps.f(n + ".ca_q.b"))
                ca = attention(qc, kl, vl, None, False, B, H, T, S, scale)
                x2 = ln(gemm(ca, ps.f(n + ".ca_o.w"), bias=ps.f(n + ".ca_o.b"), residual=x1), n + ".ln_ca")
                g = gemm(x2, ps.f(n + ".fc1.w"), bias=ps.f(n + ".fc1.b"), act=t.activation_function)
                x = ln(gemm(g, ps.f(n + ".fc2.w"), bias=ps.f(n + ".fc2.b"), residual=x2), n + ".ln_f")
        if t.final_layer_norm:
            x = ln(x, "d.ln_final")
        return x

    def logits(self, pixel_values, decoder_input_ids, attention_mask=None):
        """[B, T, V] fp32 logits (lm_head tied to the embedding + final_logits_bias)."""
        e = self.e
        B, T = decoder_input_ids.shape
        assert B * T * e.t.vocab_size * 4 <= 8 << 30, "fp32 verification path: logits would exceed 8 GB"
        ids = decoder_input_ids.to(e.dev, I32).contiguous().view(-1)
        km = None if attention_mask is None else attention_mask.to(e.dev, I32).contiguous()
        enc = self.encode(pixel_values)
        hf = self.decoder(ids, km, enc, B, T)
        z = gemm(hf, e.ps.f("shared"), b_nk=True, bias=e.ps.f("flb"))
        return z.view(B, T, -1)

    def loss(self, pixel_values, decoder_input_ids, attention_mask, labels, label_smoothing=0.0):
        """main.py:658-680: label-smoothed CE averaged over the unmasked label positions."""
        e = self.e
        z = self.logits(pixel_values, decoder_input_ids, attention_mask)
        B, T, V = z.shape
        lab = labels.to(e.dev, I32).contiguous().view(-1)
        row_loss = torch.empty((B * T,), dtype=F32, device=e.dev)
        lse = torch.empty((B * T,), dtype=F32, device=e.dev)
        z2 = z.view(B * T, V)
        check(lib().mic_f32_ce_rows(_s(), _p(z2), _ld(z2), _p(lab), B * T, V, float(label_smoothing), _p(row_loss),
                                    _p(lse)), "mic_f32_ce_rows")
        m = attention_mask.to(e.dev, F32).reshape(-1)
        return (row_loss.double() * m.double()).sum() / m.double().sum(), lse.view(B, T), z
