"""Operator layer: torch tensors in, C-ABI calls out (include/mic_b200.h).

torch is used only for device memory and the current CUDA stream; every function enqueues hand-written
sm_100a kernels from libmic_b200.so.  Nothing here computes on the host and nothing falls back to
torch math.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import lib, check

ACT = {"none": 0, None: 0, "gelu": 1, "quick_gelu": 2}
BF16 = torch.bfloat16
F32 = torch.float32
I32 = torch.int32

LAUNCHES = [0]      # number of C-ABI kernel-launching calls (bench.py reports it)


def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


TIMED = {}          # C-ABI name -> list of (start, end) CUDA events recorded around each call (bench.py)
TIMED_FLOPS = {}    # C-ABI name -> list of 2*M*N*K of the same calls (GEMM family only)
TIMED_SHAPES = []   # (M, N, K, a_mn, b_mn, block_n, split_k, act, accumulate, d_f32) per timed mic_gemm_bf16 call (tools/gemm_table.py)


def _call(name, *args):
    LAUNCHES[0] += 1
    ev = TIMED.get(name)
    if ev is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        check(getattr(lib(), name)(_s(), *args), name)
        e.record()
        ev.append((s, e))
        return
    check(getattr(lib(), name)(_s(), *args), name)


def _drop(dropout):
    """dropout = (seed_tensor (device int32[1]), site:int, p:float) or None -> C-ABI triple"""
    if dropout is None or dropout[2] <= 0.0:
        return None, 0, 0.0
    return dropout[0].data_ptr(), int(dropout[1]) & 0xFFFFFFFF, float(dropout[2])


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, f"need a row-major 2-D view, got {tuple(t.shape)} {t.stride()}"
    return t.stride(0)


# ------------------------------------------------------------------------------------------------
# GEMM family
# ------------------------------------------------------------------------------------------------
def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=BF16, bias=None, act="none", pre_act_out=None,
         residual=None, accumulate=False, block_n=0, group_m=0, split_k=0, dropout=None, act_bwd=None):
    """D[M,N] = act(A[M,K] B[N,K]^T + bias) + residual.
    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True, Flax kernel)."""
    assert a.dtype == BF16 and b.dtype == BF16
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N)
    d_f32 = out.dtype == F32
    if pre_act_out is not None:
        assert pre_act_out.dtype == BF16 and _ld(pre_act_out) == _ld(out)
    if bias is not None:
        assert bias.dtype == F32 and bias.numel() == N
    act_code = ACT[act]
    if act_bwd is not None:          # (activation name, saved pre-activation U): D = (A B^T) o act'(U)
        assert residual is None and bias is None and act_code == 0
        act_code, residual, block_n = -ACT[act_bwd[0]], act_bwd[1], 256
    if "mic_gemm_bf16" in TIMED:
        TIMED_FLOPS.setdefault("mic_gemm_bf16", []).append(2.0 * M * N * K)
        TIMED_SHAPES.append((M, N, K, int(a_mn), int(b_mn), block_n, split_k, act_code, int(accumulate), int(d_f32)))
    _call("mic_gemm_bf16", int(a_mn), int(b_mn), _p(a), _ld(a), _p(b), _ld(b), M, N, K, _p(out), _ld(out),
          int(d_f32), int(accumulate), _p(bias), act_code, _p(pre_act_out), _p(residual),
          _ld(residual) if residual is not None else 0, block_n, group_m, split_k, *_drop(dropout))
    return out


def lm_head_num_partials(V):
    return lib().mic_lm_head_num_partials(V)


def lm_head_search_num_partials(M):
    return lib().mic_lm_head_search_num_partials(M)


def pack_kmajor_tiles_bytes(rows, K, tile_rows):
    return lib().mic_pack_kmajor_tiles_bytes(rows, K, tile_rows)


def pack_kmajor_tiles(src, tile_rows, out):
    rows, K = src.shape
    _call("mic_pack_kmajor_tiles", _p(src), _ld(src), rows, K, tile_rows, _p(out))


def _gumbel_key(key):
    """(k0, k1) uint32 pair -> host array for the C-ABI (None = no sampling noise)."""
    if key is None:
        return None, None
    arr = (ctypes.c_uint32 * 2)(int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF)
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def lm_head_search_packed(h_tiles, e_tiles, bias, mask_token, M, V, K, ws, second_pass=False, active=None,
                          gumbel_key=None):
    keep, gk = _gumbel_key(gumbel_key)
    _call("mic_lm_head_search_packed", _p(h_tiles), _p(e_tiles), _p(bias), mask_token, M, V, K, _p(ws["pmax"]),
          _p(ws["psum"]), _p(ws["cand_val"]), _p(ws["cand_idx"]), _p(ws["last_val"]) if second_pass else None,
          _p(ws["last_idx"]) if second_pass else None, _p(active), gk)


def lm_head_ce_stats(h, emb, bias, labels, ws, logits_out=None):
    M, K = h.shape
    V = emb.shape[0]
    if "mic_lm_head_ce_stats" in TIMED:
        TIMED_FLOPS.setdefault("mic_lm_head_ce_stats", []).append(2.0 * M * V * K)
    _call("mic_lm_head_ce_stats", _p(h), _ld(h), _p(emb), _ld(emb), _p(bias), _p(labels), M, V, K, _p(ws["pmax"]),
          _p(ws["psum"]), _p(ws["psumz"]), _p(ws["zlabel"]), _p(logits_out),
          _ld(logits_out) if logits_out is not None else 0)


def ce_softmax_bwd_workspace_floats(M, ld):
    return lib().mic_ce_softmax_bwd_workspace_floats(M, ld)


def ce_softmax_bwd(logits_inout, labels, ws, conf, low, V, dbias, workspace):
    M = logits_inout.shape[0]
    _call("mic_ce_softmax_bwd", _p(logits_inout), _ld(logits_inout), _p(labels), _p(ws["lse"]), _p(ws["row_w"]),
          float(conf), float(low), M, V, _p(dbias), _p(workspace), _p(counters(logits_inout.device)))


def ce_finalize(ws, mask, M, V, label_smoothing, with_loss=True):
    _call("mic_ce_finalize", _p(ws["pmax"]), _p(ws["psum"]), _p(ws.get("psumz")), _p(ws.get("zlabel")), _p(mask),
          ws["nparts"], M, V, float(label_smoothing), _p(ws["lse"]), _p(ws["row_loss"]) if with_loss else None,
          _p(ws["row_w"]) if with_loss else None, _p(ws["out"]) if with_loss else None)


def lm_head_ce_grad(h, emb, bias, labels, ws, conf, low, dlogits):
    M, K = h.shape
    V = emb.shape[0]
    _call("mic_lm_head_ce_grad", _p(h), _ld(h), _p(emb), _ld(emb), _p(bias), _p(labels), _p(ws["lse"]),
          _p(ws["row_w"]), float(conf), float(low), M, V, K, _p(dlogits), _ld(dlogits))


def lm_head_search(h, emb, bias, mask_token, ws, second_pass=False, active=None, gumbel_key=None):
    M, K = h.shape
    V = emb.shape[0]
    keep, gk = _gumbel_key(gumbel_key)
    _call("mic_lm_head_search", _p(h), _ld(h), _p(emb), _ld(emb), _p(bias), int(mask_token), M, V, K,
          _p(ws["pmax"]), _p(ws["psum"]), _p(ws["cand_val"]), _p(ws["cand_idx"]),
          _p(ws["last_val"]) if second_pass else None, _p(ws["last_idx"]) if second_pass else None, _p(active), gk)


def search_merge(ws, R, second_pass=False):
    """Top-8 per row of the slab partials -> ws["row_lp"/"row_tok"] [R, cpr] (cpr = 8, or 16 for 5..8 beams: the
    second pass fills columns 8..15 with the normaliser of the first)."""
    cpr = ws["row_lp"].shape[1]
    _call("mic_search_merge", _p(ws["pmax"]), _p(ws["psum"]), _p(ws["cand_val"]), _p(ws["cand_idx"]), ws["nparts"],
          R, _p(ws["row_lp"]), _p(ws["row_tok"]), _p(ws["row_ml"]), cpr, 8 if second_pass else 0, int(second_pass),
          _p(ws.get("last_val")), _p(ws.get("last_idx")))


# ------------------------------------------------------------------------------------------------
# normalisation / embedding / elementwise
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, out=None, mean=None, rstd=None):
    M, d = x.shape
    if out is None:
        out = torch.empty_like(x)
    _call("mic_layernorm_fwd", _p(x), _p(gamma), _p(beta), float(eps), _p(out), _p(mean), _p(rstd), M, d)
    return out


def residual_ln_fwd(acc, bias, x, gamma, beta, eps, out):
    M, d = x.shape
    _call("mic_residual_ln_fwd", _p(acc), _p(bias), _p(x), _p(gamma), _p(beta), float(eps), _p(out), M, d)
    return out


_LAYER_FIELDS = ("ln_sa_g", "ln_sa_b", "sa_qkv_w", "sa_qkv_b", "sa_o_w", "sa_o_b", "ln_ca_g", "ln_ca_b", "ca_q_w", "ca_q_b",
                 "ca_o_w", "ca_o_b", "ln_f_g", "ln_f_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "self_kv", "enc_k", "enc_v")
_BUFFER_FIELDS = ("x", "q", "a_tiles", "o_tiles", "g_tiles", "acc", "q_acc", "ancestors", "h_out", "ln_out_g", "ln_out_b",
                  "h_out_tiles", "cross_kv_tiles")


class DecoderLayerT(ctypes.Structure):        # mic_decoder_layer_t
    _fields_ = [(n, ctypes.c_void_p) for n in _LAYER_FIELDS]


class DecoderBuffersT(ctypes.Structure):      # mic_decoder_buffers_t
    _fields_ = [(n, ctypes.c_void_p) for n in _BUFFER_FIELDS]


def decoder_plan_bytes(num_layers):
    return lib().mic_decoder_plan_bytes(num_layers)


def decoder_packed_bytes(num_layers, d_model, ffn_dim):
    return lib().mic_decoder_packed_bytes(num_layers, d_model, ffn_dim)


def decoder_layers_struct(layers):
    """layers: list of dicts field -> tensor -> ctypes array of mic_decoder_layer_t"""
    arr = (DecoderLayerT * len(layers))()
    for i, l in enumerate(layers):
        for n in _LAYER_FIELDS:
            setattr(arr[i], n, l[n].data_ptr())
    return arr


def decoder_pack_weights(layers_struct, d_model, ffn_dim, packed):
    _call("mic_decoder_pack_weights", ctypes.cast(layers_struct, ctypes.c_void_p), len(layers_struct), d_model, ffn_dim,
          _p(packed))


def decoder_plan_init(plan, layers_struct, buffers, packed, R, d_model, heads, ffn_dim, cache_len, enc_tokens,
                      rows_per_image, ld_enc, act, eps):
    bt = DecoderBuffersT()
    for n in _BUFFER_FIELDS:
        setattr(bt, n, _p(buffers[n]))
    _call("mic_decoder_plan_init", _p(plan), ctypes.cast(layers_struct, ctypes.c_void_p), len(layers_struct),
          ctypes.cast(ctypes.pointer(bt), ctypes.c_void_p), _p(packed), R, d_model, heads, ffn_dim, cache_len,
          enc_tokens, rows_per_image, ld_enc, ACT[act], eps)


def decoder_cross_kv_tiles_bytes(B, num_layers, heads):
    return lib().mic_decoder_cross_kv_tiles_bytes(B, num_layers, heads)


def decoder_pack_cross_kv(enc_kv, B, S, num_layers, heads, d_model, out):
    _call("mic_decoder_pack_cross_kv", _p(enc_kv), _ld(enc_kv), B, S, num_layers, heads, d_model, _p(out))


DECODER_STEP_OPTS = [int(__import__("os").environ.get("MIC_DECODER_OPTS", "0"))]     # tuning switches (include/mic_b200.h)


def decoder_step(plan, num_layers, R, pos, sync, phase_times=None, active=None, opts=None):
    _call("mic_decoder_step", _p(plan), num_layers, R, pos, _p(sync), _p(phase_times), _p(active),
          DECODER_STEP_OPTS[0] if opts is None else int(opts))


def launch_options(pdl=-1, gemm_b_static=-1, gemm_sm_margin=-1):
    """mic_launch_options: programmatic dependent launch on/off, weights-are-static hint for the decode loop, SMs the
    persistent GEMMs leave free for a concurrent NCCL collective."""
    lib().mic_launch_options(int(pdl), int(gemm_b_static), int(gemm_sm_margin))


_COUNTERS = {}


def counters(device):
    """Ticket counters of the 'last CTA reduces' kernels: zero-initialised once per device, kernels reset them."""
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    c = _COUNTERS.get(key)
    if c is None:
        c = torch.zeros(1024, dtype=torch.int32, device=dev)
        _COUNTERS[key] = c
    return c


def ln_bwd_workspace_floats(M, d):
    return lib().mic_layernorm_bwd_workspace_floats(M, d)


def layernorm_bwd(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, workspace):
    M, d = x.shape
    _call("mic_layernorm_bwd", _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx), _p(dgamma),
          _p(dbeta), _p(workspace), _p(counters(x.device)), M, d)
    return dx


def colsum_workspace_floats(M, N):
    return lib().mic_colsum_workspace_floats(M, N)


def act_bwd_colsum(dy, u, act, du, dbias, workspace, accumulate=False, dropout=None, side=False):
    """side=True: second half of the ticket counters (a launch on another stream may run concurrently with a
    main-stream launch of the same kernel family; N <= 131072 either way)."""
    M, N = dy.shape
    cnt = counters(dy.device)
    if side:
        cnt = cnt[512:]
    _call("mic_act_bwd_colsum", _p(dy), _ld(dy), _p(u), _ld(u) if u is not None else 0, ACT[act], _p(du),
          _ld(du) if du is not None else 0, _p(dbias), int(accumulate), _p(workspace), _p(cnt), M, N,
          *_drop(dropout))


def embed_ln_fwd(ids, pos_ids, pos_mod, pos_offset, table, pos_table, scale, gamma, beta, eps, emb, out,
                 mean=None, rstd=None, dropout=None):
    M = ids.numel()
    d = table.shape[1]
    _call("mic_embed_ln_fwd", _p(ids), _p(pos_ids), int(pos_mod), int(pos_offset), _p(table), _p(pos_table),
          float(scale), _p(gamma), _p(beta), float(eps), _p(emb), _p(out), _p(mean), _p(rstd), M, d, *_drop(dropout))
    return out


def embed_bwd(ids, d_emb, scale, d_table, d_pos_rows, B, T, hot_id=-1):
    d = d_emb.shape[-1]
    _call("mic_embed_bwd", _p(ids), _p(d_emb), float(scale), _p(d_table), _p(d_pos_rows), B, T, d, int(hot_id))


def batch_sum(x, B, T, d, out, out_ld):
    _call("mic_batch_sum", _p(x), B, T, d, _p(out), out_ld)


def patchify(pixels, out, B, image_size, patch, channel_first=False, trunc_int=False, mean=None, std=None):
    """fp32 pixels (already normalised, the reference's input contract) or uint8 pixels (input hand-off: x/255 and
    Normalize(mean, std) of main.py:173-174 are applied in the kernel)."""
    assert pixels.is_contiguous()
    if pixels.dtype == torch.uint8:
        m3 = (ctypes.c_float * 3)(*[float(x) for x in mean])
        s3 = (ctypes.c_float * 3)(*[float(x) for x in std])
        _call("mic_patchify_u8", _p(pixels), _p(out), B, image_size, patch, int(channel_first), int(trunc_int),
              ctypes.cast(m3, ctypes.c_void_p), ctypes.cast(s3, ctypes.c_void_p))
        return out
    assert pixels.dtype == F32
    _call("mic_patchify", _p(pixels), _p(out), B, image_size, patch, int(channel_first), int(trunc_int))
    return out


def resize_crop_u8(blob, desc, n, size, out, channel_first=False):
    """Resize([size], BICUBIC) + CenterCrop(size) of n packed uint8 CHW images (desc: int64 [n,8], see mic_b200.h)."""
    assert blob.dtype == torch.uint8 and out.dtype == torch.uint8 and desc.dtype == torch.int64 and desc.shape == (n, 8)
    assert blob.is_contiguous() and out.is_contiguous() and desc.is_contiguous()
    _call("mic_resize_crop_u8", _p(blob), _p(desc), n, size, int(channel_first), _p(out))
    return out


def vit_embed_ln_fwd(patch_out, patch_bias, cls, pos, gamma, beta, eps, use_ln, emb, out, mean, rstd, B, S):
    d = patch_out.shape[1]
    _call("mic_vit_embed_ln_fwd", _p(patch_out), _p(patch_bias), _p(cls), _p(pos), _p(gamma), _p(beta), float(eps),
          int(use_ln), _p(emb), _p(out), _p(mean), _p(rstd), B, S, d)
    return out


def drop_cls_rows(d_emb, out, B, S):
    d = d_emb.shape[-1]
    _call("mic_drop_cls_rows", _p(d_emb), _p(out), B, S, d)
    return out


def adamw(p, m, v, g, shadow, lr, b1, b2, eps, weight_decay, bias_corr1, bias_corr2, grad_scale):
    """The eight scalars are passed by value (kernel arguments fixed at enqueue time)."""
    _call("mic_adamw", _p(p), _p(m), _p(v), _p(g), _p(shadow), p.numel(), float(lr), float(b1), float(b2), float(eps),
          float(weight_decay), float(bias_corr1), float(bias_corr2), float(grad_scale))


def cast_f32_to_bf16(src, dst):
    _call("mic_cast_f32_to_bf16", _p(src), _p(dst), src.numel())
    return dst


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def attention_fwd(q, k, v, out, lse, key_mask, causal, B, H, Tq, Tk, scale):
    _call("mic_attention_fwd", _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(out), _ld(out), _p(lse),
          _p(key_mask), int(causal), B, H, Tq, Tk, 64, float(scale))
    return out


def attention_impl(impl):
    """A/B switch: 0 = row-tiled kernels (default; single-kernel backward), 1 = one-CTA-per-head kernels where they
    apply (<= 64 tokens), 2 = row-tiled with the backward as a dQ and a dK/dV kernel."""
    check(lib().mic_attention_impl(int(impl)), "mic_attention_impl")       # no stream argument: a host-side switch


def attention_bwd(q, k, v, o, do, lse, key_mask, causal, dq, dk, dv, B, H, Tq, Tk, scale):
    _call("mic_attention_bwd", _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(o), _ld(o), _p(do), _ld(do), _p(lse),
          _p(key_mask), int(causal), _p(dq), _ld(dq), _p(dk), _ld(dk), _p(dv), _ld(dv), B, H, Tq, Tk, 64,
          float(scale))


def decode_attention(q, k_cache, v_cache, ldkv, ancestors, cache_len, n_keys, rows_per_kv, out, R, H, scale):
    _call("mic_decode_attention", _p(q), _ld(q), _p(k_cache), _p(v_cache), int(ldkv), _p(ancestors), cache_len,
          n_keys, rows_per_kv, _p(out), _ld(out), R, H, 64, float(scale))
    return out


# ------------------------------------------------------------------------------------------------
# search steps
# ------------------------------------------------------------------------------------------------
def beam_step(ws, st, forced_token, B, K, L, V, cur_len, eos, early_stopping, length_penalty):
    _call("mic_beam_step", _p(ws["row_lp"]), _p(ws["row_tok"]), int(forced_token), B, K, L, V, cur_len, eos,
          int(early_stopping), float(length_penalty), _p(st["running_seq"]), _p(st["running_scores"]),
          _p(st["sequences"]), _p(st["scores"]), _p(st["finished"]), _p(st["ancestors"]), _p(st["next_token"]),
          _p(st["active"]), ws["row_lp"].shape[1])


def beam_cond(st, B, K, cur_len, max_length, length_penalty, early_stopping):
    _call("mic_beam_cond", _p(st["running_scores"]), _p(st["scores"]), _p(st["finished"]), B, K, cur_len, max_length,
          float(length_penalty), int(early_stopping), _p(st["active"]))


def beam_finalize(st, B, K, L, out_seq, out_scores):
    _call("mic_beam_finalize", _p(st["sequences"]), _p(st["scores"]), _p(st["finished"]), _p(st["running_seq"]),
          _p(st["running_scores"]), B, K, L, _p(out_seq), _p(out_scores))


def greedy_step(ws, st, forced_token, R, L, cur_len, eos, pad):
    _call("mic_greedy_step", _p(ws["row_tok"]), int(forced_token), R, L, cur_len, eos, pad, _p(st["sequences"]),
          _p(st["finished"]), _p(st["next_token"]), _p(st["active"]))


def greedy_cond(st, R, cur_len, max_length):
    _call("mic_greedy_cond", _p(st["finished"]), R, cur_len, max_length, _p(st["active"]))
