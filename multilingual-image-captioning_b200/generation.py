"""KV-cached decode step and the greedy / beam-search drivers.

Replaces `generation_clip_vision_utils.py` (`generate` :128-336, `_greedy_search` :422-535,
`_beam_search` :665-990) and the cached `decode` path (`modeling_clip_vision_mbart.py:249-282,519-693`).
Differences in mechanism, not in results:
  * cross-attention K/V of the visual tokens are projected once per image (the reference re-projects
    them every step) and shared by the beams of an image;
  * the self-attention cache is never gathered on beam reorder (:945-953): beams carry an ancestor table;
  * the loop runs a fixed trip count with a device-side `active` flag mirroring the while_loop
    condition, so the host never synchronises inside the loop (CUDA-graph capturable).
"""
from __future__ import annotations

import math

import torch

from . import ops

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


class DecodeCache:
    """Device state of the cached decoder: self-attention K|V per layer, cross K|V, ancestor table."""

    def __init__(self, engine, rows, max_length, enc_kv, rows_per_image, use_ancestors):
        t = engine.t
        dev = engine.dev
        self.rows, self.T, self.rows_per_image = rows, max_length, rows_per_image
        self.self_kv = engine.bufs.get("gen.self_kv", (t.decoder_layers, rows, max_length, 2 * t.d_model))
        self.enc_kv = enc_kv
        self.ancestors = None
        if use_ancestors:       # stable address (the fused decoder plan holds it): refilled in place
            self.ancestors = engine.bufs.get("gen.ancestors", (rows, max_length), I32)
            self.ancestors.copy_(torch.arange(rows, dtype=I32, device=dev)[:, None].expand(rows, max_length))
        self.index = 0
        self.fused = None


def decode_step(engine, cache: DecodeCache, tokens, pos: int):
    """One token per row through the decoder; returns final hidden states [R, d].

    Weight-streaming regime (R = batch*beams rows): every GEMM uses 64-wide tiles so a weight matrix is
    pulled by many SMs at once; the three residual GEMMs of a layer (self out_proj, cross out_proj, fc2)
    additionally split K and reduce-add into one fp32 accumulator, which the next LayerNorm kernel folds
    into the residual stream (x += acc + bias; y = LN(x); acc = 0) - no bias/residual epilogue, no memset.

    pre-LN (mBART):  x is the residual stream, `a` = LN(x) feeds each sub-block.
    post-LN (BART, the flax_vit_bart variant): h = LN(h + sublayer(h)) — the SAME fused kernel applied to
    (cur, other): other = LN(cur + acc + bias); the normalised `other` then IS the residual stream, so the two
    buffers swap roles after every sub-block; BART has no final layer_norm."""
    t, ps, b = engine.t, engine.ps, engine.bufs
    R, d, H, T = cache.rows, t.d_model, t.decoder_attention_heads, cache.T
    S = engine.c.num_tokens
    L = t.decoder_layers
    eps = t.layer_norm_eps
    scale = 1.0 / math.sqrt(t.head_dim)
    x = b.get("gen.x", (R, d))
    a = b.get("gen.a", (R, d))
    q = b.get("gen.q", (R, d))
    o = b.get("gen.o", (R, d))
    g = b.get("gen.g", (R, t.decoder_ffn_dim))
    acc = b.zeros("gen.acc", (R, d))                                 # zeroed once; kernels hand it back zeroed
    sk_d = max(1, min(4, d // 256))
    sk_f = max(1, min(8, t.decoder_ffn_dim // 512))
    L2 = L * 2 * d
    pre = t.pre_layernorm
    # embedding (+ positions) -> x ; every row sits at the same position: pos_mod=1 -> 0 + (pos + offset)
    ops.embed_ln_fwd(tokens, None, 1, pos + t.position_offset, ps.w("shared"), ps.w("d.pos"), engine.emb_scale,
                     ps.f("d.ln_emb.scale"), ps.f("d.ln_emb.bias"), eps, None, x)
    if pre:
        engine._ln_fwd(x, "d.0.ln_sa", eps, a)
        inp = a
    else:
        inp = x                                                      # post-LN: sub-blocks read the stream itself
    cur, other = x, a
    for l in range(L):
        n = f"d.{l}"

        def close(bias_name, ln_name):
            """Fold the sub-block's output (acc + bias) into the stream and normalise.  Returns the next input."""
            nonlocal cur, other
            ops.residual_ln_fwd(acc, ps.f(bias_name), cur, ps.f(ln_name + ".scale"), ps.f(ln_name + ".bias"), eps, other)
            if pre:
                return other                                         # cur stays the stream, other = LN(cur)
            cur, other = other, cur                                  # post-LN: the normalised tensor is the stream
            return cur
        wqkv, bqkv = ps.w(n + ".sa_qkv.w"), ps.f(n + ".sa_qkv.b")
        ops.gemm(inp, wqkv[:, :d], b_mn=True, bias=bqkv[:d], out=q)
        kv_slot = cache.self_kv[l, :, pos, :]                      # [R, 2d] view, row pitch T*2d: written in place
        ops.gemm(inp, wqkv[:, d:], b_mn=True, bias=bqkv[d:], out=kv_slot)
        kc = cache.self_kv[l].view(R * T, 2 * d)
        ops.decode_attention(q, kc[:, :d], kc[:, d:], 2 * d, cache.ancestors, T, pos + 1, 1, o, R, H, scale)
        ops.gemm(o, ps.w(n + ".sa_o.w"), b_mn=True, out=acc, accumulate=True, split_k=sk_d, block_n=64)
        inp = close(n + ".sa_o.b", n + (".ln_ca" if pre else ".ln_sa"))
        ops.gemm(inp, ps.w(n + ".ca_q.w"), b_mn=True, bias=ps.f(n + ".ca_q.b"), out=q)
        ek = cache.enc_kv[:, l * 2 * d: l * 2 * d + d]
        ev = cache.enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]
        ops.decode_attention(q, ek, ev, L2, None, S, S, cache.rows_per_image, o, R, H, scale)
        ops.gemm(o, ps.w(n + ".ca_o.w"), b_mn=True, out=acc, accumulate=True, split_k=sk_d, block_n=64)
        inp = close(n + ".ca_o.b", n + (".ln_f" if pre else ".ln_ca"))
        ops.gemm(inp, ps.w(n + ".fc1.w"), b_mn=True, bias=ps.f(n + ".fc1.b"), act=t.activation_function, out=g)
        ops.gemm(g, ps.w(n + ".fc2.w"), b_mn=True, out=acc, accumulate=True, split_k=sk_f, block_n=64)
        if pre:
            nxt = f"d.{l + 1}.ln_sa" if l + 1 < L else ("d.ln_final" if t.final_layer_norm else None)
            if nxt is None:
                raise NotImplementedError("pre-LN decoder without a final LayerNorm is not wired for cached decode")
            inp = close(n + ".fc2.b", nxt)
        else:
            inp = close(n + ".fc2.b", n + ".ln_f")
    if not pre and t.final_layer_norm:
        raise NotImplementedError("post-LN decoder with a final LayerNorm is not wired for cached decode")
    return inp


def fused_cache_rowmajor(engine, cache: DecodeCache):
    """The self-attention cache of the FUSED path, converted to the per-op path's [layer, row, pos, k|v] layout
    (tests / debugging).  The persistent kernel keeps it head-major — [row][head][K plane | V plane][pos][64], the
    16-byte chunk c of a (pos) row stored at c ^ (pos & 7) — so that one beam's K (or V) history of a head is a
    few contiguous runs that bulk copies fetch (csrc/decode_device.cuh: decode_self_attn_runs)."""
    t = engine.t
    L, R, T, H = t.decoder_layers, cache.rows, cache.T, t.decoder_attention_heads
    kv = cache.self_kv.reshape(L, R, H, 2, T, 8, 8)
    pos = torch.arange(T, device=kv.device)
    src = (torch.arange(8, device=kv.device)[None, :] ^ (pos[:, None] & 7))          # [T, 8]: stored chunk of logical c
    idx = src[None, None, None, None, :, :, None].expand(L, R, H, 2, T, 8, 8)
    logical = torch.gather(kv, 5, idx).reshape(L, R, H, 2, T, 64)
    return logical.permute(0, 1, 4, 3, 2, 5).reshape(L, R, T, 2 * H * 64).contiguous()


def _aligned(engine, name, nbytes, align=1024):
    """Cached uint8 device buffer with an aligned start (tile-image buffers are bulk-copy sources)."""
    raw = engine.bufs.get(name, (nbytes + align,), torch.uint8)
    off = (-raw.data_ptr()) % align
    return raw[off:off + nbytes]


def _fused_plan(engine, cache: DecodeCache, tiled_out=False):
    """Phase table of the persistent decoder-step kernel for this buffer set (built once per buffer set).
    tiled_out: the final hidden states are written as a tile image for the packed lm_head search."""
    t, ps, b = engine.t, engine.ps, engine.bufs
    R, d, T, L = cache.rows, t.d_model, cache.T, t.decoder_layers
    F = t.decoder_ffn_dim
    Rp = (R + 127) // 128 * 128
    bufs = {"x": b.get("gen.x", (R, d)), "q": b.get("gen.q", (R, d)), "h_out": b.get("gen.a", (R, d)),
            "a_tiles": _aligned(engine, "gen.a_tiles", Rp * d * 2), "o_tiles": _aligned(engine, "gen.o_tiles", Rp * d * 2),
            "g_tiles": _aligned(engine, "gen.g_tiles", Rp * F * 2), "ancestors": cache.ancestors,
            "ln_out_g": ps.f("d.ln_final.scale"), "ln_out_b": ps.f("d.ln_final.bias"),
            "h_out_tiles": _aligned(engine, "gen.h_tiles", Rp * d * 2) if tiled_out else None,
            "cross_kv_tiles": _aligned(engine, "gen.xkv_tiles", ops.decoder_cross_kv_tiles_bytes(
                R // cache.rows_per_image, L, t.decoder_attention_heads))}
    for n in ("acc", "q_acc"):
        bufs[n] = b.zeros("gen." + n, (R, d))                        # zeroed once; the kernel hands it back zeroed
    packed = _aligned(engine, "gen.wpack", ops.decoder_packed_bytes(L, d, F))
    key = (R, T, tiled_out, cache.rows_per_image, cache.enc_kv.data_ptr(), cache.self_kv.data_ptr(), ps.shadow.data_ptr(),
           ps.master.data_ptr(), packed.data_ptr()) + tuple(0 if v is None else v.data_ptr() for v in bufs.values())
    # every plan ever built stays alive: captured generate() graphs replay on its device buffers
    plans = engine.__dict__.setdefault("_fused_plans", {})
    entry = plans.get(key)
    if entry is not None:
        plans["decoder"] = (key, entry)
        return entry
    assert not torch.cuda.is_current_stream_capturing(), "fused decoder plan must be built before graph capture"
    layers = []
    for l in range(L):
        n = f"d.{l}"
        layers.append({
            "ln_sa_g": ps.f(n + ".ln_sa.scale"), "ln_sa_b": ps.f(n + ".ln_sa.bias"),
            "sa_qkv_w": ps.w(n + ".sa_qkv.w"), "sa_qkv_b": ps.f(n + ".sa_qkv.b"),
            "sa_o_w": ps.w(n + ".sa_o.w"), "sa_o_b": ps.f(n + ".sa_o.b"),
            "ln_ca_g": ps.f(n + ".ln_ca.scale"), "ln_ca_b": ps.f(n + ".ln_ca.bias"),
            "ca_q_w": ps.w(n + ".ca_q.w"), "ca_q_b": ps.f(n + ".ca_q.b"),
            "ca_o_w": ps.w(n + ".ca_o.w"), "ca_o_b": ps.f(n + ".ca_o.b"),
            "ln_f_g": ps.f(n + ".ln_f.scale"), "ln_f_b": ps.f(n + ".ln_f.bias"),
            "fc1_w": ps.w(n + ".fc1.w"), "fc1_b": ps.f(n + ".fc1.b"),
            "fc2_w": ps.w(n + ".fc2.w"), "fc2_b": ps.f(n + ".fc2.b"),
            "self_kv": cache.self_kv[l],
            "enc_k": cache.enc_kv[:, l * 2 * d: l * 2 * d + d], "enc_v": cache.enc_kv[:, l * 2 * d + d: (l + 1) * 2 * d]})
    lstruct = ops.decoder_layers_struct(layers)
    plan = torch.empty(ops.decoder_plan_bytes(L) + 128, dtype=torch.uint8, device=engine.dev)
    plan = plan[(-plan.data_ptr()) % 128:]
    sync = torch.zeros(256, dtype=I32, device=engine.dev)
    for n in ("a_tiles", "o_tiles", "g_tiles", "h_out_tiles"):
        if bufs[n] is not None:
            bufs[n].zero_()                                   # rows beyond R of the last row tile stay finite
    ops.decoder_plan_init(plan, lstruct, bufs, packed, R, d, t.decoder_attention_heads, F, T, engine.c.num_tokens,
                          cache.rows_per_image, L * 2 * d, t.activation_function, t.layer_norm_eps)
    fp = {"plan": plan, "sync": sync, "bufs": bufs, "layers": lstruct, "packed": packed}
    plans[key] = fp
    plans["decoder"] = (key, fp)           # most recent (tests / tools)
    return fp


def fused_prepare(engine, cache: DecodeCache, packed_search=False):
    """Once per generate() call: (re)build the plan if the buffers moved and re-pack the current weights
    (and, for the packed lm_head search, the tied embedding table as 256-row tile images)."""
    t = engine.t
    fp = _fused_plan(engine, cache, tiled_out=packed_search)
    ops.decoder_pack_weights(fp["layers"], t.d_model, t.decoder_ffn_dim, fp["packed"])
    ops.decoder_pack_cross_kv(cache.enc_kv, cache.rows // cache.rows_per_image, engine.c.num_tokens, t.decoder_layers,
                              t.decoder_attention_heads, t.d_model, fp["bufs"]["cross_kv_tiles"])
    if packed_search:
        emb = engine.ps.w("shared")
        fp["e_tiles"] = _aligned(engine, "gen.e_tiles", ops.pack_kmajor_tiles_bytes(emb.shape[0], emb.shape[1], 256))
        ops.pack_kmajor_tiles(emb, 256, fp["e_tiles"])
    return fp


def decode_step_fused(engine, cache: DecodeCache, tokens, pos: int, fp=None, active=None):
    """decode_step with all decoder layers in ONE persistent kernel (csrc/decoder_step.cu)."""
    t, ps = engine.t, engine.ps
    assert t.pre_layernorm and t.final_layer_norm, "fused cached decode is wired for the pre-LN (mBART) decoder"
    if fp is None:
        fp = fused_prepare(engine, cache)
    ops.embed_ln_fwd(tokens, None, 1, pos + t.position_offset, ps.w("shared"), ps.w("d.pos"), engine.emb_scale,
                     ps.f("d.ln_emb.scale"), ps.f("d.ln_emb.bias"), t.layer_norm_eps, None, fp["bufs"]["x"])
    ops.decoder_step(fp["plan"], t.decoder_layers, cache.rows, pos, fp["sync"], active=active)
    return fp["bufs"]["h_out_tiles"] if fp["bufs"]["h_out_tiles"] is not None else fp["bufs"]["h_out"]


def _search_ws(engine, R, cand_per_row=8):
    b, V = engine.bufs, engine.t.vocab_size
    n = ops.lm_head_search_num_partials(R)
    return {"nparts": n, "pmax": b.get("gen.pmax", (n, R), F32), "psum": b.get("gen.psum", (n, R), F32),
            "cand_val": b.get("gen.cand_val", (n, R, 8), F32), "cand_idx": b.get("gen.cand_idx", (n, R, 8), I32),
            "row_lp": b.get("gen.row_lp", (R, cand_per_row), F32), "row_tok": b.get("gen.row_tok", (R, cand_per_row), I32),
            "row_ml": b.get("gen.row_ml", (R, 2), F32),
            "last_val": b.get("gen.last_val", (R,), F32), "last_idx": b.get("gen.last_idx", (R,), I32)}


# ---- jax.random key handling for `_sample` (host side; the per-element stream is generated in the search kernel) ----
def _threefry2x32_host(k0, k1, x0, x1):
    """threefry2x32, 20 rounds, on Python ints (jax._src.random at the pinned jax==0.2.16, [MEMORY] risk U9)."""
    m = 0xFFFFFFFF
    ks = (k0 & m, k1 & m, (k0 ^ k1 ^ 0x1BD11BDA) & m)
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    x0, x1 = (x0 + ks[0]) & m, (x1 + ks[1]) & m
    for g in range(5):
        for r in rot[g & 1]:
            x0 = (x0 + x1) & m
            x1 = ((x1 << r) | (x1 >> (32 - r))) & m
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & m
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & m
    return x0, x1


def prng_key_pair(prng_key=None):
    """A jax PRNG key (uint32[2]) / int seed / None (= PRNGKey(0), generation_clip_vision_utils.py:565) -> (k0, k1)."""
    if prng_key is None:
        return (0, 0)
    if isinstance(prng_key, int):
        return ((prng_key >> 32) & 0xFFFFFFFF, prng_key & 0xFFFFFFFF)
    if hasattr(prng_key, "detach"):
        prng_key = prng_key.detach().cpu().numpy()
    k = [int(x) & 0xFFFFFFFF for x in list(prng_key)]
    assert len(k) == 2, "a PRNG key is two uint32 words"
    return (k[0], k[1])


def prng_split(key):
    """jax.random.split(key) -> (first, second): threefry_2x32(key, iota(4)) with counts split into halves."""
    a0, a1 = _threefry2x32_host(key[0], key[1], 0, 2)
    b0, b1 = _threefry2x32_host(key[0], key[1], 1, 3)
    return (a0, b0), (a1, b1)


def _min_length_applies(cur_len, min_length):
    """FlaxMinLengthLogitsProcessor (transformers@0085e712 generation_flax_logits_process.py, risk U4 [MEMORY]):
    `apply_penalty = 1 - clip(cur_len - min_length, 0, 1)` — EOS is masked while cur_len <= min_length, one step
    longer than the PyTorch processor's `cur_len < min_length`."""
    return 1 - min(max(cur_len - min_length, 0), 1) == 1


def _forced_token(cur_len, max_length, forced_bos, forced_eos):
    """FlaxForcedBOS (cur_len == 1) then FlaxForcedEOS (cur_len == max_length-1): the later processor wins."""
    f = -1
    if forced_bos is not None and cur_len == 1:
        f = forced_bos
    if forced_eos is not None and cur_len == max_length - 1:
        f = forced_eos
    return f


def _step(engine, cache, tokens, pos, active=None):
    """`active` = the device-side while_loop condition: once it is 0 the persistent decoder step and the lm_head
    search return immediately (the reference's loop would have ended, generation_clip_vision_utils.py:798-820)."""
    if cache.fused is not None:
        return decode_step_fused(engine, cache, tokens, pos, cache.fused, active=active)
    return decode_step(engine, cache, tokens, pos)


def _lm_head_search(engine, cache, hf, mask_token, ws, active=None, gumbel_key=None):
    """lm_head + log-softmax partials + per-row candidates.  8 candidates per row cover 2*num_beams for <= 4 beams;
    5..8 beams run the search a second time restricted to what ranks after the first pass's 8th (exact: both passes
    compute bit-identical logits)."""
    ps, t = engine.ps, engine.t
    passes = 2 if ws["row_lp"].shape[1] > 8 else 1
    for i in range(passes):
        if cache.fused is not None and "e_tiles" in cache.fused:
            ops.lm_head_search_packed(hf, cache.fused["e_tiles"], ps.f("flb"), int(mask_token), cache.rows, t.vocab_size,
                                      t.d_model, ws, second_pass=i == 1, active=active, gumbel_key=gumbel_key)
        else:
            ops.lm_head_search(hf, ps.w("shared"), ps.f("flb"), mask_token, ws, second_pass=i == 1, active=active,
                               gumbel_key=gumbel_key)
        ops.search_merge(ws, cache.rows, second_pass=i == 1)


def _search_loop(engine, px, *, max_length, pad_token_id, eos_token_id, decoder_start_token_id, num_beams, min_length,
                 forced_bos_token_id, forced_eos_token_id, length_penalty, early_stopping, trace_cb=None,
                 sample_key=None):
    """Enqueue encode + the whole search loop on the current stream (no host synchronisation inside:
    the while_loop condition lives in the device flag `active`).  Capturable into one CUDA graph.
    trace_cb(cur_len, ws, st) — eager runs only — is called after the lm_head search of every un-forced step, before
    the bookkeeping kernel consumes `ws` (parity tests read the candidate lists and the search state there)."""
    t, ps = engine.t, engine.ps
    dev = engine.dev
    B = px.shape[0]
    K, Lmax, V = num_beams, max_length, t.vocab_size
    R = B * K
    enc = engine.encode(px, trunc_int=True, save=False, tag="gen.enc")
    enc_kv = engine.cross_kv(enc, tag="gen.enc")
    cache = DecodeCache(engine, R, Lmax, enc_kv, K, use_ancestors=K > 1)
    # the persistent decoder-step kernel stages <= 64 keys per attention item and <= 4 beams per image; longer
    # searches (the model default max_length is 200) take the per-op path
    fused_ok = (Lmax <= 64 and engine.c.num_tokens <= 64 and K <= 8 and t.pre_layernorm and t.final_layer_norm
                and t.activation_function == "gelu" and t.decoder_layers <= 12)
    import os
    if getattr(engine, "fused_decoder", True) and fused_ok and os.environ.get("MIC_FUSED_DECODER", "1") != "0":
        cache.fused = fused_prepare(engine, cache, packed_search=os.environ.get("MIC_PACKED_SEARCH", "1") != "0")
    ws = _search_ws(engine, R, 8 if K <= 4 else 16)
    active = torch.ones(1, dtype=I32, device=dev)
    next_token = torch.full((R,), decoder_start_token_id, dtype=I32, device=dev)
    mask_eos = min_length is not None and eos_token_id is not None and min_length > -1

    if K == 1:
        st = {"sequences": torch.full((R, Lmax), pad_token_id, dtype=I32, device=dev),
              "finished": torch.zeros(R, dtype=I32, device=dev), "next_token": next_token, "active": active}
        st["sequences"][:, 0] = decoder_start_token_id
        key = sample_key
        for cur_len in range(1, Lmax):
            forced = _forced_token(cur_len, Lmax, forced_bos_token_id, forced_eos_token_id)
            last = cur_len == Lmax - 1
            gk = None
            if sample_key is not None:
                # `_sample` :616-625: split the key, draw from the RAW logits (processors / warpers do not act)
                gk, key = prng_split(key)
                forced = -1
            if not (last and forced >= 0):
                hf = _step(engine, cache, st["next_token"], cur_len - 1, active)
            if forced < 0:
                mt = eos_token_id if (mask_eos and _min_length_applies(cur_len, min_length)) else -1
                _lm_head_search(engine, cache, hf, mt, ws, active, gumbel_key=gk)
                if trace_cb is not None:
                    trace_cb(cur_len, ws, st)
            ops.greedy_step(ws, st, forced, R, Lmax, cur_len, eos_token_id, pad_token_id)
            ops.greedy_cond(st, R, cur_len + 1, Lmax)
        return {"sequences": st["sequences"]}

    st = {"running_seq": torch.full((B, K, Lmax), pad_token_id, dtype=I32, device=dev),
          "sequences": torch.full((B, K, Lmax), pad_token_id, dtype=I32, device=dev),
          "running_scores": torch.full((B, K), -1.0e7, dtype=F32, device=dev),     # [0, -1e7, ...] set below
          "scores": torch.full((B, K), -1.0e7, dtype=F32, device=dev),
          "finished": torch.zeros((B, K), dtype=I32, device=dev),
          "ancestors": cache.ancestors, "next_token": next_token, "active": active}
    st["running_seq"][:, :, 0] = decoder_start_token_id
    st["running_scores"][:, 0] = 0.0
    for cur_len in range(1, Lmax):
        forced = _forced_token(cur_len, Lmax, forced_bos_token_id, forced_eos_token_id)
        last = cur_len == Lmax - 1
        if not (last and forced >= 0):
            hf = _step(engine, cache, st["next_token"], cur_len - 1, active)
        if forced < 0:
            mt = eos_token_id if (mask_eos and _min_length_applies(cur_len, min_length)) else -1
            _lm_head_search(engine, cache, hf, mt, ws, active)
            if trace_cb is not None:
                trace_cb(cur_len, ws, st)
        ops.beam_step(ws, st, forced, B, K, Lmax, V, cur_len, eos_token_id, early_stopping, length_penalty)
        ops.beam_cond(st, B, K, cur_len + 1, Lmax, length_penalty, early_stopping)
    out_seq = torch.empty((B, Lmax), dtype=I32, device=dev)
    out_scores = torch.empty((B,), dtype=F32, device=dev)
    ops.beam_finalize(st, B, K, Lmax, out_seq, out_scores)
    return {"sequences": out_seq, "scores": out_scores}


@torch.no_grad()
def generate(engine, pixel_values, *, use_cuda_graph=True, pdl=False, prefetch_weights=None, trace_cb=None, **kw):
    """`generate` :128-336.  encode() truncates pixels to int32 first (modeling_clip_vision_mbart.py:330).
    The first call for a given (batch, search settings) runs eagerly (allocates every buffer); the whole
    loop is then captured into ONE CUDA graph and later calls only copy the pixels in and replay it."""
    if kw["num_beams"] > 8:
        raise NotImplementedError("beam search keeps 2*num_beams <= 16 candidates per image row (num_beams <= 8)")
    px = pixel_values.to(engine.dev).contiguous() if pixel_values.dtype == torch.uint8 else \
        pixel_values.to(engine.dev, F32).contiguous()
    # parameters are frozen while the loop runs: GEMMs prefetch weight tiles ahead of their dependency wait
    prefetch_weights = pdl if prefetch_weights is None else prefetch_weights
    ops.launch_options(pdl=int(pdl), gemm_b_static=int(prefetch_weights))
    if trace_cb is not None:
        use_cuda_graph = False
        kw = dict(kw, trace_cb=trace_cb)
    if kw.get("sample_key") is not None:
        use_cuda_graph = False                  # the per-step keys are launch arguments: a replay would repeat them
    try:
        return _generate(engine, px, use_cuda_graph, (pdl, prefetch_weights), kw)
    finally:
        ops.launch_options(pdl=0, gemm_b_static=0)


def _generate(engine, px, use_cuda_graph, pdl, kw):
    if not use_cuda_graph:
        return _search_loop(engine, px, **kw)
    key = (tuple(px.shape), px.dtype, pdl) + tuple(sorted(kw.items()))
    graphs = engine.__dict__.setdefault("_gen_graphs", {})
    entry = graphs.get(key)
    if entry is None:
        static_px = px.clone()
        out = _search_loop(engine, static_px, **kw)          # eager warm-up: result of this call
        result = {k: v.clone() for k, v in out.items()}
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            gout = _search_loop(engine, static_px, **kw)
        graphs[key] = (g, static_px, gout)
        return result
    g, static_px, gout = entry
    static_px.copy_(px)
    g.replay()
    return {k: v.clone() for k, v in gout.items()}
