"""ORACLE — test infrastructure only.  PARITY UNPINNED BY THE REFERENCE (no tests/goldens exist).

CPU restatement of the vendored generation code
`models/flax_clip_vision_mbart/generation_clip_vision_utils.py` (greedy `:422-535`, beam search
`:665-990`, processors `:368-420`) and of the cached decode path it drives
(`modeling_clip_vision_mbart.py:249-282,519-693`).  fp32 arithmetic is kept in numpy float32 exactly
where the reference computes in float32 (scores, -1e7 penalties, length penalty).

Upstream pieces restated from their published algorithms (transformers@0085e712,
`generation_flax_logits_process.py`; jax 0.2.16 `lax.top_k`):
  * FlaxMinLength / ForcedBOS / ForcedEOS processors (risk U4)
  * lax.top_k: values descending, ties -> lower index first (risk U5)
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import reference_model as rm

NEG = np.float32(-1.0e7)


# ----------------------------------------------------------------------------------------------
def top_k(x: np.ndarray, k: int):
    """lax.top_k along the last axis: descending, stable (lower index wins ties, incl. -inf ties)."""
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(x, idx, axis=-1), idx


def log_softmax(x: np.ndarray) -> np.ndarray:
    """jax.nn.log_softmax in float32: x - max - log(sum(exp(x - max)))."""
    x = x.astype(np.float32)
    s = x - x.max(axis=-1, keepdims=True)
    return (s - np.log(np.exp(s).sum(axis=-1, keepdims=True, dtype=np.float32))).astype(np.float32)


def apply_processors(scores: np.ndarray, cur_len: int, *, min_length, eos_token_id, forced_bos_token_id,
                     forced_eos_token_id, max_length) -> np.ndarray:
    """`_get_logits_processor` (:368-420) order: MinLength, ForcedBOS, ForcedEOS.
    FlaxMinLength: apply_penalty = 1 - clip(cur_len - min_length, 0, 1) -> scores[:, eos] = -inf while
                   cur_len <= min_length (risk U4 [MEMORY]: one step longer than the PyTorch processor)
    FlaxForcedBOS: at cur_len == 1 -> all -inf except forced id = 0
    FlaxForcedEOS: at cur_len == max_length - 1 -> all -inf except forced id = 0"""
    s = scores
    if min_length is not None and eos_token_id is not None and min_length > -1:
        if 1 - int(np.clip(cur_len - min_length, 0, 1)):
            s = s.copy()
            s[:, eos_token_id] = -np.inf
    if forced_bos_token_id is not None and cur_len == 1:
        s = np.full_like(s, -np.inf)
        s[:, forced_bos_token_id] = 0.0
    if forced_eos_token_id is not None and cur_len == max_length - 1:
        s = np.full_like(s, -np.inf)
        s[:, forced_eos_token_id] = 0.0
    return s


# ----------------------------------------------------------------------------------------------
class DecodeCache:
    """init_cache (:249-282) + FlaxMBartAttention cache semantics.
    Mathematically: step t attends over cached keys 0..t (the all-ones static mask at :669 and the
    causal row make slots > t invisible), so the cache is kept as a growing list."""

    def __init__(self, n_layers):
        self.k = [None] * n_layers
        self.v = [None] * n_layers
        self.cross = [None] * n_layers
        self.index = 0

    def gather(self, rows):
        rows = torch.as_tensor(rows, dtype=torch.long)
        for i in range(len(self.k)):
            if self.k[i] is not None:
                self.k[i] = self.k[i][rows]
                self.v[i] = self.v[i][rows]


def decode_step(params, token, position, enc, cache: DecodeCache, config):
    """`decode` (:519-651) for one token per row: returns logits (R, V) and updates the cache.
    Cross-attention K/V are a pure function of `enc`, so they are projected once and memoised
    (the reference re-projects each step; results identical — SURVEY.md risk U6)."""
    t = config.mbart_config
    H, hd = t.decoder_attention_heads, t.head_dim
    dp = params["model"]["decoder"]
    ids = torch.as_tensor(token, dtype=torch.long).reshape(-1, 1)
    R = ids.shape[0]
    pos = torch.full((R, 1), int(position), dtype=torch.long)
    h = rm._embed(params, ids, pos, config)

    def self_attn(x, lp, i):
        q = rm.dense(x, lp["q_proj"]).view(R, 1, H, hd) / math.sqrt(hd)
        k = rm.dense(x, lp["k_proj"]).view(R, 1, H, hd)
        v = rm.dense(x, lp["v_proj"]).view(R, 1, H, hd)
        cache.k[i] = k if cache.k[i] is None else torch.cat([cache.k[i], k], 1)
        cache.v[i] = v if cache.v[i] is None else torch.cat([cache.v[i], v], 1)
        w = torch.softmax(torch.einsum("bqhd,bkhd->bhqk", q, cache.k[i]), -1)
        o = torch.einsum("bhqk,bkhd->bqhd", w, cache.v[i]).reshape(R, 1, H * hd)
        return rm.dense(o, lp["out_proj"])

    def cross_attn(x, lp, i):
        if cache.cross[i] is None:
            S = enc.shape[1]
            cache.cross[i] = (rm.dense(enc, lp["k_proj"]).view(R, S, H, hd),
                              rm.dense(enc, lp["v_proj"]).view(R, S, H, hd))
        ck, cv = cache.cross[i]
        q = rm.dense(x, lp["q_proj"]).view(R, 1, H, hd) / math.sqrt(hd)
        w = torch.softmax(torch.einsum("bqhd,bkhd->bhqk", q, ck), -1)
        o = torch.einsum("bhqk,bkhd->bqhd", w, cv).reshape(R, 1, H * hd)
        return rm.dense(o, lp["out_proj"])

    eps, act = t.layer_norm_eps, t.activation_function
    for i in range(t.decoder_layers):
        lp = dp["layers"][str(i)]
        if t.pre_layernorm:
            h = h + self_attn(rm.layer_norm(h, lp["self_attn_layer_norm"], eps), lp["self_attn"], i)
            h = h + cross_attn(rm.layer_norm(h, lp["encoder_attn_layer_norm"], eps), lp["encoder_attn"], i)
            f = rm.layer_norm(h, lp["final_layer_norm"], eps)
            h = h + rm.dense(rm.activation(rm.dense(f, lp["fc1"]), act), lp["fc2"])
        else:
            h = rm.layer_norm(h + self_attn(h, lp["self_attn"], i), lp["self_attn_layer_norm"], eps)
            h = rm.layer_norm(h + cross_attn(h, lp["encoder_attn"], i), lp["encoder_attn_layer_norm"], eps)
            h = rm.layer_norm(h + rm.dense(rm.activation(rm.dense(h, lp["fc1"]), act), lp["fc2"]),
                              lp["final_layer_norm"], eps)
    if t.final_layer_norm:
        h = rm.layer_norm(h, dp["layer_norm"], eps)
    cache.index += 1
    return rm.lm_head(params, h)[:, 0].numpy().astype(np.float32)


def _gen_defaults(config, max_length, pad_token_id, eos_token_id, decoder_start_token_id, num_beams,
                  min_length, forced_bos_token_id, forced_eos_token_id, length_penalty, early_stopping):
    t = config.mbart_config
    d = dict(
        max_length=max_length if max_length is not None else t.max_length,
        pad_token_id=pad_token_id if pad_token_id is not None else t.pad_token_id,
        eos_token_id=eos_token_id if eos_token_id is not None else t.eos_token_id,
        # `if decoder_start_token_id` (truthiness) at generation_clip_vision_utils.py:216-220
        decoder_start_token_id=decoder_start_token_id if decoder_start_token_id else t.decoder_start_token_id,
        num_beams=num_beams if num_beams is not None else t.num_beams,
        min_length=min_length if min_length is not None else t.min_length,
        forced_bos_token_id=forced_bos_token_id if forced_bos_token_id is not None else t.forced_bos_token_id,
        forced_eos_token_id=forced_eos_token_id if forced_eos_token_id is not None else t.forced_eos_token_id,
        length_penalty=length_penalty if length_penalty is not None else t.length_penalty,
        early_stopping=early_stopping if early_stopping is not None else t.early_stopping)
    return d


@torch.no_grad()
def generate(params, pixel_values, config, max_length=None, pad_token_id=None, eos_token_id=None,
             decoder_start_token_id=None, num_beams=None, min_length=None, forced_bos_token_id=None,
             forced_eos_token_id=None, length_penalty=None, early_stopping=None, return_trace=False,
             do_sample=False, prng_key=None):
    """`generate` (:128-336): encode once (with the int32 pixel cast of encode(), modeling...:330),
    then greedy (num_beams == 1) or beam search."""
    g = _gen_defaults(config, max_length, pad_token_id, eos_token_id, decoder_start_token_id, num_beams,
                      min_length, forced_bos_token_id, forced_eos_token_id, length_penalty, early_stopping)
    p = rm.to_torch_tree(params)
    enc = rm.encode(p, pixel_values, config, int32_cast=True)
    if do_sample:
        if g["num_beams"] != 1:
            raise NotImplementedError("`Beam sampling is currently not implemented.")
        return _sample(p, enc, config, g, prng_key, return_trace)
    if g["num_beams"] == 1:
        return _greedy_search(p, enc, config, g, return_trace)
    return _beam_search(p, enc, config, g, return_trace)


# ----------------------------------------------------------------------------------------------
# jax._src.random at jax==0.2.16 (requirements.txt:13), restated from the published algorithm [MEMORY, risk U9]:
# PRNGKey(seed) = uint32[seed >> 32, seed & 0xffffffff]; threefry2x32 (20 rounds); split; random_bits; uniform; gumbel;
# categorical(key, logits) = argmax(logits + gumbel(key, logits.shape)).
# ----------------------------------------------------------------------------------------------
def _threefry2x32(k0, k1, x0, x1):
    k0, k1 = np.uint32(k0), np.uint32(k1)
    x0, x1 = x0.astype(np.uint32).copy(), x1.astype(np.uint32).copy()
    ks = [k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA))]
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for g in range(5):
            for r in rot[g & 1]:
                x0 += x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 ^= x0
            x0 += ks[(g + 1) % 3]
            x1 += ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def threefry_2x32(key, count):
    """jax `threefry_2x32(keypair, count)`: counts split into halves (odd sizes padded with one 0)."""
    c = np.asarray(count, dtype=np.uint32).ravel()
    odd = c.size % 2
    if odd:
        c = np.concatenate([c, np.zeros(1, np.uint32)])
    h = c.size // 2
    y0, y1 = _threefry2x32(key[0], key[1], c[:h], c[h:])
    out = np.concatenate([y0, y1])
    return (out[:-1] if odd else out).reshape(np.shape(count))


def prng_key(seed=0):
    return np.array([(int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def prng_split(key, num=2):
    return threefry_2x32(key, np.arange(num * 2, dtype=np.uint32)).reshape(num, 2)


def gumbel(key, shape):
    size = int(np.prod(shape))
    bits = threefry_2x32(key, np.arange(size, dtype=np.uint32)).reshape(shape)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    tiny = np.finfo(np.float32).tiny
    u = np.maximum(tiny, f * (np.float32(1.0) - tiny) + tiny).astype(np.float32)
    return (-np.log(-np.log(u))).astype(np.float32)


def categorical(key, logits):
    return np.argmax(gumbel(key, logits.shape) + logits.astype(np.float32), axis=-1)


def _sample(p, enc, config, g, key, return_trace=False):
    """`_sample` (:537-663).  Quirk reproduced: the processed / warped logits are computed and then IGNORED — the
    token is drawn from the RAW logits (`jax.random.categorical(prng_key, model_outputs.logits[:, -1])`, :623-625),
    so forced BOS/EOS, min_length, top-k/top-p and temperature have no effect."""
    B = enc.shape[0]
    L, pad, eos = g["max_length"], g["pad_token_id"], g["eos_token_id"]
    sequences = np.full((B, L), pad, dtype=np.int32)
    sequences[:, 0] = g["decoder_start_token_id"]
    finished = np.zeros((B,), dtype=bool)
    running = sequences[:, 0].copy()
    cache = DecodeCache(config.mbart_config.decoder_layers)
    key = prng_key(0) if key is None else np.asarray(key, dtype=np.uint32)
    cur_len, pos, margins = 1, 0, []
    while not (cur_len == L or finished.all()):
        k_use, key = prng_split(key)                                          # :616
        logits = decode_step(p, running, pos, enc, cache, config)
        noisy = gumbel(k_use, logits.shape) + logits
        nxt = np.argmax(noisy, axis=-1).astype(np.int32)
        if return_trace:
            srt = np.sort(noisy, axis=-1)
            margins.append(srt[:, -1] - srt[:, -2])
        finished = finished | (nxt == eos)
        nxt = np.where(finished, pad, nxt).astype(np.int32)
        sequences[:, cur_len] = nxt
        running = nxt
        cur_len += 1
        pos += 1
    out = {"sequences": sequences}
    if return_trace:
        out["margins"] = np.stack(margins, 1) if margins else np.zeros((B, 0), np.float32)
    return out


def _greedy_search(p, enc, config, g, return_trace=False):
    """`_greedy_search` (:422-535)."""
    B = enc.shape[0]
    L, pad, eos = g["max_length"], g["pad_token_id"], g["eos_token_id"]
    sequences = np.full((B, L), pad, dtype=np.int32)
    sequences[:, 0] = g["decoder_start_token_id"]
    finished = np.zeros((B,), dtype=bool)
    running = sequences[:, 0].copy()
    cache = DecodeCache(config.mbart_config.decoder_layers)
    cur_len, pos, margins = 1, 0, []
    while not (cur_len == L or finished.all()):                       # :480-487
        logits = decode_step(p, running, pos, enc, cache, config)     # raw logits (:497)
        logits = apply_processors(logits, cur_len, min_length=g["min_length"], eos_token_id=eos,
                                  forced_bos_token_id=g["forced_bos_token_id"],
                                  forced_eos_token_id=g["forced_eos_token_id"], max_length=L)
        nxt = logits.argmax(-1).astype(np.int32)                      # first max on ties (:499)
        if return_trace:
            srt = np.sort(logits, axis=-1)
            margins.append(srt[:, -1] - srt[:, -2])
        finished = finished | (nxt == eos)                            # :501-503
        nxt = np.where(finished, pad, nxt).astype(np.int32)           # :504-507
        sequences[:, cur_len] = nxt
        running = nxt
        cur_len += 1
        pos += 1
    out = {"sequences": sequences}
    if return_trace:
        out["margins"] = np.stack(margins, 1) if margins else np.zeros((B, 0), np.float32)
    return out


def _beam_search(p, enc, config, g, return_trace=False, logits_fn=None, batch_size=None):
    """`_beam_search` (:665-990), statement by statement; arrays are (batch, beams, ...)."""
    # logits_fn(cur_len, running_sequences[B*K, L]) -> logits (B*K, V): test hook that replaces the decoder so
    # the bookkeeping can be compared bit-for-bit with the CUDA kernels on identical log-probs
    B = enc.shape[0] if logits_fn is None else batch_size
    K, L, pad, eos = g["num_beams"], g["max_length"], g["pad_token_id"], g["eos_token_id"]
    lp, early = g["length_penalty"], bool(g["early_stopping"])
    V = config.mbart_config.vocab_size
    enc_rows = None if logits_fn is not None else \
        enc[:, None].expand(B, K, *enc.shape[1:]).reshape(B * K, *enc.shape[1:])          # :299-307, flatten
    sequences = np.full((B, K, L), pad, dtype=np.int32)
    running_sequences = np.full((B, K, L), pad, dtype=np.int32)
    running_sequences[:, :, 0] = g["decoder_start_token_id"]
    initial_running_flat = running_sequences.reshape(B * K, L).copy()   # closure captured at :852
    is_sent_finished = np.zeros((B, K), dtype=bool)
    running_scores = np.tile(np.array([0.0] + [NEG] * (K - 1), dtype=np.float32), (B, 1))
    scores = np.full((B, K), NEG, dtype=np.float32)
    cache = DecodeCache(config.mbart_config.decoder_layers)
    cur_len, pos = 1, 0
    bidx = np.arange(B)[:, None]
    trace = []

    def cond():
        not_max = cur_len < L
        best_running = running_scores[:, -1:] / np.float32(float(L) ** lp)
        worst_finished = np.where(is_sent_finished, scores.min(axis=1, keepdims=True), NEG)
        improvement = bool(np.all(worst_finished < best_running))
        still_open = not (bool(is_sent_finished.all()) and early)
        return not_max and still_open and improvement

    first = True
    while first or cond():
        first = False
        token = running_sequences[:, :, cur_len - 1].reshape(B * K)                  # :830-836
        if logits_fn is not None:
            logits = logits_fn(cur_len, running_sequences.reshape(B * K, L))
        else:
            logits = decode_step(p, token, pos, enc_rows, cache, config)              # (B*K, V)
        log_probs = log_softmax(logits)                                               # :850
        log_probs = apply_processors(log_probs, cur_len, min_length=g["min_length"], eos_token_id=eos,
                                     forced_bos_token_id=g["forced_bos_token_id"],
                                     forced_eos_token_id=g["forced_eos_token_id"], max_length=L)
        log_probs = log_probs.reshape(B, K, V) + running_scores[:, :, None]           # :857
        flat = log_probs.reshape(B, K * V).astype(np.float32)
        topk_log_probs, topk_indices = top_k(flat, 2 * K)                             # :873
        topk_raw = topk_log_probs.copy()
        topk_beam = topk_indices // V
        topk_running = running_sequences[bidx, topk_beam]                             # (B, 2K, L)
        topk_ids = (topk_indices % V).astype(np.int32)
        topk_sequences = topk_running.copy()
        topk_sequences[:, :, cur_len] = topk_ids                                      # :882-885
        did_finish = topk_sequences[:, :, cur_len] == eos                             # :889
        topk_log_probs = (topk_log_probs + did_finish.astype(np.float32) * NEG).astype(np.float32)  # :890
        next_topk_indices = np.flip(top_k(topk_log_probs, K)[1], axis=1)              # :895-897
        next_running_sequences = topk_sequences[bidx, next_topk_indices]
        next_running_scores = topk_log_probs[bidx, next_topk_indices]
        if return_trace:
            rest = np.sort(flat, axis=-1)[:, -(2 * K + 1)]
            trace.append({"cur_len": cur_len, "running_sequences": running_sequences.copy(),
                          "running_scores": running_scores.copy(),
                          "topk_log_probs": topk_log_probs.copy(), "topk_raw": topk_raw,
                          "did_finish": did_finish.copy(),
                          "topk_indices": topk_indices.copy(), "ninth": rest.copy()})
        # :910-919 (re-uses the penalised topk_log_probs)
        topk_log_probs = (topk_log_probs / np.float32(float(cur_len) ** lp)).astype(np.float32)
        beams_full = np.broadcast_to(is_sent_finished.all(axis=-1, keepdims=True), did_finish.shape) & early
        add_penalty = (~did_finish) | beams_full
        topk_log_probs = (topk_log_probs + add_penalty.astype(np.float32) * NEG).astype(np.float32)
        # :925-940
        merged_sequences = np.concatenate([sequences, topk_sequences], axis=1)
        merged_scores = np.concatenate([scores, topk_log_probs], axis=1)
        merged_finished = np.concatenate([is_sent_finished, did_finish], axis=1)
        topk_merged = np.flip(top_k(merged_scores, K)[1], axis=1)
        sequences = merged_sequences[bidx, topk_merged]
        scores = merged_scores[bidx, topk_merged]
        is_sent_finished = merged_finished[bidx, topk_merged]
        # :945-953 cache reorder
        next_running_indices = topk_beam[bidx, next_topk_indices]                     # (B, K) old beam per new beam
        rows = (np.arange(B)[:, None] * K + next_running_indices).reshape(-1)
        cache.gather(rows)
        running_sequences, running_scores = next_running_sequences, next_running_scores
        cur_len += 1
        pos += 1
    none_finished = is_sent_finished.any(axis=1)                                      # :980-984
    out_seq = np.where(none_finished[:, None, None], sequences, running_sequences)
    out_scores = np.where(none_finished[:, None], scores, running_scores)
    out = {"sequences": out_seq[:, -1].astype(np.int32), "scores": out_scores[:, -1].astype(np.float32)}
    if return_trace:
        out["trace"] = trace
        out["all_sequences"] = out_seq
        out["state"] = {"running_seq": running_sequences, "running_scores": running_scores, "sequences": sequences,
                        "scores": scores, "finished": is_sent_finished.astype(np.int32), "cur_len": cur_len}
    _ = initial_running_flat  # processors ignore input_ids; kept to document the :852 closure quirk
    return out
