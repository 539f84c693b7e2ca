"""Full-size parity of the configs the headline metric is quoted on (VERDICT r01 "Next round" item 2), driven through
the product path — captured-graph generate() with the FUSED persistent decoder step, and the bf16 training tape —
against goldens the CPU oracle wrote (tests/golden/gen_golden_full.py; the weights are rebuilt from the numpy seeds).

  * BASELINE configs[3] shape at full size (CLIP-ViT-B/32 + mBART-50, V = 250,054, 12 layers), B = 8: greedy and
    beam-4, max_length 64, forced_bos es_XX, for the random-init set and for the "peaked" set where EOS fires
    naturally.  Token ids are compared EXACTLY wherever the oracle's decision margin exceeds the tolerance written
    below (north-star rule); a divergence is accepted only right after a sub-tolerance margin, and that row is not
    compared further.
  * BASELINE configs[0/1] shape at full size, B = 8: loss and the gradient of EVERY parameter tensor.

PARITY UNPINNED: the oracle restates the reference (it cannot run here) and is pinned against the HF PyTorch twins.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import generation as gen, synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GEN = os.path.join(HERE, "golden", "config3_full_gen_golden.npz")
GRAD = os.path.join(HERE, "golden", "config1_full_grad_golden.npz")
GEN_KW = dict(max_length=64, forced_bos_token_id=250005)
B, K, L, V = 8, 4, 64, 250054

# tolerances (bf16 compute, fp32 accumulation, against the fp32 oracle)
TOL_GREEDY_MARGIN = 0.10     # top-1 minus top-2 logit below which a greedy position is not compared (and ends the row)
TOL_STEP_LOGPROB = 0.08      # per-step log-prob of a kept candidate (increment over its beam's running score)
TOL_BEAM_GAP = 0.30          # cumulative beam-score gap below which a beam decision may legitimately differ
                             # (up to 63 accumulated bf16 steps)

_CACHE = {}


def _model():
    if "m" not in _CACHE:
        _CACHE["cfg"] = mic_b200.clip_mbart_config()
        _CACHE["m"] = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(_CACHE["cfg"], seed=0)
        _CACHE["set"] = None
    return _CACHE["m"], _CACHE["cfg"]


def _use(name):
    model, cfg = _model()
    if _CACHE["set"] != name:
        if name == "init":
            p = synthetic.make_params(cfg, seed=1)
        elif name == "peaked":
            p = synthetic.make_peaked_params(cfg, seed=11)
        else:
            p = synthetic.make_params(cfg, seed=1, perturbed=True)
        model.params = p
        _CACHE["set"] = name
    return model, cfg


def _px(cfg):
    return synthetic.make_batch(cfg, B, 64, seed=7)["pixel_values"]


@pytest.mark.parametrize("name", ["init", "peaked"])
def test_full_size_greedy_tokens_match_golden_where_margin_is_clear(name):
    g = np.load(GEN)
    model, cfg = _use(name)
    assert model.engine.__dict__.get("fused_decoder", True)
    # eager warm-up + capture, then the graph replay: BOTH are held to the golden (they need not equal each other at
    # sub-tolerance margins: the split-K reductions of the decoder step are fp32 atomics, i.e. order-dependent)
    runs = [model.generate(_px(cfg), num_beams=1, **GEN_KW).sequences.cpu().numpy() for _ in range(2)]
    assert "decoder" in model.engine._fused_plans, "the persistent decoder-step kernel must be the path under test"
    ref, margins = g[f"{name}_greedy_seq"], g[f"{name}_greedy_margins"]
    for seq in runs:
        compared = 0
        for b in range(B):
            done = False
            for pos in range(1, L):
                if not done and pos - 1 < margins.shape[1] and margins[b, pos - 1] < TOL_GREEDY_MARGIN:
                    break                                   # sub-tolerance decision: this row is not compared further
                assert seq[b, pos] == ref[b, pos], (name, b, pos, seq[b, :pos + 2], ref[b, :pos + 2])
                compared += 1
                done = done or ref[b, pos] == 1             # after EOS the row is pad (processors no longer matter)
        assert compared >= (8 * 63 if name == "init" else 100), compared


def _cuda_top2k(row_lp, row_tok, running_scores):
    """What beam_step_kernel builds: candidates = row log-probs + the beam's running score, flat id = beam*V + token,
    ordered by (value descending, flat id ascending)."""
    cand = row_lp.reshape(B, K, -1) + running_scores[:, :, None]
    flat = (np.arange(K)[None, :, None] * V + row_tok.reshape(B, K, -1)).astype(np.int64)
    cand, flat = cand.reshape(B, -1), flat.reshape(B, -1)
    order = np.lexsort((flat, -cand), axis=-1)[:, :2 * K]
    return np.take_along_axis(cand, order, 1), np.take_along_axis(flat, order, 1)


@pytest.mark.parametrize("name", ["init", "peaked"])
def test_full_size_beam4_trace_matches_golden(name):
    g = np.load(GEN)
    model, cfg = _use(name)
    cur_lens = g[f"{name}_beam_cur_len"].tolist()
    o_raw, o_ninth = g[f"{name}_beam_topk_raw"], g[f"{name}_beam_ninth"]
    o_idx, o_rseq, o_rsc = g[f"{name}_beam_topk_indices"], g[f"{name}_beam_running_seq"], g[f"{name}_beam_running_scores"]
    alive = np.ones(B, bool)
    prev_min_gap = np.full(B, np.inf)
    stats = {"steps": 0, "cands": 0, "ids": 0}

    def cb(cur_len, ws, st):
        torch.cuda.synchronize()
        if cur_len not in cur_lens:           # the oracle's loop had ended for every row (early stopping)
            return
        s = cur_lens.index(cur_len)
        rs = st["running_scores"].cpu().numpy()
        rseq = st["running_seq"].cpu().numpy()
        row_lp, row_tok = ws["row_lp"].cpu().numpy(), ws["row_tok"].cpu().numpy()
        c_val, c_flat = _cuda_top2k(row_lp, row_tok, rs)
        with np.errstate(invalid="ignore"):
            allv = np.concatenate([o_raw[s], o_ninth[s][:, None]], 1)
            gaps = np.abs(np.diff(allv, axis=1))
            gaps = np.where(np.isfinite(gaps), gaps, np.inf)
            big = (np.abs(allv[:, :-1]) > 1e6) & (np.abs(allv[:, 1:]) > 1e6)       # -1e7-level pairs: integer-rounded scores,
            gaps = np.where(big, np.inf, gaps)                                     # exact ties resolve by index on both sides
        for b in range(B):
            if not alive[b]:
                continue
            if not np.array_equal(rseq[b], o_rseq[s, b]):
                # the kept beams differ from the oracle's: legitimate only right after a sub-tolerance decision
                assert prev_min_gap[b] < TOL_BEAM_GAP, (name, "row", b, "diverged at cur_len", cur_len,
                                                         "although the previous decision margin was", prev_min_gap[b])
                alive[b] = False
                continue
            # same beams: running scores agree, this step's candidate log-probs agree, clear candidates have equal ids
            fin = np.abs(o_rsc[s, b]) < 1e6
            assert np.all(np.abs(rs[b][fin] - o_rsc[s, b][fin]) <= TOL_BEAM_GAP), (name, b, cur_len, rs[b], o_rsc[s, b])
            assert np.all(np.abs(rs[b][~fin] - o_rsc[s, b][~fin]) <= 64.0)          # -1e7-level entries: ulp is 1.0
            for j in range(2 * K):
                if not np.isfinite(o_raw[s, b, j]) or abs(o_raw[s, b, j]) > 1e6:
                    continue
                beam, tok = int(o_idx[s, b, j] // V), int(o_idx[s, b, j] % V)
                inc_ref = o_raw[s, b, j] - o_rsc[s, b, beam]
                hit = np.where(row_tok[b * K + beam] == tok)[0]
                margin_to_rest = o_raw[s, b, j] - o_ninth[s, b]
                if len(hit) == 0:
                    assert margin_to_rest < TOL_BEAM_GAP, (name, b, cur_len, "oracle candidate", beam, tok, "missing")
                    continue
                assert abs(row_lp[b * K + beam, hit[0]] - inc_ref) <= TOL_STEP_LOGPROB, \
                    (name, b, cur_len, beam, tok, row_lp[b * K + beam, hit[0]], inc_ref)
                stats["cands"] += 1
                lo = gaps[b, j - 1] if j > 0 else np.inf
                if min(lo, gaps[b, j]) > TOL_BEAM_GAP:                              # rank j is unambiguous
                    assert c_flat[b, j] == o_idx[s, b, j], (name, b, cur_len, j, c_flat[b], o_idx[s, b])
                    stats["ids"] += 1
            prev_min_gap[b] = gaps[b].min()
        stats["steps"] += 1

    full = dict(pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, num_beams=K, min_length=0, forced_eos_token_id=2,
                length_penalty=1.0, early_stopping=True, **GEN_KW)
    out = gen.generate(model.engine, torch.from_numpy(_px(cfg)).cuda(), trace_cb=cb, **full)
    assert "decoder" in model.engine._fused_plans
    seq, sc = out["sequences"].cpu().numpy(), out["scores"].cpu().numpy()
    n_final = 0
    for b in range(B):
        if alive[b] and prev_min_gap[b] >= TOL_BEAM_GAP:
            np.testing.assert_array_equal(seq[b], g[f"{name}_beam_seq"][b])
            assert abs(sc[b] - g[f"{name}_beam_scores"][b]) <= 1.0 + 1e-6 * abs(sc[b]), (b, sc[b], g[f"{name}_beam_scores"][b])
            n_final += 1
    # the captured-graph product path: rows whose EVERY decision in the oracle trace is clear must equal the golden
    with np.errstate(invalid="ignore"):
        allv = np.concatenate([o_raw, o_ninth[:, :, None]], 2)
        gp = np.abs(np.diff(allv, axis=2))
        gp = np.where(np.isfinite(gp), gp, np.inf)
        gp = np.where((np.abs(allv[:, :, :-1]) > 1e6) & (np.abs(allv[:, :, 1:]) > 1e6), np.inf, gp)
    fully_clear = gp.min(axis=(0, 2)) >= TOL_BEAM_GAP
    for _ in range(2):
        rep = model.generate(_px(cfg), num_beams=K, **GEN_KW).sequences.cpu().numpy()
        for b in np.where(fully_clear)[0]:
            np.testing.assert_array_equal(rep[b], g[f"{name}_beam_seq"][b])
    # Beam decisions among ranks 2..8 of 1,000,216 candidates are mostly near ties (the random-init beams repeat one
    # token with almost equal scores), so rows legitimately leave the comparison early; what IS pinned: every kept
    # candidate's per-step log-prob while a row follows the oracle, ids/order at unambiguous ranks, and the final
    # sequence + score of rows that never met a sub-tolerance margin (n_final, informational at these sizes).
    print(f"[{name}] rows alive to the end {int(alive.sum())}/8, final rows compared {n_final}, {stats}")
    assert stats["steps"] >= 55 and stats["cands"] >= 100, stats
    if name == "peaked":
        ref_len = (g["peaked_beam_seq"] != 1).sum(1)
        assert (ref_len < 10).sum() >= 2 and (ref_len == 64).sum() >= 2          # EOS fired naturally in the golden
        assert stats["ids"] >= 2, stats


@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_full_size_gradients_match_golden(eps):
    """Every parameter tensor's gradient at full size (bf16 tape vs the oracle's fp32 autograd), B = 8."""
    g = np.load(GRAD)
    model, cfg = _use("perturbed")
    batch = synthetic.make_batch(cfg, B, 64, seed=3)
    eng = model.engine
    eng.dropout_p = 0.0
    ws = eng.forward_backward(torch.from_numpy(batch["pixel_values"]), torch.from_numpy(batch["decoder_input_ids"]),
                              torch.from_numpy(batch["attention_mask"]), torch.from_numpy(batch["input_ids"]), eps)
    torch.cuda.synchronize()
    loss = float(ws["out"][0])
    assert abs(loss - float(g[f"loss_eps{eps}"])) < 2e-2, (loss, g[f"loss_eps{eps}"])
    grads = model.store.to_numpy_tree(model.store.grad)
    flat = dict(("/".join(k), v) for k, v in synthetic.tree_flatten(grads))
    names = [str(n) for n in g["grad_names"]]
    assert sorted(flat) == names
    ref_norms = g[f"gradnorms_eps{eps}"]
    worst = ("", 0.0)
    for n, rn in zip(names, ref_norms):
        if "post_layernorm" in n:
            assert float(np.abs(flat[n]).max()) == 0.0          # dead parameters (pooled output unused): zero gradient
            continue
        gn = float(np.linalg.norm(flat[n].astype(np.float64)))
        # 5 % of the norm, plus an absolute floor for gradients that are analytically ~0 (a key-projection bias shifts
        # every score of a softmax row equally: its true gradient is 0 and ours is bf16 noise)
        rel = max(abs(gn - rn) - 3e-5, 0.0) / (rn + 1e-12)
        if rel > worst[1]:
            worst = (n, rel)
        assert rel < 0.05, (n, gn, rn)
    if eps == 0.0:
        def relerr(a, b):
            return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-20))
        rows = g["emb_rows"]
        checks = {
            "shared.embedding rows": relerr(flat["model/shared/embedding"][rows], g["grad_emb_rows"]),
            "final_logits_bias rows": relerr(flat["final_logits_bias"][0, rows], g["grad_flb_rows"]),
            "final_logits_bias[:4096]": relerr(flat["final_logits_bias"][0, :4096], g["grad_flb_head"]),
            "visual_projection.kernel": relerr(flat["model/visual_projection/kernel"], g["grad_proj_kernel"]),
            "visual_projection.bias": relerr(flat["model/visual_projection/bias"], g["grad_proj_bias"]),
            "decoder.0.fc1.bias": relerr(flat["model/decoder/layers/0/fc1/bias"], g["grad_dec0_fc1_bias"]),
            "decoder.11.q_proj.kernel[:64]": relerr(flat["model/decoder/layers/11/self_attn/q_proj/kernel"][:64],
                                                    g["grad_dec11_q_kernel_head"]),
            "embed_positions[:80]": relerr(flat["model/decoder/embed_positions/embedding"][:80], g["grad_pos_emb_head"]),
            "vit.0.layer_norm1.scale": relerr(flat["model/encoder/vision_model/encoder/layers/0/layer_norm1/scale"],
                                              g["grad_vit0_ln1_scale"]),
            "patch_embedding.kernel[:2,:2]": relerr(
                flat["model/encoder/vision_model/embeddings/patch_embedding/kernel"][:2, :2], g["grad_patch_kernel_slice"]),
        }
        print("full-size gradient slices, relative L2 error:", {k: round(v, 4) for k, v in checks.items()},
              "worst norm", worst)
        for k, v in checks.items():
            assert v < 0.08, (k, v)
