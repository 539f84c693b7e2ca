"""Full-size goldens for the configs the headline metric is quoted on.

    python tests/golden/gen_golden_full.py            (~15 min of CPU, ~12 GB RSS)

Writes
  config3_full_gen_golden.npz  BASELINE configs[3] at full size (CLIP-ViT-B/32 + mBART-50, V=250,054, 12 layers),
      B=8: greedy and beam-4, max_length=64, forced_bos es_XX=250005, two weight sets
        "init"   = synthetic.make_params(seed=1)            (BASELINE's random init; no natural EOS)
        "peaked" = synthetic.make_peaked_params(seed=11)    (peaked next-token distribution, EOS fires naturally)
      with the per-step decision-margin trace (top-2K raw candidate scores + the best rejected one, did_finish)
      so the GPU test compares token ids exactly up to the first position whose margin is below its tolerance.
  config1_full_grad_golden.npz BASELINE configs[0/1] shape at full size, B=8: loss and the gradient of EVERY
      parameter tensor (norms) plus slices of shared.embedding / final_logits_bias / visual_projection.
The weights are NOT stored: both sides rebuild them from the numpy seeds.
The oracle is pinned against the HF PyTorch twins only (the reference has no goldens): PARITY UNPINNED.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import mic_b200  # noqa: E402
from mic_b200 import synthetic  # noqa: E402
from oracle import reference_model as rm  # noqa: E402
from oracle import reference_generate as rg  # noqa: E402

GEN_KW = dict(max_length=64, forced_bos_token_id=250005)
B = 8


def gen():
    cfg = mic_b200.clip_mbart_config()
    px = synthetic.make_batch(cfg, B, 64, seed=7)["pixel_values"]
    out = {}
    for name, params in (("init", synthetic.make_params(cfg, seed=1)),
                         ("peaked", synthetic.make_peaked_params(cfg, seed=11))):
        t0 = time.time()
        r = rg.generate(params, px, cfg, num_beams=1, return_trace=True, **GEN_KW)
        out[f"{name}_greedy_seq"] = r["sequences"]
        out[f"{name}_greedy_margins"] = r["margins"].astype(np.float32)
        print(name, "greedy", time.time() - t0, "s", r["sequences"][:2, :12], flush=True)
        t0 = time.time()
        r = rg.generate(params, px, cfg, num_beams=4, return_trace=True, **GEN_KW)
        out[f"{name}_beam_seq"] = r["sequences"]
        out[f"{name}_beam_scores"] = r["scores"]
        out[f"{name}_beam_all_seq"] = r["all_sequences"].astype(np.int32)
        tr = r["trace"]
        out[f"{name}_beam_cur_len"] = np.array([s["cur_len"] for s in tr], np.int32)
        out[f"{name}_beam_topk_raw"] = np.stack([s["topk_raw"] for s in tr]).astype(np.float32)        # [steps, B, 2K]
        out[f"{name}_beam_ninth"] = np.stack([s["ninth"] for s in tr]).astype(np.float32)             # [steps, B]
        out[f"{name}_beam_did_finish"] = np.stack([s["did_finish"] for s in tr])
        out[f"{name}_beam_topk_indices"] = np.stack([s["topk_indices"] for s in tr]).astype(np.int64)
        out[f"{name}_beam_running_seq"] = np.stack([s["running_sequences"] for s in tr]).astype(np.int32)   # at step entry
        out[f"{name}_beam_running_scores"] = np.stack([s["running_scores"] for s in tr]).astype(np.float32)
        print(name, "beam", time.time() - t0, "s; steps", len(tr), r["sequences"][:2, :12], r["scores"], flush=True)
        del params
    np.savez_compressed(os.path.join(HERE, "config3_full_gen_golden.npz"), **out)


def grad():
    cfg = mic_b200.clip_mbart_config()
    params = synthetic.make_params(cfg, seed=1, perturbed=True)
    batch = synthetic.make_batch(cfg, B, 64, seed=3)
    out = {}
    for eps in (0.0, 0.1):
        t0 = time.time()
        loss, grads, _ = rm.loss_and_grads(params, batch, cfg, eps)
        flat = dict(("/".join(k), v) for k, v in synthetic.tree_flatten(grads))
        names = sorted(flat)
        out[f"loss_eps{eps}"] = np.float32(loss)
        out[f"gradnorms_eps{eps}"] = np.array([np.linalg.norm(flat[k].astype(np.float64)) for k in names], np.float64)
        if eps == 0.0:
            out["grad_names"] = np.array(names)
            lab = np.unique(batch["input_ids"])
            rows = np.unique(np.r_[lab[:24], 0, 1, 2, 3, 1000, 125000, 250003:250008, 250053])
            out["emb_rows"] = rows.astype(np.int64)
            out["grad_emb_rows"] = flat["model/shared/embedding"][rows]
            out["grad_flb_rows"] = flat["final_logits_bias"][0, rows]
            out["grad_flb_head"] = flat["final_logits_bias"][0, :4096]
            out["grad_proj_kernel"] = flat["model/visual_projection/kernel"]
            out["grad_proj_bias"] = flat["model/visual_projection/bias"]
            out["grad_dec0_fc1_bias"] = flat["model/decoder/layers/0/fc1/bias"]
            out["grad_dec11_q_kernel_head"] = flat["model/decoder/layers/11/self_attn/q_proj/kernel"][:64]
            out["grad_pos_emb_head"] = flat["model/decoder/embed_positions/embedding"][:80]
            out["grad_vit0_ln1_scale"] = flat["model/encoder/vision_model/encoder/layers/0/layer_norm1/scale"]
            out["grad_patch_kernel_slice"] = flat["model/encoder/vision_model/embeddings/patch_embedding/kernel"][:2, :2]
        print("grad eps", eps, "loss", loss, time.time() - t0, "s", flush=True)
    np.savez_compressed(os.path.join(HERE, "config1_full_grad_golden.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["grad", "gen"]
    torch.set_num_threads(os.cpu_count())
    if "grad" in which:
        grad()
    if "gen" in which:
        gen()
