"""Generates tests/golden/*.npz from the CPU oracle (oracle/).  The reference itself cannot run here
(JAX/Flax/optax are not installable offline), so these vectors pin the ORACLE — which is in
turn pinned against the HF PyTorch twins (tests/test_oracle_vs_hf.py) — and give the GPU tests fixed
targets that need no oracle run on the GPU box.

    python tests/golden/gen_golden.py [--full]      (--full adds BASELINE config 1: full size, B=8, ~3 min)
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import mic_b200  # noqa: E402
from mic_b200 import synthetic  # noqa: E402
from oracle import reference_model as rm  # noqa: E402
from oracle import reference_generate as rg  # noqa: E402


def tiny():
    cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
    params = synthetic.make_params(cfg, seed=1, perturbed=True, std=0.05)
    batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=0, min_len=4)
    out = {}
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        logits = rm.forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], None, cfg)
        enc = rm.encode(p, batch["pixel_values"], cfg)
    out["logits"] = logits.numpy()
    out["enc"] = enc.numpy()
    for eps in (0.0, 0.1):
        loss, grads, _ = rm.loss_and_grads(params, batch, cfg, eps)
        out[f"loss_eps{eps}"] = np.float32(loss)
        flat = dict(("/".join(k), v) for k, v in synthetic.tree_flatten(grads))
        out[f"gradnorms_eps{eps}"] = np.array([np.linalg.norm(flat[k]) for k in sorted(flat)], np.float32)
        if eps == 0.1:
            out["grad_names"] = np.array(sorted(flat))
            out["grad_proj_kernel"] = flat["model/visual_projection/kernel"]
            out["grad_flb"] = flat["final_logits_bias"]
    gp = synthetic.make_params(cfg, seed=5, perturbed=True, std=0.3)
    gb = synthetic.make_batch(cfg, 3, seq_len=16, seed=0, min_len=4)
    for beams in (1, 4):
        r = rg.generate(gp, gb["pixel_values"], cfg, num_beams=beams, max_length=12, forced_bos_token_id=1001,
                        return_trace=True)
        out[f"seq_beams{beams}"] = r["sequences"]
        if beams == 4:
            out["scores_beams4"] = r["scores"]
            L = 12
            clear = np.ones(3, bool)
            for step in r["trace"]:
                if step["cur_len"] in (1, L - 1):
                    continue
                v = step["topk_log_probs"]
                gaps = np.abs(np.diff(np.concatenate([v, step["ninth"][:, None]], 1), axis=1))
                clear &= np.all((gaps > 0.05) | ~np.isfinite(gaps), axis=1)
            out["clear_beams4"] = clear
        else:
            out["margins_beams1"] = r["margins"]
    np.savez_compressed(os.path.join(HERE, "tiny_golden.npz"), **out)
    print("tiny golden written:", {k: getattr(v, "shape", ()) for k, v in out.items()})


def full():
    """BASELINE configs[0]: CLIP-ViT-B/32 + mBART-50 forward+loss, random init, batch 8, 64-token captions."""
    cfg = mic_b200.clip_mbart_config()
    params = synthetic.make_params(cfg, seed=1, perturbed=False)
    batch = synthetic.make_batch(cfg, 8, 64, seed=0)
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        logits = rm.forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], None, cfg)
        out = {"shape": np.array(logits.shape)}
        for eps in (0.0, 0.1):
            out[f"loss_eps{eps}"] = np.float32(rm.loss_fn(logits, batch["input_ids"], batch["attention_mask"], eps))
        cols = np.r_[0:64, 125000:125064, 250003:250054]
        out["cols"] = cols
        out["logits_slice"] = logits[:, :, cols].numpy()
        out["logits_absmax"] = np.float32(logits.abs().max())
        out["lse"] = torch.logsumexp(logits, -1).numpy()
        out["argmax"] = logits.argmax(-1).numpy().astype(np.int32)
        srt = torch.topk(logits, 2, dim=-1).values
        out["top2_gap"] = (srt[..., 0] - srt[..., 1]).numpy()
    np.savez_compressed(os.path.join(HERE, "config1_full_golden.npz"), **out)
    print("full golden written; loss", out["loss_eps0.0"], "shape", out["shape"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    tiny()
    if a.full:
        full()
