"""Per-shape table of every tcgen05 GEMM of one training step (B=256): launches, time, achieved TFLOP/s, and the time
above the measured sustained peak ("excess") - where the GEMM part of the step loses against the roofline.

    python tools/gemm_table.py [--model vit-bart] > profiles/r02_gemm_table.txt
"""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mic_b200  # noqa: E402
from mic_b200 import ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="clip-mbart")
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 1401.9
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    vb = args.model == "vit-bart"
    cfg = mic_b200.vit_bart_config() if vb else mic_b200.clip_mbart_config()
    cls = mic_b200.FlaxViTBartForConditionalGeneration if vb else mic_b200.FlaxCLIPVisionMBartForConditionalGeneration
    model = cls(cfg, seed=0, device=dev)
    state = mic_b200.TrainState(model, mic_b200.create_learning_rate_fn(10_000_000, args.batch, 7, 1000, 5e-5))
    hb = synthetic.make_batch(cfg, args.batch, 64, seed=2)
    devb = {k: torch.from_numpy(v).to(dev) for k, v in hb.items()}
    for _ in range(3):
        mic_b200.train_step(state, devb, use_cuda_graph=False)
    torch.cuda.synchronize()
    model.engine.overlap_wgrad = False
    rows = collections.defaultdict(lambda: [0, 0.0, 0.0])
    reps = 3
    for _ in range(reps):
        ops.TIMED["mic_gemm_bf16"] = []
        ops.TIMED_SHAPES.clear()
        mic_b200.train_step(state, devb, use_cuda_graph=False)
        torch.cuda.synchronize()
        evs = ops.TIMED.pop("mic_gemm_bf16")
        fl = ops.TIMED_FLOPS.pop("mic_gemm_bf16")
        for (s, e), f, shp in zip(evs, fl, ops.TIMED_SHAPES):
            r = rows[shp]
            r[0] += 1
            r[1] += s.elapsed_time(e)
            r[2] += f
    print(f"# every mic_gemm_bf16 launch of one eager training step ({args.model}, B={args.batch}), CUDA events per launch, "
          f"mean of {reps} steps; peak = {peak} TFLOP/s (measured sustained)")
    print(f"# {'M':>7s} {'N':>7s} {'K':>7s} aT bT  bn sk act acc f32 | {'n':>4s} {'ms':>8s} {'TF/s':>7s} {'frac':>5s} {'excess ms':>9s}")
    tot_ms = tot_fl = 0.0
    for shp, (n, ms, f) in sorted(rows.items(), key=lambda kv: -(kv[1][1] - kv[1][2] / peak / 1e9)):
        n, ms, f = n / reps, ms / reps, f / reps
        tf = f / ms / 1e9
        M, N, K, am, bm, bn, sk, act, acc, f32 = shp
        print(f"  {M:7d} {N:7d} {K:7d} {am:2d} {bm:2d} {bn:3d} {sk:2d} {act:3d} {acc:3d} {f32:3d} | {n:4.0f} {ms:8.3f} {tf:7.1f} {tf / peak:5.2f} "
              f"{ms - f / peak / 1e9:9.3f}")
        tot_ms += ms
        tot_fl += f
    print(f"# total {tot_ms:.2f} ms, {tot_fl / 1e12:.2f} TFLOP, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s = {tot_fl / tot_ms / 1e9 / peak:.3f} of peak; "
          f"excess over peak {tot_ms - tot_fl / peak / 1e9:.2f} ms")


if __name__ == "__main__":
    main()
