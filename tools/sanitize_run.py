"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): one tiny training step through
its eager, graph-capture and graph-replay forms, a fused beam-4 generate (eager + replay: persistent decoder-step
kernel with its grid barrier, cross-proxy fences and bulk copies), a greedy generate, and the per-op decode path.

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mic_b200
from mic_b200 import synthetic

what = sys.argv[1] if len(sys.argv) > 1 else "all"
cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
model.params = synthetic.make_params(cfg, seed=1, perturbed=True, std=0.05)
batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=0, min_len=4)
if what in ("all", "train"):
    state = mic_b200.TrainState(model, mic_b200.create_learning_rate_fn(1000, 4, 1, 10, 1e-3))
    losses = []
    for i in range(4):           # eager, eager + capture, replay, replay
        _, m = mic_b200.train_step(state, batch)
        losses.append(float(m["loss"]))
    print("train losses", losses)
    assert all(np.isfinite(losses))
if what in ("all", "gen"):
    kw = dict(num_beams=4, max_length=8, forced_bos_token_id=1001)
    a = model.generate(batch["pixel_values"], **kw).sequences.cpu().numpy()      # eager warm-up + capture
    b = model.generate(batch["pixel_values"], **kw).sequences.cpu().numpy()      # graph replay
    assert (a == b).all()
    g = model.generate(batch["pixel_values"], num_beams=1, max_length=8, forced_bos_token_id=1001).sequences.cpu().numpy()
    model.engine.fused_decoder = False
    model.engine.__dict__.pop("_gen_graphs", None)
    c = model.generate(batch["pixel_values"], **kw).sequences.cpu().numpy()      # per-op decode path
    print("beam fused", a[0].tolist(), "per-op", c[0].tolist(), "greedy", g[0].tolist())
torch.cuda.synchronize()
print("sanitize_run ok")
