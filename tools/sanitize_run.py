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
if what in ("all", "attn"):
    # row-tiled attention (forward, fused backward <= 128 tokens, dQ + dK/dV kernels beyond), masked / causal / ragged
    # shapes, and the image transform: the kernels added after the first sanitizer pass of this round
    from mic_b200 import ops, transforms
    dev = "cuda:0"
    for (Tq, Tk, causal, masked) in [(50, 50, False, False), (64, 64, True, True), (64, 50, False, False),
                                     (197, 197, False, False), (64, 197, False, True), (130, 100, True, True), (5, 5, False, False)]:
        B, H, d = 2, 2, 128
        q = torch.randn(B * Tq, d, device=dev).bfloat16()
        kv = torch.randn(B * Tk, 2 * d, device=dev).bfloat16()
        km = None
        if masked:
            km = torch.ones(B, Tk, dtype=torch.int32, device=dev)
            km[0, Tk // 2:] = 0
        o = torch.empty_like(q)
        lse = torch.empty(B, H, Tq, device=dev)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        for impl in (0, 2):
            ops.attention_impl(impl)
            ops.attention_fwd(q, kv[:, :d], kv[:, d:], o, lse, km, causal, B, H, Tq, Tk, 0.125)
            ops.attention_bwd(q, kv[:, :d], kv[:, d:], o, torch.randn_like(q), lse, km, causal, dq, dkv[:, :d], dkv[:, d:],
                              B, H, Tq, Tk, 0.125)
        ops.attention_impl(0)
        assert torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
    rng = np.random.RandomState(0)
    imgs = [rng.randint(0, 256, (3, h, w)).astype(np.uint8) for h, w in [(41, 67), (90, 33), (32, 32), (5, 200)]]
    for cf in (False, True):
        u8 = transforms.BatchTransform(32, dev, channel_first=cf)(imgs)
    print("attention + transform ok", tuple(u8.shape))
torch.cuda.synchronize()
print("sanitize_run ok")
