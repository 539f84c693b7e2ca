"""Per-phase timing of the persistent decoder-step kernel (globaltimer stamps at every grid-barrier arrival)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mic_b200
from mic_b200 import synthetic, generation as gen, ops

cfg = mic_b200.clip_mbart_config()
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
eng = model.engine
B, K, T = (int(sys.argv[2]) if len(sys.argv) > 2 else 64), 4, 64
R = B * K
px = torch.from_numpy(synthetic.make_batch(cfg, B, 64, seed=7)["pixel_values"]).cuda()
enc = eng.encode(px, trunc_int=True, save=False, tag="gen.enc")
enc_kv = eng.cross_kv(enc, tag="gen.enc")
cache = gen.DecodeCache(eng, R, T, enc_kv, K, use_ancestors=True)
tokens = torch.randint(4, 250000, (R,), dtype=torch.int32, device="cuda")
for pos in range(40):                      # fill the cache a bit
    gen.decode_step_fused(eng, cache, tokens, pos)
fp = gen.fused_prepare(eng, cache)
plan, sync = fp['plan'], fp['sync']
L = cfg.mbart_config.decoder_layers
G = torch.cuda.get_device_properties(0).multi_processor_count
P = 1 + 11 * L
prof = torch.zeros(P * G + 256, dtype=torch.int64, device="cuda")
pos = int(sys.argv[1]) if len(sys.argv) > 1 else 40
opts = int(sys.argv[3]) if len(sys.argv) > 3 else ops.DECODER_STEP_OPTS[0]
for _ in range(3):
    ops.decoder_step(plan, L, R, pos, sync, prof, opts=opts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.decoder_step(plan, L, R, pos, sync, None, opts=opts)
e1.record()
torch.cuda.synchronize()
print(f"opts={opts}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (20 back-to-back launches, CUDA events)")
raw = prof.cpu().numpy().astype(np.int64)
t = raw[:P * G].reshape(P, G)
tr = raw[P * G:]
end = t.max(axis=1)                        # phase complete = last arrival
first = t.min(axis=1)
names = ["qkv", "self_attn", "sa_o", "ln_ca", "ca_q", "cross_attn", "ca_o", "ln_f", "fc1", "fc2", "ln_next"]
dur = np.diff(end).reshape(L, 11)          # phase 0 (LN_0) is the time origin
spread = (end - first)[1:].reshape(L, 11)
print(f"pos={pos}  total step (first arrival of phase 0 -> last of phase {P - 1}): {(end[-1] - first[0]) / 1e3:.1f} us")
print("phase        mean us   (arrival spread us)")
for i, n in enumerate(names):
    print(f"{n:11s} {dur[:, i].mean() / 1e3:8.2f}   {spread[:, i].mean() / 1e3:8.2f}")
print(f"per layer: {dur.sum(axis=1).mean() / 1e3:.1f} us")

TP = 1 + 8 + 11 * 5
t0 = end[TP - 1]
f = lambda xs: " ".join(f"{(x - t0) / 1e3:.2f}" for x in xs)
print(f"fine trace of fc1/layer5 on CTA 1 (us after the previous phase completed; CTA's own arrival {(t[TP, 1] - t0) / 1e3:.2f}):")
print("  dependency seen by the activation producer:", f(tr[129:130]))
print("  W issue :", f(tr[32:48]))
print("  A issue :", f(tr[0:16]))
print("  full bar:", f(tr[64:80]))
print("  epilogue: accumulator ready", f(tr[128:129]), " stores issued", f(tr[130:131]), " proxy fence done", f(tr[131:132]))
print("  epilogue detail: tmem loaded", f(tr[132:133]), " gelu done", f(tr[133:134]))
TA = 1 + 1 + 11 * 5
t0 = end[TA - 1]
print(f"fine trace of self_attn/layer5 on CTA 1 warp 4: phase entered {f(tr[169:170])}, items done {f(tr[168:169])}, arrival {(t[TA, 1] - t0) / 1e3:.2f}")
