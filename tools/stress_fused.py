"""Poison the caching allocator with NaN, then run tiny greedy/beam generation through the fused and the per-op
decode paths: an uninitialised read shows up deterministically."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mic_b200
from mic_b200 import synthetic, generation as gen

def poison(gb=6):
    xs = [torch.full((256 * 1024 * 1024,), 37.0, device="cuda") for _ in range(gb)]   # 1 GB each
    small = [torch.full((n,), 37.0, device="cuda") for n in (1000, 10000, 100000, 1000000) for _ in range(20)]
    del xs, small

cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
params = synthetic.make_params(cfg, seed=7, perturbed=True, std=0.12)
params["final_logits_bias"] = params["final_logits_bias"].copy(); params["final_logits_bias"][0, 2] += 6.0
batch = synthetic.make_batch(cfg, 5, seq_len=16, seed=7)
which = sys.argv[1] if len(sys.argv) > 1 else "poison"
if which == "poison":
    poison()
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg)
model.params = params
px = torch.from_numpy(batch["pixel_values"]).cuda()
kws = [dict(num_beams=4, max_length=8, forced_bos_token_id=1001), dict(num_beams=1, max_length=12, forced_bos_token_id=1001),
       dict(num_beams=2, max_length=6), dict(num_beams=1, max_length=12)]
base = dict(pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, min_length=0, forced_eos_token_id=2,
            length_penalty=1.0, early_stopping=True, forced_bos_token_id=None)
bad = 0
for rep in range(6):
    for kw in kws:
        full = dict(base); full.update(kw)
        for n, t_ in list(model.engine.bufs.t.items()):
            if n.startswith("gen.") and n not in ("gen.acc", "gen.q_acc") and t_.dtype in (torch.bfloat16, torch.float32, torch.uint8):
                t_.fill_(37 if t_.dtype != torch.uint8 else 66)          # stale finite garbage in every decode buffer
        model.engine.fused_decoder = True
        b = gen.generate(model.engine, px, use_cuda_graph=False, **full)["sequences"].cpu().numpy()
        model.engine.fused_decoder = False
        a = gen.generate(model.engine, px, use_cuda_graph=False, **full)["sequences"].cpu().numpy()
        same = (a == b).all(axis=1)
        if not same.all():
            bad += 1
            print("rep", rep, kw, "rows differing:", np.where(~same)[0].tolist(), "\n  per-op:", a[~same][0].tolist(), "\n  fused :", b[~same][0].tolist(), flush=True)
print("mismatching calls:", bad, "of", 6 * len(kws))
