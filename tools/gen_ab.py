"""A/B: beam-4 generation at BASELINE configs[3] size, per-op decode path vs the persistent decoder-step kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mic_b200
from mic_b200 import synthetic, generation as gen

cfg = mic_b200.clip_mbart_config()
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
px = torch.from_numpy(synthetic.make_batch(cfg, B, 64, seed=7)["pixel_values"]).cuda()
kw = dict(max_length=64, pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, num_beams=4, min_length=0,
          forced_bos_token_id=250005, forced_eos_token_id=2, length_penalty=1.0, early_stopping=True)
res = {}
for fused in (False, True):
    model.engine.fused_decoder = fused
    model.engine.__dict__.pop("_gen_graphs", None)
    out = gen.generate(model.engine, px, **kw)
    out = gen.generate(model.engine, px, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = gen.generate(model.engine, px, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[fused] = out
    base = res[False]
    same = (base["sequences"] == out["sequences"]).all(dim=1).float().mean().item()
    print(f"fused_decoder={fused}: {ms:.2f} ms -> {B / ms * 1e3:.1f} captions/s; rows identical to per-op path: {same:.2f}; "
          f"score diff {(base['scores'] - out['scores']).abs().max().item():.4f}", flush=True)
