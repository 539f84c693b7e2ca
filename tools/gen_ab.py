"""A/B: beam-4 generation at BASELINE configs[3] size with programmatic dependent launch on / off."""
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
for graph in (False, True):
    for pdl, pre in ((False, False), (True, False), (True, True)):
        out = gen.generate(model.engine, px, use_cuda_graph=graph, pdl=pdl, prefetch_weights=pre, **kw)
        out = gen.generate(model.engine, px, use_cuda_graph=graph, pdl=pdl, prefetch_weights=pre, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = gen.generate(model.engine, px, use_cuda_graph=graph, pdl=pdl, prefetch_weights=pre, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[(graph, pdl, pre)] = out
        base = res[(False, False, False)]
        print(f"graph={graph} pdl={pdl} prefetch={pre}: {ms:.2f} ms -> {B / ms * 1e3:.1f} captions/s; identical to baseline: "
              f"{torch.equal(base['sequences'], out['sequences'])}, score diff {(base['scores'] - out['scores']).abs().max().item():.4f}", flush=True)
