"""Eager (no CUDA graph) beam-4 generation at BASELINE configs[3] size, for `ncu` launch lists."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mic_b200
from mic_b200 import synthetic, generation as gen

cfg = mic_b200.clip_mbart_config()
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64
px = torch.from_numpy(synthetic.make_batch(cfg, B, 64, seed=7)["pixel_values"]).cuda()
kw = dict(max_length=L, pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, num_beams=4, min_length=0,
          forced_bos_token_id=250005, forced_eos_token_id=2, length_penalty=1.0, early_stopping=True)
out = gen.generate(model.engine, px, use_cuda_graph=False, **kw)
torch.cuda.synchronize()
t0 = time.time()
out = gen.generate(model.engine, px, use_cuda_graph=False, **kw)
torch.cuda.synchronize()
print("eager generate ms:", (time.time() - t0) * 1e3, out["sequences"][0, :8].tolist())
