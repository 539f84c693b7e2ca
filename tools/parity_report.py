"""Measured deviation of both compute modes from the oracle's golden vectors of BASELINE configs[0] (batch 8, 64 tokens)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import mic_b200
from mic_b200 import synthetic
gold = np.load("tests/golden/config1_full_golden.npz")
cfg = mic_b200.clip_mbart_config()
params = synthetic.make_params(cfg, seed=1, perturbed=False)
batch = synthetic.make_batch(cfg, 8, 64, seed=0)
for dt in ("float32", "bfloat16"):
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, dtype=dt, _do_init=False)
    model.params = params
    logits = model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"]).logits
    cols = torch.from_numpy(gold["cols"]).to(logits.device)
    sl = logits[:, :, cols].float().cpu().numpy()
    rel = np.abs(sl - gold["logits_slice"]).max() / float(gold["logits_absmax"])
    lse = torch.logsumexp(logits.float(), -1).cpu().numpy()
    del logits
    out = [dt, "logits rel err %.2e" % rel, "lse max err %.2e" % np.abs(lse - gold["lse"]).max()]
    for eps in (0.0, 0.1):
        loss = float(model.loss(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], batch["input_ids"], eps))
        out.append("loss(eps=%.1f) err %.2e" % (eps, abs(loss - float(gold[f"loss_eps{eps}"]))))
    print(" | ".join(out), flush=True)
    del model
    torch.cuda.empty_cache()
