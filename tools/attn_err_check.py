"""Error of both attention kernel families against an fp32 torch reference (rms and max), and the tiny ViT-BART logits check."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mic_b200
from mic_b200 import ops, synthetic
from oracle import reference_model as rm

DEV = "cuda:0"
def ref(q, k, v, causal, scale):
    w = torch.einsum("bqhd,bkhd->bhqk", q * scale, k)
    if causal:
        Tq, Tk = w.shape[-2:]
        w = w.masked_fill(~torch.tril(torch.ones(Tq, Tk, dtype=torch.bool, device=q.device)), float("-inf"))
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(w, -1), v)

for (Tq, Tk, causal) in [(5, 5, False), (17, 17, False), (50, 50, False), (64, 64, True), (64, 50, False)]:
    B, H, hd = 4, 4, 64
    d = H * hd
    g = torch.Generator().manual_seed(1)
    q = (torch.randn(B * Tq, d, generator=g) * 1.5).bfloat16().to(DEV)
    kv = (torch.randn(B * Tk, 2 * d, generator=g) * 1.5).bfloat16().to(DEV)
    do = torch.randn(B * Tq, d, generator=g).bfloat16().to(DEV)
    for impl in (1, 0):
        ops.attention_impl(impl)
        out = torch.empty(B * Tq, d, dtype=torch.bfloat16, device=DEV)
        lse = torch.empty(B, H, Tq, device=DEV)
        ops.attention_fwd(q, kv[:, :d], kv[:, d:], out, lse, None, causal, B, H, Tq, Tk, 0.125)
        qf = q.float().view(B, Tq, H, hd).requires_grad_(True)
        kf = kv[:, :d].float().reshape(B, Tk, H, hd).requires_grad_(True)
        vf = kv[:, d:].float().reshape(B, Tk, H, hd).requires_grad_(True)
        want = ref(qf, kf, vf, causal, 0.125)
        dq = torch.empty_like(q); dkv = torch.empty_like(kv)
        ops.attention_bwd(q, kv[:, :d], kv[:, d:], out, do, lse, None, causal, dq, dkv[:, :d], dkv[:, d:], B, H, Tq, Tk, 0.125)
        want.backward(do.float().view(B, Tq, H, hd))
        torch.cuda.synchronize()
        def e(a, b):
            a = a.float().reshape(b.shape); return f"rms {float((a-b).pow(2).mean().sqrt()/b.pow(2).mean().sqrt()):.2e} max {float((a-b).abs().max()/b.abs().max()):.2e}"
        lref = torch.logsumexp(torch.einsum("bqhd,bkhd->bhqk", qf * 0.125, kf).masked_fill(
            ~torch.tril(torch.ones(Tq, Tk, dtype=torch.bool, device=DEV)) if causal else torch.zeros(Tq, Tk, dtype=torch.bool, device=DEV), float("-inf")), -1)
        print(f"T=({Tq},{Tk}) causal={causal} impl={impl}: out {e(out, want.detach())} | dq {e(dq, qf.grad)} | dk {e(dkv[:, :d], kf.grad)} | dv {e(dkv[:, d:], vf.grad)} | lse max {float((lse-lref).abs().max()):.2e}")
ops.attention_impl(0)

cfg = mic_b200.tiny_vit_bart_config(vocab_size=1003, layers=2)
params = synthetic.make_params(cfg, seed=4, perturbed=True, std=0.12)
batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=4, min_len=4)
with torch.no_grad():
    want = rm.forward_logits(rm.to_torch_tree(params), batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], None, cfg).numpy()
for impl in (1, 0):
    ops.attention_impl(impl)
    model = mic_b200.FlaxViTBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    got = model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"]).logits.float().cpu().numpy()
    print(f"tiny ViT-BART logits impl={impl}: max err {np.abs(got-want).max():.4f} (max |want| {np.abs(want).max():.3f}) rms {np.sqrt(((got-want)**2).mean()):.4f}")
ops.attention_impl(0)
