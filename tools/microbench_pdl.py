"""Per-kernel cost of dependent chains inside a CUDA graph, with / without programmatic dependent launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mic_b200
from mic_b200 import ops

def graph_timeit(fn, n=100):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

dev, bf = "cuda", torch.bfloat16
R, d, F = 256, 1024, 4096
x = torch.randn(R, d, device=dev).to(bf); y = torch.empty_like(x); o = torch.empty_like(x)
g1 = torch.ones(d, device=dev); b1 = torch.zeros(d, device=dev)
# 24 distinct weight matrices so the stream really comes from HBM/L2 like the decode loop
ws = [torch.randn(d, d, device=dev).to(bf) * 0.03 for _ in range(48)]
w1 = [torch.randn(d, F, device=dev).to(bf) * 0.03 for _ in range(12)]
w2 = [torch.randn(F, d, device=dev).to(bf) * 0.03 for _ in range(12)]
bias = torch.zeros(d, device=dev); biasF = torch.zeros(F, device=dev)
h = torch.empty(R, F, device=dev, dtype=bf)
acc = torch.zeros(R, d, device=dev)
T, H = 64, 16
cache = torch.randn(R * T, 2 * d, device=dev).to(bf)
anc = torch.arange(R, device=dev, dtype=torch.int32)[:, None].expand(R, T).contiguous()
i = [0]
def chain_gemm():
    i[0] = (i[0] + 1) % 48
    ops.gemm(x, ws[i[0]], b_mn=True, bias=bias, out=y, block_n=64)
    i[0] = (i[0] + 1) % 48
    ops.gemm(y, ws[i[0]], b_mn=True, bias=bias, out=x, block_n=64)
def chain_ln_gemm():
    i[0] = (i[0] + 1) % 48
    ops.layernorm_fwd(x, g1, b1, 1e-5, out=y)
    ops.gemm(y, ws[i[0]], b_mn=True, bias=bias, out=x, block_n=64)
def chain_attn_gemm():
    i[0] = (i[0] + 1) % 48
    ops.decode_attention(x, cache[:, :d], cache[:, d:], 2 * d, anc, T, 32, 1, o, R, H, 0.125)
    ops.gemm(o, ws[i[0]], b_mn=True, bias=bias, out=x, block_n=64)
def chain_ffn():
    i[0] = (i[0] + 1) % 12
    ops.gemm(x, w1[i[0]], b_mn=True, bias=biasF, act="gelu", out=h)
    ops.gemm(h, w2[i[0]], b_mn=True, out=acc, accumulate=True, split_k=8, block_n=64)
    ops.residual_ln_fwd(acc, bias, y, g1, b1, 1e-5, x)
def chain_ln():
    ops.layernorm_fwd(x, g1, b1, 1e-5, out=y)
    ops.layernorm_fwd(y, g1, b1, 1e-5, out=x)
for name, fn, k in (("gemm->gemm", chain_gemm, 2), ("ln->gemm", chain_ln_gemm, 2), ("attn->gemm", chain_attn_gemm, 2),
                    ("fc1->fc2->resln", chain_ffn, 3), ("ln->ln", chain_ln, 2)):
    row = []
    for pdl, pre in ((0, 0), (1, 0), (1, 1)):
        ops.launch_options(pdl=pdl, gemm_b_static=pre)
        row.append(graph_timeit(fn) / k)
    print(f"{name:18s} us/kernel: no-pdl {row[0]:6.2f}   pdl {row[1]:6.2f}   pdl+prefetch {row[2]:6.2f}", flush=True)
