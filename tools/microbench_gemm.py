"""Fixed-overhead microbenchmarks for the tcgen05 GEMM (CUDA events, no profiler)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mic_b200
from mic_b200 import ops

def timeit(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def graph_timeit(fn, n=200):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

dev = "cuda"
bf = torch.bfloat16
for (M, N, K, bn) in [(256, 1024, 64, 64), (256, 1024, 1024, 64), (256, 1024, 1024, 128), (256, 1024, 1024, 256), (256, 4096, 1024, 64),
                      (256, 1024, 4096, 64), (256, 1024, 4096, 128), (16384, 1024, 1024, 256), (16384, 1024, 64, 256), (128, 64, 64, 64)]:
    a = torch.randn(M, K, device=dev).to(bf); w = torch.randn(K, N, device=dev).to(bf) * 0.05
    bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=bf)
    f = lambda: ops.gemm(a, w, b_mn=True, bias=bias, out=out, block_n=bn)
    print(f"gemm M={M} N={N} K={K} bn={bn}: eager {timeit(f):7.1f} us   graph {graph_timeit(f):7.1f} us")
x = torch.randn(256, 1024, device=dev).to(bf); g = torch.ones(1024, device=dev); b = torch.zeros(1024, device=dev); y = torch.empty_like(x)
ln = lambda: ops.layernorm_fwd(x, g, b, 1e-5, out=y)
print(f"layernorm 256x1024: eager {timeit(ln):7.1f} us  graph {graph_timeit(ln):7.1f} us")
a = torch.randn(256, 1024, device=dev).to(bf); w = torch.randn(1024, 1024, device=dev).to(bf) * 0.05
bias = torch.zeros(1024, device=dev); out = torch.empty(256, 1024, device=dev, dtype=bf)
def both():
    ops.layernorm_fwd(x, g, b, 1e-5, out=y)
    ops.gemm(y, w, b_mn=True, bias=bias, out=out, block_n=64)
print(f"LN + gemm alternating: graph {graph_timeit(both):7.1f} us per pair")
# decode attention
R, H, T = 256, 16, 64
q = torch.randn(R, 1024, device=dev).to(bf); cache = torch.randn(R * T, 2048, device=dev).to(bf)
anc = torch.arange(R, device=dev, dtype=torch.int32)[:, None].expand(R, T).contiguous(); o = torch.empty(R, 1024, device=dev, dtype=bf)
for nk in (1, 32, 64):
    f = lambda: ops.decode_attention(q, cache[:, :1024], cache[:, 1024:], 2048, anc, T, nk, 1, o, R, H, 0.125)
    print(f"decode self-attn n_keys={nk}: graph {graph_timeit(f):7.1f} us")
enc = torch.randn(64 * 50, 24576, device=dev).to(bf)
f = lambda: ops.decode_attention(q, enc[:, :1024], enc[:, 1024:2048], 24576, None, 50, 50, 4, o, R, H, 0.125)
print(f"decode cross-attn S=50: graph {graph_timeit(f):7.1f} us")
