"""What each fused epilogue feature costs on the short-K forward GEMMs of the decoder (M=16384, N=1024, K=1024 and
friends): plain store / + bias / + residual / + dropout / all, per layout, CUDA-graph timed over rotating buffers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mic_b200  # noqa: E402
from mic_b200 import ops  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=10, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3


def main():
    torch.manual_seed(0)
    seed = torch.zeros(1, dtype=torch.int32, device=DEV)
    shapes = [(16384, 1024, 1024), (16384, 1024, 4096), (12800, 768, 768), (12800, 768, 3072), (16384, 4096, 1024), (16384, 3072, 1024)]
    bns = [int(x) for x in os.environ.get("BNS", "0").split(",")]
    print(f"{'M':>6s} {'N':>5s} {'K':>5s} {'bn':>4s} | " + " ".join(f"{n:>12s}" for n in ["plain", "+bias", "+residual", "+dropout", "bias+res+drop", "dgrad(0,0)"]))
    for M, N, K in shapes:
        nset = 4
        A = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(nset)]
        W = [(torch.randn(K, N, device=DEV) * 0.03).bfloat16() for _ in range(nset)]
        Wt = [w.t().contiguous() for w in W]           # [N, K]
        R = [torch.randn(M, N, device=DEV).bfloat16() for _ in range(nset)]
        O = [torch.empty(M, N, dtype=torch.bfloat16, device=DEV) for _ in range(nset)]
        bias = torch.randn(N, device=DEV)
        for bn in bns:
            res = []
            for kw in [dict(), dict(bias=bias), dict(residual=True), dict(dropout=(seed, 7, 0.1)), dict(bias=bias, residual=True, dropout=(seed, 7, 0.1))]:
                st = {"i": 0}

                def f():
                    i = st["i"] % nset
                    st["i"] += 1
                    k = dict(kw)
                    if k.get("residual"):
                        k["residual"] = R[i]
                    ops.gemm(A[i], W[i], b_mn=True, out=O[i], block_n=bn, **k)
                res.append(timeit(f))
            st = {"i": 0}

            def f2():
                i = st["i"] % nset
                st["i"] += 1
                ops.gemm(A[i], Wt[i], b_mn=False, out=O[i], block_n=bn)
            res.append(timeit(f2))
            fl = 2.0 * M * N * K
            print(f"{M:6d} {N:5d} {K:5d} {bn:4d} | " + " ".join(f"{t:6.1f}us {fl / t / 1e6:4.0f}" for t in res))


if __name__ == "__main__":
    main()
