"""Attention kernel A/B: one-CTA-per-head kernels (impl 1, <= 64 tokens) / previous general kernels vs the row-tiled
kernels (impl 2: two-kernel backward; impl 0: default, single-kernel backward) at the shapes of the two models, timed with CUDA events over inputs larger than L2.

    python tools/attn_bench.py            # prints a table; copy to profiles/
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mic_b200  # noqa: E402
from mic_b200 import ops  # noqa: E402

DEV = "cuda:0"
SHAPES = [  # name, B, H, Tq, Tk, causal, fused qkv
    ("ViT-B/16 self (197)", 256, 12, 197, 197, False),
    ("BART cross (64x197)", 256, 12, 64, 197, False),
    ("BART self (64, causal)", 256, 12, 64, 64, True),
    ("CLIP-B/32 self (50)", 256, 12, 50, 50, False),
    ("mBART self (64, causal)", 256, 16, 64, 64, True),
    ("mBART cross (64x50)", 256, 16, 64, 50, False),
]


def timeit(fn, iters=12, reps=5):
    """`iters` launches captured in one CUDA graph (the ctypes call costs more host time than the small shapes run)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if os.environ.get("ATTN_EAGER") == "1":      # under ncu: the three launches above are what gets profiled
        return float("nan")
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
    print(f"{'shape':28s} {'impl':>6s} {'fwd us':>9s} {'bwd us':>9s} {'fwd TF/s':>9s} {'bwd TF/s':>9s} {'fwd GB/s':>9s} {'bwd GB/s':>9s}")
    for name, B, H, Tq, Tk, causal in SHAPES:
        d = H * 64
        # rotate over several buffer sets so that consecutive launches do not hit L2 (set = ~0.3-0.5 GB)
        nset = 3
        sets = []
        for _ in range(nset):
            q = torch.randn(B * Tq, d, device=DEV).bfloat16()
            kv = torch.randn(B * Tk, 2 * d, device=DEV).bfloat16()
            o = torch.empty(B * Tq, d, dtype=torch.bfloat16, device=DEV)
            do = torch.randn(B * Tq, d, device=DEV).bfloat16()
            lse = torch.empty(B, H, Tq, device=DEV)
            dq = torch.empty_like(q)
            dkv = torch.empty_like(kv)
            sets.append((q, kv, o, do, lse, dq, dkv))
        scale = 1 / math.sqrt(64)
        flops_f = 4.0 * B * H * Tq * Tk * 64 * (0.5 if causal else 1.0)
        bytes_f = 2.0 * (2 * B * Tq * d + 2 * B * Tk * d)
        bytes_b = 2.0 * (4 * B * Tq * d + 4 * B * Tk * d)
        for impl in (1, 2, 0):
            if impl == 1 and max(Tq, Tk) > 64:
                continue                      # the one-CTA-per-head kernels hold at most 64 tokens
            impl_name = {1: "1cta", 2: "tiled2", 0: "tiled"}[impl]      # tiled2: backward as dQ + dK/dV kernels
            ops.attention_impl(impl)
            st = {"i": 0}

            def fwd():
                q, kv, o, do, lse, dq, dkv = sets[st["i"] % nset]
                st["i"] += 1
                ops.attention_fwd(q, kv[:, :d], kv[:, d:], o, lse, None, causal, B, H, Tq, Tk, scale)

            def bwd():
                q, kv, o, do, lse, dq, dkv = sets[st["i"] % nset]
                st["i"] += 1
                ops.attention_bwd(q, kv[:, :d], kv[:, d:], o, do, lse, None, causal, dq, dkv[:, :d], dkv[:, d:], B, H, Tq, Tk,
                                  scale)

            for _ in range(nset):
                fwd()
            tf = timeit(fwd)
            tb = timeit(bwd)
            print(f"{name:28s} {impl_name:>6s} {tf:9.1f} {tb:9.1f} {flops_f / tf / 1e6:9.1f} {2.5 * flops_f / tb / 1e6:9.1f} "
                  f"{bytes_f / tf / 1e3:9.0f} {bytes_b / tb / 1e3:9.0f}")
        ops.attention_impl(0)


def one_shape():
    """ATTN_ONE=1: a few launches of one self-attention shape only (what the ncu capture profiles); ATTN_T=64 ATTN_H=16
    ATTN_CAUSAL=1 selects the mBART decoder shape instead of ViT-B/16's 197 tokens."""
    B, H, T = 256, int(os.environ.get("ATTN_H", "12")), int(os.environ.get("ATTN_T", "197"))
    causal = os.environ.get("ATTN_CAUSAL") == "1"
    d = H * 64
    qkv = torch.randn(B * T, 3 * d, device=DEV).bfloat16()
    o = torch.empty(B * T, d, dtype=torch.bfloat16, device=DEV)
    do = torch.randn(B * T, d, device=DEV).bfloat16()
    lse = torch.empty(B, H, T, device=DEV)
    dqkv = torch.empty_like(qkv)
    for _ in range(3):
        ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, lse, None, causal, B, H, T, T, 0.125)
        ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, do, lse, None, causal, dqkv[:, :d], dqkv[:, d:2 * d],
                          dqkv[:, 2 * d:], B, H, T, T, 0.125)
    torch.cuda.synchronize()


if __name__ == "__main__":
    one_shape() if os.environ.get("ATTN_ONE") == "1" else main()
