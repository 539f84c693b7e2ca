"""GPU diagnostic for the tcgen05 GEMM: runs each (layout, block_n) in a subprocess with a timeout so a
hung pipeline cannot take the box down, and prints where errors sit (row/col/k structure)."""
import subprocess
import sys

CHILD = r'''
import sys, math, torch
sys.path.insert(0, ".")
import mic_b200
from mic_b200 import ops
layout, bn, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
torch.manual_seed(0)
a = torch.randn(M, K).to(torch.bfloat16).cuda()
b = torch.randn(N, K).to(torch.bfloat16).cuda()
a_mn = layout == "mm"; b_mn = layout in ("kn", "mm")
A = a.t().contiguous() if a_mn else a
B = b.t().contiguous() if b_mn else b
out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, block_n=bn)
torch.cuda.synchronize()
want = a.float() @ b.float().t()
err = (out.float() - want).abs()
tol = 0.05 * math.sqrt(K / 64) + 1e-2 * want.abs()
bad = err > tol
print(f"{layout} bn={bn} M={M} N={N} K={K}: max_err={float(err.max()):.4g} bad={int(bad.sum())}/{bad.numel()}")
if bad.any():
    rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
    print("  bad rows (first 16):", rows[:16].tolist(), "count", len(rows))
    print("  bad cols (first 16):", cols[:16].tolist(), "count", len(cols))
    print("  out[0,:8] ", out[0,:8].float().tolist())
    print("  want[0,:8]", want[0,:8].tolist())
    # which K-slices are missing/wrong? project with one-hot K blocks
    for kb in range(0, min(K, 256), 16):
        a2 = torch.zeros_like(a); a2[:, kb:kb+16] = a[:, kb:kb+16]
        A2 = a2.t().contiguous() if a_mn else a2
        o2 = ops.gemm(A2, B, a_mn=a_mn, b_mn=b_mn, block_n=bn).float()
        w2 = a2.float() @ b.float().t()
        e2 = float((o2 - w2).abs().max())
        print(f"  k[{kb}:{kb+16}] max_err={e2:.4g}")
'''

def main():
    cfgs = []
    for layout in ("kk", "kn", "mm"):
        for bn in (128, 256, 192):
            cfgs.append((layout, bn, 256, 512, 128))
    cfgs += [("kk", 256, 2048, 1024, 1024), ("kn", 256, 2048, 1024, 1024), ("mm", 256, 1024, 1024, 2048),
             ("kn", 0, 72, 1000, 136)]
    for c in cfgs:
        try:
            r = subprocess.run([sys.executable, "-c", CHILD] + [str(x) for x in c], capture_output=True, text=True,
                               timeout=120)
            print(r.stdout.strip())
            if r.returncode != 0:
                print("  FAILED rc", r.returncode, r.stderr.strip()[-1500:])
        except subprocess.TimeoutExpired:
            print(c, "TIMEOUT (hang)")
        sys.stdout.flush()

if __name__ == "__main__":
    main()
