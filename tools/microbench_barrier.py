"""Cost of one grid barrier of the persistent decoder-step kernel (148 CTAs, L2 atomic + polling)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mic_b200
from mic_b200 import _lib

lib = _lib.lib()
sync = torch.zeros(256, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
names = {0: "counter (atom.add.release)", 1: "poll without sleep", 2: "+ __threadfence per thread", 3: "256 pollers per CTA",
         4: "per-CTA flag words, warp polls"}
for variant in (0, 1, 2, 3, 4):
    for n in (2, 1002):
        for _ in range(3):
            lib.mic_barrier_bench(s, sync.data_ptr(), n, variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.mic_barrier_bench(s, sync.data_ptr(), n, variant)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e3
        if n == 2:
            base = t
        else:
            print(f"variant {variant} ({names[variant]:28s}): {(t - base) / 1000:.3f} us per barrier (kernel with 2 barriers: {base:.1f} us)")
assert int(sync[0].item()) == 0
