"""NCCL all-reduce bandwidth on an otherwise idle box (torchrun): sizes of the gradient buckets the training step uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import mic_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
mic_b200.training.init_distributed(local)
buf = torch.zeros(512 << 20, dtype=torch.float32, device="cuda")       # 2 GiB
for mb in (16, 64, 97, 256, 353, 1024, 1836, 2048):
    n = (mb << 20) // 4
    x = buf[:n]
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if rank == 0:
        alg = mb / 1024 / (ms / 1e3) * 1.073741824
        print(f"all_reduce {mb:5d} MiB fp32: {ms:7.3f} ms  algbw {alg:6.1f} GB/s  busbw {alg * 2 * (world - 1) / world:6.1f} GB/s")
# the same 1.84 GB as 256 MiB buckets issued back to back with async_op (what train_step does)
n = (1836 << 20) // 4
x = buf[:n]
be = (256 << 20) // 4
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
hs = [dist.all_reduce(x[lo:min(n, lo + be)], async_op=True) for lo in range(0, n, be)]
for h in hs:
    h.wait()
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"1836 MiB as 256 MiB async buckets: {e0.elapsed_time(e1):.3f} ms")
dist.destroy_process_group()
