"""CUDA-event timeline of one data-parallel training step (torchrun, NCCL): when each gradient bucket's all-reduce
starts and ends relative to the backward graph segments and the AdamW slices (name the residual).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_timeline.py [steps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import mic_b200
from mic_b200 import synthetic

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    mic_b200.training.init_distributed(local)
cfg = mic_b200.clip_mbart_config()
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0, device=torch.device("cuda", local))
state = mic_b200.TrainState(model, mic_b200.create_learning_rate_fn(10_000_000, 256 * world, 7, 1000, 5e-5))
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(cfg, 256, 64, seed=2 + rank).items()}
for _ in range(4):
    mic_b200.train_step(state, batch)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rows = []
for it in range(steps):
    state.timeline = []
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    mic_b200.train_step(state, batch)
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    torch.cuda.synchronize()
    tl = [(n, e0.elapsed_time(ev)) for n, ev in state.timeline] + [("step end (host-enqueued work done)", e0.elapsed_time(e1))]
    rows.append(tl)
state.timeline = None
# steady-state step time without stamps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    mic_b200.train_step(state, batch)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# one compact line per rank: do all GPUs reach the collectives at the same time, or does one straggler gate them?
names = [n for n, _ in rows[-1]]
med = [sorted(row[i][1] for row in rows if len(row) == len(names))[len(rows) // 2] for i in range(len(names))]
mine = torch.tensor(med, dtype=torch.float64, device="cuda")
allr = [torch.empty_like(mine) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, mine)
else:
    allr = [mine]
if rank == 0:
    print("per-rank medians (ms after step start):")
    for i, n in enumerate(names):
        print(f"  {n:48s} " + " ".join(f"{float(a[i]):7.2f}" for a in allr))
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank and rank in (0, world - 1):
        print(f"--- rank {rank}/{world}: ms after step start (last of {steps} steps; median over steps in brackets)")
        names = [n for n, _ in rows[-1]]
        for i, n in enumerate(names):
            vals = sorted(row[i][1] for row in rows if len(row) == len(names))
            print(f"  {rows[-1][i][1]:8.2f}  [{vals[len(vals) // 2]:8.2f}]  {n}")
if rank == 0:
    print(f"steady-state step: {float(ms):.2f} ms at {world} GPU(s) = {256 * world / float(ms) * 1e3:.0f} samples/s")
if world > 1:
    dist.destroy_process_group()
