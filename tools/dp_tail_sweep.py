"""Steady-state data-parallel step time as a function of `dp_vision_tail_layers` (how many encoder layers' gradients ride in
the last, exposed all-reduce bucket).  One process group, the model rebuilt per setting.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/dp_tail_sweep.py 4,2,1,0
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
tails = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4,1").split(",")]
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(cfg, 256, 64, seed=2 + rank).items()}
model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0, device=torch.device("cuda", local))
for rep in range(2):
    for tail in tails:
        model.engine.dp_vision_tail_layers = tail
        state = mic_b200.TrainState(model, mic_b200.create_learning_rate_fn(10_000_000, 256 * world, 7, 1000, 5e-5))
        for _ in range(4):
            mic_b200.train_step(state, batch)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(12):
            mic_b200.train_step(state, batch)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 12], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"rep {rep} dp_vision_tail_layers={tail}: {float(ms):.2f} ms/step at {world} GPUs = {256 * world / float(ms) * 1e3:.0f} samples/s", flush=True)
        del state
if world > 1:
    dist.destroy_process_group()
