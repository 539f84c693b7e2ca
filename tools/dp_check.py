"""2-rank data-parallel correctness check (torchrun, NCCL): replicas stay identical and equal a single-process
step whose gradient is the UNWEIGHTED mean of the per-shard gradients (lax.pmean, main.py:679,698)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import mic_b200
from mic_b200 import synthetic, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
mic_b200.training.init_distributed(local)
cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
params = synthetic.make_params(cfg, seed=1, perturbed=True, std=0.05)
shards = [synthetic.make_batch(cfg, 4, seq_len=16, seed=10 + r, min_len=3 + 5 * r) for r in range(world)]
sched = mic_b200.create_learning_rate_fn(1000, 8, 1, 0, 1e-2)

model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg)
model.params = params
state = mic_b200.TrainState(model, sched, weight_decay=0.01, bucket_bytes=1 << 16, dropout=0.0)
losses = []
for step in range(3):      # eager, warm, graph-replay paths
    state, m = mic_b200.train_step(state, shards[rank])
    losses.append(float(m["loss"]))
torch.cuda.synchronize()
mine = model.store.master.clone()
other = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(other, mine)
if rank == 0:
    for r in range(1, world):
        assert torch.equal(other[0], other[r]), "replicas diverged"
    # single-process reference: mean of shard gradients
    ref = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg)
    ref.params = params
    rstate = mic_b200.TrainState(ref, sched, weight_decay=0.01, dropout=0.0)
    rstate.world = 1
    ref_losses = []
    for step in range(3):
        gs, ls = [], []
        for sh in shards:
            ws = ref.engine.forward_backward(torch.from_numpy(sh["pixel_values"]), torch.from_numpy(sh["decoder_input_ids"]),
                                             torch.from_numpy(sh["attention_mask"]), torch.from_numpy(sh["input_ids"]), 0.0)
            gs.append(ref.store.grad.clone()); ls.append(float(ws["out"][0]))
        ref.store.grad.copy_(sum(gs) / len(gs))
        rstate.apply_gradients()
        ref_losses.append(sum(ls) / len(ls))
    torch.cuda.synchronize()
    diff = (ref.store.master - mine).abs().max().item()
    moved = (ref.store.master - torch.cat([torch.from_numpy(v).flatten() for _, v in synthetic.tree_flatten(params)])[:1].cuda()).abs().max().item()
    print("DP check: replicas identical; max |param diff| vs mean-gradient reference =", diff,
          "losses", losses, "ref", ref_losses)
    assert diff < 2e-3, diff
    assert all(abs(a - b) < 1e-3 for a, b in zip(losses, ref_losses))
    print("DP_CHECK_OK")
dist.barrier()
dist.destroy_process_group()
