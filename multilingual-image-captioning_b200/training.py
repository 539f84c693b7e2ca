"""Training step of the reference (`main.py:684-707`) on the B200 engine.

  train_step:  loss, grad = value_and_grad(compute_loss)(params)      main.py:688-697
               grad = lax.pmean(grad, "batch")                          main.py:698   -> NCCL all-reduce / N
               state.apply_gradients(grads=grad)  (optax.adamw)         main.py:701,629-635
               metrics = pmean({"loss", "learning_rate"})               main.py:703-704

Data parallelism = one process per GPU (torch.distributed, NCCL over NVLink); every rank holds a full
replica; the gradient all-reduce is an UNWEIGHTED mean of the per-rank token-normalised gradients, as
`pmean` does (main.py:679,698).  The all-reduce runs bucket by bucket over the flat gradient buffer.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops

F32 = torch.float32


def create_learning_rate_fn(train_ds_size, train_batch_size, num_train_epochs, num_warmup_steps, learning_rate):
    """main.py:281-292: linear warm-up 0 -> lr over `num_warmup_steps`, then linear decay to 0."""
    steps_per_epoch = train_ds_size // train_batch_size
    num_train_steps = steps_per_epoch * num_train_epochs

    def schedule(step):
        if step < num_warmup_steps:
            return learning_rate * step / max(num_warmup_steps, 1)
        frac = (step - num_warmup_steps) / max(num_train_steps - num_warmup_steps, 1)
        return learning_rate * (1.0 - min(max(frac, 0.0), 1.0))
    return schedule


def init_distributed(local_rank: int, backend: str = "nccl"):
    """`torch.distributed` set-up for one process per GPU.  The NCCL streams are HIGH PRIORITY: the gradient all-reduce
    runs concurrently with backward kernels and must get SM slots as soon as they free up (see `train_step`)."""
    import os
    if dist.is_initialized():
        return
    if backend == "nccl":
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
        except Exception:           # older torch: plain defaults
            opts = None
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    else:
        dist.init_process_group(backend)


def bucketed_allreduce_sum(flat: torch.Tensor, bucket_elems: int):
    """In-place SUM all-reduce of a flat buffer in fixed-size buckets (launch-latency sized, not link sized:
    NVSwitch gives every peer full bandwidth).  Works for any backend (NCCL on GPU, gloo in the CPU tests)."""
    n = flat.numel()
    handles = []
    for lo in range(0, n, bucket_elems):
        handles.append(dist.all_reduce(flat[lo:min(n, lo + bucket_elems)], op=dist.ReduceOp.SUM, async_op=True))
    for h in handles:
        h.wait()


class AdamWConfig:
    """`optax.adamw(learning_rate=..., b1, b2, eps, weight_decay)` as the reference builds it (main.py:629-635): the
    hyper-parameters only — the update itself is the fused AdamW kernel.  Pass it as `tx=` to TrainState.create."""

    def __init__(self, learning_rate, b1=0.9, b2=0.999, eps=1e-8, weight_decay=0.0):
        if not callable(learning_rate):
            lr = float(learning_rate)
            learning_rate = lambda step: lr         # noqa: E731  (optax accepts a constant or a schedule)
        self.learning_rate, self.b1, self.b2, self.eps, self.weight_decay = learning_rate, b1, b2, eps, weight_decay


def adamw(learning_rate, b1=0.9, b2=0.999, eps=1e-8, weight_decay=0.0):
    """Drop-in for the `optax.adamw(...)` call of main.py:629-635."""
    return AdamWConfig(learning_rate, b1, b2, eps, weight_decay)


class TrainState:
    """flax TrainState analogue (main.py:247-251,638): params + AdamW state + step, living on the GPU."""

    def __init__(self, model, learning_rate_fn, b1=0.9, b2=0.999, eps=1e-8, weight_decay=0.0,
                 bucket_bytes=256 << 20, dropout=None, dropout_seed=0):
        self.model = model
        self.store = model.store
        self.store.ensure_optimizer()
        self.learning_rate_fn = learning_rate_fn
        self.b1, self.b2, self.eps, self.weight_decay = b1, b2, eps, weight_decay
        self.step = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket_elems = max(int(bucket_bytes) // 4, 1)
        self.metrics_buf = torch.zeros(2, dtype=F32, device=self.store.device)
        self.comm_stream = None
        # SMs the persistent GEMMs of the vision backward may leave free while a gradient all-reduce is in flight (one
        # tcgen05 GEMM CTA owns a whole SM).  Measured at 8 GPUs (profiles/r02_dp_timeline_n8.txt): a margin of 16 / 32
        # SMs made the step 0.6 / 1.5 ms SLOWER — the all-reduce is not SM-starved but shares HBM and the power budget
        # with the GEMMs — so the default is 0; the knob stays for other topologies.
        import os
        self.comm_sm_margin = int(os.environ.get("MIC_COMM_SM_MARGIN", "0")) if self.world > 1 else 0
        # train=True semantics (main.py:692): decoder dropout at mbart_config.dropout unless overridden;
        # one fresh mask per step and per rank (dropout_rng split / shard_prng_key, main.py:251,686)
        self.dropout = model.config.mbart_config.dropout if dropout is None else float(dropout)
        self.dropout_seed = int(dropout_seed)
        self.rank = dist.get_rank() if self.world > 1 else 0

    @classmethod
    def create(cls, apply_fn=None, params=None, tx=None, model=None, dropout_rng=None, **kw):
        """`TrainState.create(apply_fn=model.__call__, params=model.params, tx=adamw, dropout_rng=...)` (main.py:638).
        `tx` must be an AdamWConfig (mic_b200.adamw(...)): any other optimiser object would be silently replaced by
        AdamW with different hyper-parameters, so it is refused.  `apply_fn` may be the model's bound `__call__`."""
        if model is None and apply_fn is not None:
            model = getattr(apply_fn, "__self__", None)
        if model is None:
            raise ValueError("TrainState.create needs `model=` (or `apply_fn=model.__call__`)")
        if tx is not None:
            if not isinstance(tx, AdamWConfig):
                raise TypeError("tx must be mic_b200.adamw(learning_rate, b1, b2, eps, weight_decay): the step runs a fused "
                                "AdamW kernel, other optax transformations are not supported")
            kw = dict(dict(b1=tx.b1, b2=tx.b2, eps=tx.eps, weight_decay=tx.weight_decay), **kw)
            kw.setdefault("learning_rate_fn", tx.learning_rate)
        elif "learning_rate_fn" not in kw:
            raise ValueError("TrainState.create needs tx=mic_b200.adamw(...) or learning_rate_fn=")
        if params is not None:
            model._use_params_permanently(params)
        if dropout_rng is not None:
            kw.setdefault("dropout_seed", _rng_to_int(dropout_rng))
        return cls(model, kw.pop("learning_rate_fn"), **kw)

    @property
    def params(self):
        """The model's parameter tree with the reference's names (live views of the fp32 master buffer)."""
        return self.model.params

    @property
    def opt_state(self):
        """The optax chain state of `optax.adamw` as flax serialises it (main.py:313-314 `to_bytes(state.opt_state)`):
        (ScaleByAdamState(count, mu, nu), AddDecayedWeightsState(), ScaleByScheduleState(count)) -> maps "0","1","2".
        mu / nu are live views of the flat Adam buffers with the parameter tree's names.  [MEMORY: optax 0.0.9 chain]"""
        import numpy as np
        cnt = np.asarray(self.step, dtype=np.int32)
        return {"0": {"count": cnt, "mu": self.store.tree(self.store.adam_m), "nu": self.store.tree(self.store.adam_v)},
                "1": {}, "2": {"count": cnt}}

    def load_opt_state(self, opt_state, step=None):
        """Inverse of `opt_state` (restore_model_checkpoint main.py:332-346)."""
        st = opt_state["0"] if "0" in opt_state else opt_state
        for name, buf in (("mu", self.store.adam_m), ("nu", self.store.adam_v)):
            _copy_tree_into(self.store.tree(buf), st[name], self.store.device)
        self.step = int(step if step is not None else st["count"])

    def _stamp(self, name, stream=None):
        """Timeline aid (tools/dp_timeline.py): when `self.timeline` is a list, record a timing event on `stream`."""
        tl = self.__dict__.get("timeline")
        if tl is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        tl.append((name, ev))

    def allreduce_grads(self, lo=0, hi=None, async_stream=False):
        """lax.pmean(grad, 'batch') over grad[lo:hi]: SUM over ranks here, the 1/N is folded into the AdamW kernel.
        async_stream=True issues the collective on a dedicated communication stream that first waits for the
        work already enqueued on the compute stream (so it overlaps whatever is enqueued afterwards)."""
        if self.world == 1:
            return
        g = self.store.grad
        hi = g.numel() if hi is None else hi
        if not async_stream:
            bucketed_allreduce_sum(g[lo:hi], self.bucket_elems)
            return
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.store.device, priority=-1)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            self._stamp(f"allreduce[{lo}:{hi}] start", self.comm_stream)
            bucketed_allreduce_sum(g[lo:hi], self.bucket_elems)
            self._stamp(f"allreduce[{lo}:{hi}] end", self.comm_stream)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        return done

    def wait_comm(self):
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def apply_gradients(self, segments=None):
        """optax.adamw with the schedule evaluated at the pre-increment count.
        segments: optional [(lo, hi, event)] — AdamW runs slice by slice, each after its all-reduce event, so the
        optimiser of the first bucket overlaps the communication of the second."""
        count = self.step
        t = count + 1
        lr = float(self.learning_rate_fn(count))
        # scalars by value: the host may run many steps ahead of the stream (no per-step sync), so nothing the
        # enqueued update reads may live in a buffer the next step rewrites
        hp = (lr, self.b1, self.b2, self.eps, self.weight_decay, 1.0 / (1.0 - self.b1 ** t), 1.0 / (1.0 - self.b2 ** t),
              1.0 / self.world)
        s = self.store
        if segments is None:
            ops.adamw(s.master, s.adam_m, s.adam_v, s.grad, s.shadow, *hp)
        else:
            cur = torch.cuda.current_stream()
            for lo, hi, ev in segments:
                if ev is not None:
                    cur.wait_event(ev)
                ops.adamw(s.master[lo:hi], s.adam_m[lo:hi], s.adam_v[lo:hi], s.grad[lo:hi], s.shadow[lo:hi], *hp)
                self._stamp(f"adamw[{lo}:{hi}] end")
        self.step = t
        return lr


def _rng_to_int(rng):
    """A PRNG key (jax uint32[2], numpy array, torch tensor or int) folded to one seed for the counter-hash dropout."""
    import numpy as np
    if hasattr(rng, "detach"):
        rng = rng.detach().cpu().numpy()
    a = np.asarray(rng).astype(np.uint64).reshape(-1)
    v = 0
    for x in a:
        v = (v * 1000003 + int(x)) & 0x3FFFFFFF
    return int(v)


def _copy_tree_into(dst, src, device):
    import numpy as np
    for k, v in dst.items():
        if k not in src:
            raise KeyError(f"missing entry {k}")
        if isinstance(v, dict):
            _copy_tree_into(v, src[k], device)
        else:
            s = src[k]
            s = torch.from_numpy(np.ascontiguousarray(s)) if isinstance(s, np.ndarray) else s
            if tuple(s.shape) != tuple(v.shape):
                raise ValueError(f"shape mismatch at {k}: {tuple(s.shape)} vs {tuple(v.shape)}")
            v.copy_(s.to(device, torch.float32))


def save_model_checkpoint(model, save_dir, state, with_opt: bool = False, overwrite: bool = False, logger=None, **kwargs):
    """main.py:299-328: <save_dir>/ckpt-<step-1>/{config.json, flax_model.msgpack[, opt_state.msgpack,
    training_state.json]} in the reference's own container format (checkpoint.py).  Hub push is out of scope."""
    import json
    import os
    from . import checkpoint as ck
    ckpt_save_dir = f"{save_dir}/ckpt-{int(state.step) - 1}"
    if os.path.exists(ckpt_save_dir) and not overwrite:
        if logger:
            logger.info("checkpoint exists, skipping overwrite")
        return ckpt_save_dir
    model.save_pretrained(ckpt_save_dir, params=state.params)
    if with_opt:
        with open(os.path.join(ckpt_save_dir, "opt_state.msgpack"), "wb") as f:
            f.write(ck.to_bytes(state.opt_state))
        with open(os.path.join(ckpt_save_dir, "training_state.json"), "w") as f:
            json.dump({"step": int(state.step)}, f)
    return ckpt_save_dir


def restore_model_checkpoint(save_dir, state, logger=None):
    """main.py:332-346: returns (params, opt_state, step) restored INTO the structure of `state` (flax from_bytes
    semantics: the key sets must match)."""
    import json
    import os
    from . import checkpoint as ck
    with open(os.path.join(save_dir, ck.FLAX_WEIGHTS_NAME), "rb") as f:
        params = ck.from_bytes(state.params, f.read())
    with open(os.path.join(save_dir, "opt_state.msgpack"), "rb") as f:
        opt_state = ck.from_bytes(state.opt_state, f.read())
    with open(os.path.join(save_dir, "training_state.json")) as f:
        step = json.load(f)["step"]
    return params, opt_state, step


def rotate_checkpoints(ckpt_dir: str, save_total_limit: int, logger=None):
    """main.py:348-357."""
    import shutil
    from pathlib import Path
    ckpts = sorted((str(x) for x in Path(ckpt_dir).glob("ckpt-*")), key=lambda x: int(x.split("-")[-1]))
    for ckpt in ckpts[:-save_total_limit] if save_total_limit > 0 else []:
        shutil.rmtree(ckpt)


_KEYS = ("pixel_values", "decoder_input_ids", "attention_mask", "input_ids")


def _static_batch(state, batch):
    """Persistent device buffers for the step inputs (so the captured graph sees fixed addresses).

    Host batches travel on a dedicated copy stream into one of two landing buffers, so the H2D transfer of
    step i overlaps the GPU work of step i-1 that is still in flight; the compute stream then takes a cheap
    device-to-device copy into the static inputs."""
    dev = state.store.device
    px = torch.as_tensor(batch["pixel_values"])
    # uint8 pixels stay uint8 all the way to the patch kernel (input hand-off: a quarter of the H2D bytes)
    px_dtype = torch.uint8 if px.dtype == torch.uint8 else F32
    key = (tuple(px.shape), tuple(torch.as_tensor(batch["decoder_input_ids"]).shape), px_dtype)
    sb = state.__dict__.get("_static")
    if sb is None or sb["key"] != key:
        B, T = key[1]

        def mk():
            return {"pixel_values": torch.empty(key[0], dtype=px_dtype, device=dev),
                    "decoder_input_ids": torch.empty((B, T), dtype=torch.int32, device=dev),
                    "attention_mask": torch.empty((B, T), dtype=torch.int32, device=dev),
                    "input_ids": torch.empty((B, T), dtype=torch.int32, device=dev)}
        sb = {"key": key, "graph": None, "land": [mk(), mk()], "slot": 0, "copy_stream": torch.cuda.Stream(device=dev),
              "consumed": [torch.cuda.Event(), torch.cuda.Event()]}
        sb.update(mk())
        state._static = sb
    cur = torch.cuda.current_stream()
    on_host = not torch.as_tensor(batch["pixel_values"]).is_cuda
    if on_host:
        slot = sb["slot"] = sb["slot"] ^ 1
        land, cs = sb["land"][slot], sb["copy_stream"]
        cs.wait_event(sb["consumed"][slot])              # the previous reader of this landing slot is done
        with torch.cuda.stream(cs):
            for k in _KEYS:
                src = torch.as_tensor(batch[k])
                if land[k].dtype != src.dtype:          # land in the SOURCE dtype: pure async memcpy, no host-side cast
                    land[k] = torch.empty(src.shape, dtype=src.dtype, device=dev)
                land[k].copy_(src, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        cur.wait_event(ready)
        for k in _KEYS:
            sb[k].copy_(land[k], non_blocking=True)
        sb["consumed"][slot].record(cur)
    else:
        for k in _KEYS:
            sb[k].copy_(torch.as_tensor(batch[k]), non_blocking=True)
    return sb


@torch.no_grad()
def train_step(state: TrainState, batch, label_smoothing_factor: float = 0.0, use_cuda_graph: bool = True):
    """One optimisation step.  batch keys as the reference's collate output (main.py:493-523):
    pixel_values (B,H,W,3) f32, input_ids = labels (B,T), attention_mask (B,T), decoder_input_ids (B,T).
    Returns (state, metrics) with metrics = {"loss": 0-d tensor, "learning_rate": float}.

    The forward+backward kernel sequence is shape-static, so after one eager step it is captured into a
    CUDA graph and replayed (no per-kernel host launch cost); the NCCL all-reduce and AdamW stay outside."""
    eng = state.model.engine
    eng.dropout_p = state.dropout
    if state.dropout > 0.0:
        mix = (state.dropout_seed * 1000003 + state.step * 7919 + state.rank * 104729 + 12345) & 0x3FFFFFFF
        eng.drop_seed.fill_(mix)      # scalar travels as a kernel argument (fixed at enqueue time; the graph reads the tensor)
    sb = _static_batch(state, batch)
    args = (sb["pixel_values"], sb["decoder_input_ids"], sb["attention_mask"], sb["input_ids"])
    ls = (label_smoothing_factor, state.dropout)
    dp = state.world > 1
    # data parallel: three graph segments.  After each one a prefix of the flat gradient buffer is final (it is laid out
    # in backward order) and its all-reduce starts on the communication stream while the next segment computes:
    #   1: forward + lm_head/CE + decoder + cross-K/V backward   -> 84 % of the bytes (tied embedding, decoder)
    #   2: visual projection + upper vision layers               -> 11 %
    #   3: last `dp_vision_tail_layers` vision layers + embeddings -> 5 %: the only all-reduce nothing can hide
    stages = (1, 2, 3) if dp else (0,)
    bounds = ([0] + eng.grad_split_offsets() + [state.store.grad.numel()]) if dp else None

    def run_stage(stage):
        # segments 2 and 3 run under an all-reduce: their GEMMs may leave `comm_sm_margin` SMs to NCCL (baked into the
        # captured graph; 0 = off, the measured optimum on NVSwitch boxes: profiles/r02_dp_timeline_n8.txt)
        margin = state.comm_sm_margin if (dp and stage >= 2) else 0
        ops.launch_options(gemm_sm_margin=margin)
        try:
            return eng.forward_backward(*args, label_smoothing=label_smoothing_factor, stage=stage)
        finally:
            ops.launch_options(gemm_sm_margin=0)

    evs = []

    def after(stage):
        # lax.pmean of everything this segment finished, on the communication stream
        if dp:
            i = stages.index(stage)
            evs.append(state.allreduce_grads(bounds[i], bounds[i + 1], async_stream=True))

    if not use_cuda_graph:
        ws = None
        for st in stages:
            r = run_stage(st)
            ws = r if r is not None else ws
            after(st)
    elif sb["graph"] is None or sb.get("ls") != ls:
        ws = None
        for st in stages:                                           # eager: allocates all buffers
            r = run_stage(st)
            ws = r if r is not None else ws
            after(st)
        if sb.get("warm"):
            state.wait_comm()
            torch.cuda.synchronize()
            graphs = []
            for st in stages:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    r = run_stage(st)
                if r is not None:
                    sb["ws"] = r
                graphs.append(g)
            sb["graph"], sb["ls"] = graphs, ls
            ws = sb["ws"]
            evs.clear()
            for i, g in enumerate(graphs):
                g.replay()
                after(stages[i])
        sb["warm"] = True
    else:
        state._stamp("step start")
        for i, g in enumerate(sb["graph"]):
            g.replay()
            state._stamp(f"graph stage {stages[i]} end")
            after(stages[i])
        ws = sb["ws"]
    if dp:
        # AdamW streams 16 GB at the full HBM rate: an all-reduce that runs next to it crawls (measured at 8 GPUs: the
        # 0.35 GB tail took 5-8 ms beside AdamW against 1.0 ms alone).  So the optimiser starts only when the LAST
        # bucket has arrived — the collectives are serialised on the communication stream, so that event covers all.
        torch.cuda.current_stream().wait_event(evs[-1])
        state._stamp("all gradients reduced")
        lr = state.apply_gradients()
        state._stamp("adamw end")
    else:
        lr = state.apply_gradients()
    loss = ws["out"][0:1].clone()
    if state.world > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        loss /= state.world
    return state, {"loss": loss[0], "learning_rate": lr}


@torch.no_grad()
def eval_step(model, batch, label_smoothing_factor: float = 0.0):
    """main.py:710-721."""
    loss = model.loss(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], batch["input_ids"],
                      label_smoothing_factor)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        loss = loss / dist.get_world_size()
    return {"loss": loss}
