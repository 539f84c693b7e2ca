"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the host-side
parameter layout reproduces the reference's pytree, schedules, and the data-parallel reduction rule."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import mic_b200
from mic_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mic_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mic_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    l = _lib.lib()
    declared = _header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/mic_b200.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.EXPORTED) == declared
    assert l.mic_abi_version() == 2


def test_sass_contains_blackwell_tensor_and_tma_ops():
    from mic_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in r.stdout, mnemonic


def test_layout_matches_reference_pytree_and_param_count():
    from mic_b200.params import Layout
    full = Layout(mic_b200.clip_mbart_config())
    total = sum(int(np.prod(shape)) for _, shape in full.storages.values())
    assert total == 547_163_590
    cfg = mic_b200.tiny_config()
    lay = Layout(cfg)
    want = {"/".join(p): v.shape for p, v in synthetic.tree_flatten(synthetic.make_params(cfg))}
    have = {}
    for path, (storage, slicer, reshape) in lay.leaves.items():
        shape = lay.storages[storage][1]
        arr = np.zeros(shape, np.int8)
        if slicer is not None:
            arr = arr[slicer]
        if reshape is not None:
            arr = arr.reshape(reshape)
        have["/".join(path)] = arr.shape
    assert have == want
    # every storage offset is 64-element aligned (TMA / vector access)
    assert all(off % 64 == 0 for off, _ in lay.storages.values())


def test_learning_rate_schedule_matches_oracle():
    from oracle import reference_model as rm
    f = mic_b200.create_learning_rate_fn(10000, 10, 7, 1000, 5e-5)
    for s in (0, 1, 999, 1000, 1001, 3500, 6999, 7000, 9000):
        assert abs(f(s) - rm.linear_warmup_decay_lr(s, 5e-5, 1000, 7000)) < 1e-15


def test_shift_tokens_right_and_batch_contract():
    cfg = mic_b200.clip_mbart_config()
    b = synthetic.make_batch(cfg, 4, 64, seed=0)
    assert b["pixel_values"].shape == (4, 224, 224, 3) and b["pixel_values"].dtype == np.float32
    assert np.all(b["decoder_input_ids"][:, 0] == 1)
    assert np.array_equal(b["decoder_input_ids"][:, 1:], b["input_ids"][:, :-1])
    assert np.array_equal(b["attention_mask"], (b["input_ids"] != 1).astype(np.int64))
    assert set(b["input_ids"][:, 0]) <= set(synthetic.LANG_CODES)


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import mic_b200
from mic_b200.training import bucketed_allreduce_sum
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
base = torch.randn(1000)
g = base * (rank + 1)                      # per-rank token-normalised gradient
bucketed_allreduce_sum(g, 96)              # ragged last bucket
mean = g / world                           # 1/N folded into the optimiser kernel
want = base * sum(r + 1 for r in range(world)) / world   # unweighted pmean (main.py:698)
assert torch.allclose(mean, want, atol=1e-6), float((mean - want).abs().max())
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_data_parallel_reduction_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_gradient_buffer_is_laid_out_in_backward_order_for_the_three_dp_segments():
    """training.train_step all-reduces a PREFIX of the flat gradient after each backward segment (engine.
    grad_split_offsets): the layout must store, in this order, head + tied embedding + decoder + cross K/V, then the
    visual projection and the upper vision layers, then the last `dp_vision_tail_layers` layers and the embeddings."""
    from mic_b200.params import Layout
    from mic_b200.engine import CaptionEngine
    lay = Layout(mic_b200.clip_mbart_config())
    off = {k: v[0] for k, v in lay.storages.items()}
    k = CaptionEngine.dp_vision_tail_layers
    s1, s2 = off["proj.w"], off[f"v.{k - 1}.fc2.w"]
    seg1 = [n for n in lay.order if off[n] < s1]
    seg2 = [n for n in lay.order if s1 <= off[n] < s2]
    seg3 = [n for n in lay.order if off[n] >= s2]
    assert seg1[0] == "flb" and seg1[1] == "shared" and all(n.startswith(("flb", "shared", "d.")) for n in seg1)
    assert all(n.startswith("proj.") or n.startswith("v.post_ln") or
               (n.startswith("v.") and n.split(".")[1].isdigit() and int(n.split(".")[1]) >= k) for n in seg2), seg2[:5]
    assert all((n.split(".")[1].isdigit() and int(n.split(".")[1]) < k) or n.split(".")[1] in ("pre_ln", "pos", "cls", "patch")
               for n in seg3), seg3
    # decoder layers appear in backward order (11 first), so their gradients are final in that order
    dec = [int(n.split(".")[1]) for n in seg1 if n.startswith("d.") and n.split(".")[1].isdigit()]
    assert dec == sorted(dec, reverse=True)
    total = lay.size * 4
    assert 0.80 < s1 * 4 / total < 0.88 and (lay.size - s2) * 4 / total < 0.07
