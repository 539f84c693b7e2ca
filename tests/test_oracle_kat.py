"""Known-answer tests that pin the oracle analytically (SURVEY.md §8c 'golden vectors to create', App. B)."""
import math

import numpy as np
import torch

import mic_b200
from mic_b200 import synthetic
from oracle import reference_model as rm
from oracle import reference_generate as rg


def test_zero_weights_loss_is_log_vocab():
    cfg = mic_b200.tiny_config(vocab_size=1003)
    params = synthetic.tree_map(np.zeros_like, synthetic.make_params(cfg))
    batch = synthetic.make_batch(cfg, 2, seq_len=8, min_len=4)
    loss, _, logits = rm.loss_and_grads(params, batch, cfg)
    assert abs(loss - math.log(1003)) < 1e-5
    assert float(np.abs(logits.numpy()).max()) == 0.0
    # full-size constant quoted in SURVEY.md
    assert abs(math.log(250054) - 12.429432) < 1e-6


def test_label_smoothing_constant():
    V, eps = 250054, 0.1
    conf, low = 1 - eps, eps / (V - 1)
    const = -(conf * math.log(conf) + (V - 1) * low * math.log(low + 1e-20))
    # uniform logits: loss = lnV - const >= 0; perfect prediction of the soft target gives 0
    soft = torch.full((1, V), low, dtype=torch.float64)
    soft[0, 7] = conf
    loss = rm.loss_fn(torch.log(soft).float(), torch.tensor([7]), torch.ones(1), eps)
    assert abs(float(loss)) < 1e-3          # fp32 sum over 250k soft-label terms
    assert const > 0


def test_int32_pixel_truncation():
    x = torch.tensor([-1.7, 0.9, 2.2, -0.2])
    assert torch.trunc(x).tolist() == [-1.0, 0.0, 2.0, -0.0]
    assert np.array([-1.7, 0.9, 2.2]).astype(np.int32).tolist() == [-1, 0, 2]


def test_top_k_tie_order():
    x = np.array([[1.0, -np.inf, 3.0, 3.0, -np.inf, 1.0]], dtype=np.float32)
    v, i = rg.top_k(x, 6)
    assert i.tolist() == [[2, 3, 0, 5, 1, 4]]


def test_fp32_penalty_rounding():
    # at 1e7 one ulp is 1.0: fp32(lp - 1e7) rounds lp to an integer (App. B)
    assert np.float32(-3.4) + rg.NEG == np.float32(-10000003.0)
    assert np.float32(np.float32(-3.4) + rg.NEG) / np.float32(5.0) == np.float32(-2000000.6)


def _gen(cfg, params, batch, **kw):
    return rg.generate(params, batch["pixel_values"], cfg, **kw)


def test_beam_step1_and_forced_eos_invariants():
    cfg = mic_b200.tiny_config(vocab_size=1003)
    params = synthetic.make_params(cfg, seed=3, perturbed=True, std=0.2)
    batch = synthetic.make_batch(cfg, 3, seq_len=8)
    L, K, fb = 12, 4, 1001
    out = _gen(cfg, params, batch, num_beams=K, max_length=L, forced_bos_token_id=fb, return_trace=True)
    tr = out["trace"]
    # step 1 is weight independent: candidates [0, -1e7 x3, -inf x4], beams 0,1,2,3 then flat idx 0..3
    t1 = tr[0]
    assert t1["cur_len"] == 1
    np.testing.assert_array_equal(t1["topk_indices"][0, :4], np.arange(4) * 1003 + fb)
    np.testing.assert_array_equal(t1["topk_indices"][0, 4:], np.arange(4))
    assert t1["topk_log_probs"][0, 0] == 0.0
    assert np.all(t1["topk_log_probs"][0, 1:4] == rg.NEG)
    assert np.all(np.isneginf(t1["topk_log_probs"][0, 4:]))
    seq = out["sequences"]
    assert seq.shape == (3, L) and seq.dtype == np.int32
    assert np.all(seq[:, 0] == 2) and np.all(seq[:, 1] == fb)
    # forced EOS at cur_len == L-1 writes eos at the last slot unless the beam finished earlier
    assert len(tr) == L - 1
    for b in range(3):
        row = seq[b]
        eos_pos = np.where(row[1:] == 2)[0]
        assert len(eos_pos) >= 1
        first = eos_pos[0] + 1
        assert np.all(row[first + 1:] == 1)        # pad after EOS
    # finished scores are fp32(fp32(lp - 1e7) / cur_len)
    assert np.all(out["scores"] < -1e7 / L - 1)


def test_greedy_pads_after_eos_and_shapes():
    cfg = mic_b200.tiny_config(vocab_size=1003)
    params = synthetic.make_params(cfg, seed=3, perturbed=True, std=0.2)
    batch = synthetic.make_batch(cfg, 2, seq_len=8)
    out = _gen(cfg, params, batch, num_beams=1, max_length=10, forced_bos_token_id=1001)
    seq = out["sequences"]
    assert seq.shape == (2, 10)
    assert np.all(seq[:, 0] == 2) and np.all(seq[:, 1] == 1001)
    # greedy writes PAD (not EOS) at the forced-EOS slot: finished flag is applied to the same token (:501-507)
    assert np.all(seq[:, -1] == 1)


def test_adamw_first_step_uses_lr_zero_under_warmup():
    p = np.ones(4, np.float32)
    g = np.full(4, 0.5, np.float32)
    lr0 = rm.linear_warmup_decay_lr(0, 5e-5, 1000, 10000)
    assert lr0 == 0.0
    p1, m1, v1 = rm.adamw_update(p, g, np.zeros(4, np.float32), np.zeros(4, np.float32), 0, lr0)
    np.testing.assert_array_equal(p1, p)
    np.testing.assert_allclose(m1, 0.05, rtol=1e-6)
    lr1 = rm.linear_warmup_decay_lr(1, 5e-5, 1000, 10000)
    p2, _, _ = rm.adamw_update(p1, g, m1, v1, 1, lr1)
    # bias-corrected update magnitude ~ 1 -> p decreases by ~lr
    np.testing.assert_allclose(p2, 1.0 - lr1, rtol=1e-5)
