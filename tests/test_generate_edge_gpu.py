"""Beam / greedy edge cases against the oracle: early EOS (finished-set merge, -1e7 reuse, early termination
of the while_loop), min_length, length_penalty, early_stopping=False, 1-3 beams, short max_length, no forced
tokens.  EOS is made likely through final_logits_bias so finished hypotheses actually occur."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import synthetic  # noqa: E402
from oracle import reference_generate as rg  # noqa: E402

_MODEL = {}
_STATS = {"rows": 0, "clear": 0, "gpos": 0, "gcmp": 0}


def _setup(eos_bias, seed=5):
    cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
    # moderate weight scale: logits of O(3) keep the bf16-vs-fp32 logit error (~0.03) far below the margins used
    params = synthetic.make_params(cfg, seed=seed, perturbed=True, std=0.12)
    params["final_logits_bias"] = params["final_logits_bias"].copy()
    params["final_logits_bias"][0, 2] += eos_bias
    if "m" not in _MODEL:
        _MODEL["m"] = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg)
    model = _MODEL["m"]
    model.params = params
    batch = synthetic.make_batch(cfg, 5, seq_len=16, seed=seed)
    return cfg, params, batch, model


def _clear_rows(ref, L, tol=0.25):
    """Rows whose every beam decision has a margin above the bf16 noise: the top-9 raw candidate scores are
    pairwise separated (set AND order of the kept 2K unambiguous) and no finished candidate sits near the .5
    rounding boundary of fp32(lp - 1e7)."""
    clear = np.ones(ref["sequences"].shape[0], bool)
    np.seterr(invalid="ignore")
    for step in ref["trace"]:
        raw = step["topk_raw"]
        allv = np.concatenate([raw, step["ninth"][:, None]], 1)
        gaps = np.abs(np.diff(allv, axis=1))
        finite = np.isfinite(allv[:, :-1]) & np.isfinite(allv[:, 1:])
        big = np.abs(allv[:, :-1]) > 1e6                      # -1e7-level entries: ulp is 1.0, exact ties are stable
        clear &= np.all((gaps > tol) | ~finite | big, axis=1)
        frac = np.abs(raw - np.floor(raw) - 0.5)
        near = step["did_finish"] & np.isfinite(raw) & (np.abs(raw) < 1e6) & (frac < tol)
        clear &= ~near.any(axis=1)
    return clear


CASES = [
    dict(num_beams=4, max_length=12, forced_bos_token_id=1001),
    dict(num_beams=4, max_length=12),                                        # no forced BOS
    dict(num_beams=4, max_length=12, forced_bos_token_id=1001, min_length=6),
    dict(num_beams=4, max_length=12, forced_bos_token_id=1001, length_penalty=2.0),
    dict(num_beams=4, max_length=12, forced_bos_token_id=1001, early_stopping=False),
    dict(num_beams=2, max_length=9, forced_bos_token_id=1001),
    dict(num_beams=3, max_length=6, forced_bos_token_id=1001),
    dict(num_beams=4, max_length=3, forced_bos_token_id=1001),
    dict(num_beams=4, max_length=16, forced_bos_token_id=1001, forced_eos_token_id=None) if False else
    dict(num_beams=4, max_length=16, forced_bos_token_id=1001),
]


@pytest.mark.parametrize("eos_bias", [0.0, 4.0, 8.0])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_beam_search_edge_cases(case, eos_bias):
    kw = CASES[case]
    cfg, params, batch, model = _setup(eos_bias)
    ref = rg.generate(params, batch["pixel_values"], cfg, return_trace=True, **kw)
    out = model.generate(batch["pixel_values"], **kw)
    seq, sc = out.sequences.cpu().numpy(), out.scores.cpu().numpy()
    L = kw["max_length"]
    assert seq.shape == ref["sequences"].shape == (5, L)
    clear = _clear_rows(ref, L)
    n_eos_early = int(((ref["sequences"][:, 1:-1] == 2).any(1)).sum())
    _STATS["rows"] += 5
    _STATS["clear"] += int(clear.sum())
    for b in range(5):
        if clear[b]:
            np.testing.assert_array_equal(seq[b], ref["sequences"][b], err_msg=f"case {kw} bias {eos_bias} row {b}")
            assert abs(sc[b] - ref["scores"][b]) <= 2e-3 * max(1.0, abs(ref["scores"][b])), (sc[b], ref["scores"][b])
    # the oracle itself must exercise the early-EOS machinery at high bias, otherwise the case tests nothing
    if eos_bias >= 8.0 and case == 0:
        assert n_eos_early >= 1


@pytest.mark.parametrize("eos_bias", [0.0, 6.0])
def test_greedy_edge_cases(eos_bias):
    cfg, params, batch, model = _setup(eos_bias, seed=7)
    for kw in (dict(num_beams=1, max_length=12, forced_bos_token_id=1001), dict(num_beams=1, max_length=12),
               dict(num_beams=1, max_length=12, forced_bos_token_id=1001, min_length=5), dict(num_beams=1, max_length=2)):
        ref = rg.generate(params, batch["pixel_values"], cfg, return_trace=True, **kw)
        seq = model.generate(batch["pixel_values"], **kw).sequences.cpu().numpy()
        L = kw["max_length"]
        margins = ref["margins"]
        for b in range(5):
            for pos in range(1, L):
                _STATS["gpos"] += 1
                if pos - 1 < margins.shape[1] and np.isfinite(margins[b, pos - 1]) and margins[b, pos - 1] < 0.3:   # bf16 + split-K summation-order noise on logits of O(3)
                    break
                _STATS["gcmp"] += 1
                assert seq[b, pos] == ref["sequences"][b, pos], (kw, eos_bias, b, pos, seq[b], ref["sequences"][b])


def test_batch_of_one_and_odd_sizes():
    cfg, params, batch, model = _setup(0.0)
    for B in (1, 3):
        px = batch["pixel_values"][:B]
        ref = rg.generate(params, px, cfg, num_beams=4, max_length=8, forced_bos_token_id=1001)
        seq = model.generate(px, num_beams=4, max_length=8, forced_bos_token_id=1001).sequences.cpu().numpy()
        assert seq.shape == (B, 8)
        assert np.all(seq[:, :2] == ref["sequences"][:, :2])


def test_zz_edge_cases_were_not_vacuous():
    """Runs last in this file: a healthy share of rows / positions had decision margins above the tolerance."""
    assert _STATS["rows"] > 0 and _STATS["clear"] / _STATS["rows"] > 0.02, _STATS   # exactness: test_beam_kernels_gpu.py
    assert _STATS["gpos"] > 0 and _STATS["gcmp"] / _STATS["gpos"] > 0.3, _STATS
