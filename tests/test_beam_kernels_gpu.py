"""Bit-exact check of the beam-search kernels (search state machine of generation_clip_vision_utils.py:822-990)
against the oracle on IDENTICAL float32 log-probs: the decoder is replaced on both sides by a deterministic
function of (step, token history), so every comparison is exact — sequences, fp32 scores (incl. the -1e7
re-use and 1-ulp rounding), finished flags, ancestor tables, early termination."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import ops  # noqa: E402
from oracle import reference_generate as rg  # noqa: E402

V = 37
I32, F32 = torch.int32, torch.float32


def make_logits_fn(seed, eos_boost, quantize):
    def fn(cur_len, seqs):
        out = np.empty((seqs.shape[0], V), np.float32)
        for r in range(seqs.shape[0]):
            h = hashlib.sha256((str(seed) + ":" + str(cur_len) + ":" + ",".join(map(str, seqs[r, :cur_len]))).encode()).digest()
            rng = np.random.default_rng(int.from_bytes(h[:8], "little"))
            x = rng.standard_normal(V).astype(np.float32) * 2.0
            if quantize:                      # provoke exact ties
                x = np.round(x)
            x[2] += eos_boost
            out[r] = x
        return out
    return fn


def cuda_beam_search(logits_fn, B, K, L, *, eos=2, pad=1, start=2, forced_bos=None, forced_eos=2, min_length=0,
                     length_penalty=1.0, early_stopping=True):
    dev = "cuda"
    R = B * K
    st = {"running_seq": torch.full((B, K, L), pad, dtype=I32, device=dev),
          "sequences": torch.full((B, K, L), pad, dtype=I32, device=dev),
          "running_scores": torch.full((B, K), -1.0e7, dtype=F32, device=dev),
          "scores": torch.full((B, K), -1.0e7, dtype=F32, device=dev),
          "finished": torch.zeros((B, K), dtype=I32, device=dev),
          "ancestors": torch.arange(R, dtype=I32, device=dev)[:, None].expand(R, L).contiguous(),
          "next_token": torch.full((R,), start, dtype=I32, device=dev),
          "active": torch.ones(1, dtype=I32, device=dev)}
    st["running_seq"][:, :, 0] = start
    st["running_scores"][:, 0] = 0.0
    cpr = 8 if K <= 4 else 16                     # candidates per row handed to the beam step (>= 2K)
    ws = {"row_lp": torch.empty((R, cpr), dtype=F32, device=dev), "row_tok": torch.empty((R, cpr), dtype=I32, device=dev)}
    anc_log = []
    for cur_len in range(1, L):
        forced = -1
        if forced_bos is not None and cur_len == 1:
            forced = forced_bos
        if forced_eos is not None and cur_len == L - 1:
            forced = forced_eos
        if forced < 0:
            seqs = st["running_seq"].reshape(R, L).cpu().numpy()
            lp = rg.log_softmax(logits_fn(cur_len, seqs))
            if min_length is not None and min_length > -1 and cur_len <= min_length:
                lp[:, eos] = -np.inf
            val, idx = rg.top_k(lp, cpr)
            ws["row_lp"].copy_(torch.from_numpy(val))
            ws["row_tok"].copy_(torch.from_numpy(idx.astype(np.int32)))
        ops.beam_step(ws, st, forced, B, K, L, V, cur_len, eos, early_stopping, length_penalty)
        ops.beam_cond(st, B, K, cur_len + 1, L, length_penalty, early_stopping)
        anc_log.append(st["ancestors"].cpu().numpy().copy())
    out_seq = torch.empty((B, L), dtype=I32, device=dev)
    out_sc = torch.empty((B,), dtype=F32, device=dev)
    ops.beam_finalize(st, B, K, L, out_seq, out_sc)
    torch.cuda.synchronize()
    return out_seq.cpu().numpy(), out_sc.cpu().numpy(), {k: v.cpu().numpy() for k, v in st.items()}


class _Cfg:
    class mbart_config:
        vocab_size = V
        decoder_layers = 1


@pytest.mark.parametrize("K", [2, 3, 4, 5, 8])
@pytest.mark.parametrize("L", [3, 7, 16])
@pytest.mark.parametrize("eos_boost", [0.0, 3.0, 6.0])
@pytest.mark.parametrize("variant", ["plain", "ties", "no_forced_bos", "lp2", "no_early", "minlen"])
def test_beam_state_machine_bit_exact(K, L, eos_boost, variant):
    B = 3
    kw = dict(forced_bos=11, forced_eos=2, min_length=0, length_penalty=1.0, early_stopping=True)
    if variant == "no_forced_bos":
        kw["forced_bos"] = None
    if variant == "lp2":
        kw["length_penalty"] = 2.0
    if variant == "no_early":
        kw["early_stopping"] = False
    if variant == "minlen":
        kw["min_length"] = 4
    fn = make_logits_fn(seed=K * 100 + L, eos_boost=eos_boost, quantize=(variant == "ties"))
    g = dict(num_beams=K, max_length=L, pad_token_id=1, eos_token_id=2, decoder_start_token_id=2,
             min_length=kw["min_length"], forced_bos_token_id=kw["forced_bos"], forced_eos_token_id=kw["forced_eos"],
             length_penalty=kw["length_penalty"], early_stopping=kw["early_stopping"])
    ref = rg._beam_search(None, None, _Cfg, g, return_trace=True, logits_fn=fn, batch_size=B)
    seq, sc, st = cuda_beam_search(fn, B, K, L, **kw)
    np.testing.assert_array_equal(seq, ref["sequences"])
    np.testing.assert_array_equal(sc.view(np.int32), ref["scores"].view(np.int32))      # bit-exact fp32 scores
    # the whole carried state after the loop ended (the device `active` flag froze it at the oracle's exit step)
    for k in ("running_seq", "sequences", "finished"):
        np.testing.assert_array_equal(st[k], ref["state"][k], err_msg=k)
    for k in ("running_scores", "scores"):
        np.testing.assert_array_equal(st[k].view(np.int32), ref["state"][k].view(np.int32), err_msg=k)
