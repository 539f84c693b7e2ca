"""Persistent decoder-step kernel (csrc/decoder_step.cu) against the per-op cached decode path it replaces
(both implement FlaxMBartDecoderLayer with past_key_values, modeling_clip_vision_mbart.py:519-651)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import generation as gen, synthetic  # noqa: E402

I32 = torch.int32


def _run(engine, fused, enc_kv, R, T, K, tokens, anc_tables):
    cache = gen.DecodeCache(engine, R, T, enc_kv, K, use_ancestors=True)
    cache.self_kv.zero_()
    outs = []
    for pos in range(tokens.shape[0]):
        cache.ancestors.copy_(anc_tables[pos])
        step = gen.decode_step_fused if fused else gen.decode_step
        hf = step(engine, cache, tokens[pos], pos)
        outs.append(hf.float().clone())
    torch.cuda.synchronize()
    kv = (gen.fused_cache_rowmajor(engine, cache) if fused else cache.self_kv).float().clone()
    return torch.stack(outs), kv


@pytest.mark.parametrize("B,K", [(6, 4), (40, 4), (3, 1), (1, 1), (80, 4), (13, 8)])
def test_fused_decoder_matches_per_op_path(B, K):
    cfg = mic_b200.tiny_config()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=3)
    model.params = synthetic.make_params(cfg, seed=5, perturbed=True, std=0.08)
    eng = model.engine
    T, R, steps = 12, B * K, 9
    px = torch.from_numpy(synthetic.make_batch(cfg, B, 16, seed=2)["pixel_values"]).cuda()
    enc = eng.encode(px, trunc_int=True, save=False, tag="gen.enc")
    enc_kv = eng.cross_kv(enc, tag="gen.enc")
    g = torch.Generator(device="cpu").manual_seed(11)
    tokens = torch.randint(4, cfg.mbart_config.vocab_size, (steps, R), generator=g).to(I32).cuda()
    # ancestor tables: history positions of a row point at arbitrary rows of the SAME image; the current position
    # always points at the row itself (beam.cu invariant)
    anc_tables = []
    for pos in range(steps):
        a = torch.arange(R)[:, None].expand(R, T).clone()
        if pos > 0:
            img = (torch.arange(R) // K)[:, None]
            a[:, :pos] = img * K + torch.randint(0, K, (R, pos), generator=g)
        anc_tables.append(a.to(I32).cuda())
    ref, ref_kv = _run(eng, False, enc_kv, R, T, K, tokens, anc_tables)
    out, out_kv = _run(eng, True, enc_kv, R, T, K, tokens, anc_tables)
    assert torch.isfinite(out).all()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item()
    assert err <= 0.03 * scale, (err, scale)
    kv_err = (out_kv - ref_kv).abs().max().item()
    assert kv_err <= 0.03 * ref_kv.abs().max().item(), kv_err
    # and the accumulators were handed back zeroed, the barrier counter re-armed
    assert float(eng.bufs.t["gen.acc"].abs().max()) == 0.0 and float(eng.bufs.t["gen.q_acc"].abs().max()) == 0.0
    assert int(eng._fused_plans["decoder"][1]["sync"][0].item()) == 0


@pytest.mark.parametrize("num_beams", [1, 4])
def test_generate_same_tokens_with_and_without_fused_decoder(num_beams):
    cfg = mic_b200.tiny_config()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=3)
    model.params = synthetic.make_params(cfg, seed=5, perturbed=True, std=0.12)
    px = torch.from_numpy(synthetic.make_batch(cfg, 5, 16, seed=2)["pixel_values"]).cuda()
    kw = dict(max_length=10, pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, num_beams=num_beams, min_length=0,
              forced_bos_token_id=7, forced_eos_token_id=2, length_penalty=1.0, early_stopping=True)
    res = {}
    for fused in (False, True):
        model.engine.fused_decoder = fused
        model.engine.__dict__.pop("_gen_graphs", None)
        a = gen.generate(model.engine, px, use_cuda_graph=False, **kw)["sequences"].cpu().numpy()
        b = gen.generate(model.engine, px, use_cuda_graph=True, **kw)["sequences"].cpu().numpy()      # eager warm-up
        c = gen.generate(model.engine, px, use_cuda_graph=True, **kw)["sequences"].cpu().numpy()      # graph replay
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
        res[fused] = a
    same = (res[False] == res[True]).all(axis=1).mean()
    assert same >= 0.6, (same, res)        # 5 rows; bf16 + summation-order noise may fork one or two late
    # (the margin-aware, oracle-anchored comparison of the fused path is tests/test_full_size_golden_gpu.py)


def test_long_generation_takes_the_per_op_path_and_matches_the_oracle_prefix():
    """max_length = 150 (> 64 keys: not eligible for the persistent kernel, > 128: beyond the old attention limit).
    Greedy; the first tokens must agree with the oracle wherever its top-2 logit margin exceeds the bf16 noise."""
    from oracle import reference_generate as rg
    cfg = mic_b200.tiny_config()
    cfg.mbart_config.max_position_embeddings = 256
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=3)
    params = synthetic.make_params(cfg, seed=5, perturbed=True, std=0.12)
    model.params = params
    batch = synthetic.make_batch(cfg, 3, seq_len=16, seed=2)
    kw = dict(max_length=150, pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, num_beams=1, min_length=0,
              forced_bos_token_id=7, forced_eos_token_id=2)
    out = model.generate(batch["pixel_values"], **kw).sequences
    out = out.cpu().numpy() if hasattr(out, "cpu") else np.asarray(out)
    assert out.shape == (3, 150)
    assert "decoder" not in model.engine.__dict__.get("_fused_plans", {}), "max_length 150 must not build a fused plan"
    ref = rg.generate(params, batch["pixel_values"], cfg, return_trace=True, length_penalty=1.0, early_stopping=True, **kw)
    ref_seq = np.asarray(ref["sequences"])
    assert (out[:, :2] == ref_seq[:, :2]).all()
    agree = (out == ref_seq).mean()
    assert agree > 0.5, agree          # bf16 drift may fork a row late; the bulk of 150 positions still agrees


@pytest.mark.parametrize("fused", [True, False])
def test_graph_replay_survives_other_shapes_and_new_params(fused):
    """A captured generate() graph holds raw pointers: running another search shape in between (which re-requests
    the decode buffers with other shapes) and changing the parameters must not invalidate its replay."""
    cfg = mic_b200.tiny_config()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=3)
    model.engine.fused_decoder = fused
    px = torch.from_numpy(synthetic.make_batch(cfg, 5, 16, seed=2)["pixel_values"]).cuda()
    kw_a = dict(max_length=12, pad_token_id=1, eos_token_id=2, decoder_start_token_id=2, num_beams=1, min_length=0,
                forced_bos_token_id=7, forced_eos_token_id=2, length_penalty=1.0, early_stopping=True)
    kw_b = dict(kw_a, max_length=2)
    kw_c = dict(kw_a, num_beams=3, max_length=7)
    p1 = synthetic.make_params(cfg, seed=5, perturbed=True, std=0.12)
    p2 = synthetic.make_params(cfg, seed=6, perturbed=True, std=0.12)
    p2["final_logits_bias"] = p2["final_logits_bias"].copy()
    p2["final_logits_bias"][0, 2] += 5.0
    model.params = p1
    gen.generate(model.engine, px, **kw_a)                       # eager warm-up + capture of graph A
    gen.generate(model.engine, px, **kw_b)                       # other cache length
    gen.generate(model.engine, px, **kw_c)                       # other row count
    model.params = p2
    replay = gen.generate(model.engine, px, **kw_a)["sequences"].cpu().numpy()
    eager = gen.generate(model.engine, px, use_cuda_graph=False, **kw_a)["sequences"].cpu().numpy()
    np.testing.assert_array_equal(replay, eager)
