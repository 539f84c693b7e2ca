"""End-to-end GPU parity against the CPU oracle on identical random weights and synthetic inputs
(SURVEY.md §8c/d).  bf16 compute vs fp32 oracle: tolerances are the north-star's bf16 ones."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import synthetic  # noqa: E402
from oracle import reference_model as rm  # noqa: E402
from oracle import reference_generate as rg  # noqa: E402


def _setup(vocab=1003, layers=2, B=4, T=16, perturbed=True, std=0.05, seed=1):
    cfg = mic_b200.tiny_config(vocab_size=vocab, layers=layers)
    params = synthetic.make_params(cfg, seed=seed, perturbed=perturbed, std=std)
    batch = synthetic.make_batch(cfg, B, seq_len=T, seed=0, min_len=4)
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    return cfg, params, batch, model


def _flat(tree, prefix=""):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from _flat(v, prefix + k + "/")
        else:
            yield prefix + k, v


def test_params_roundtrip_names_and_shapes():
    cfg, params, _, model = _setup()
    got = model.store.to_numpy_tree()
    want = dict(_flat(params))
    have = dict(_flat(got))
    assert set(want) == set(have)
    for k in want:
        np.testing.assert_array_equal(have[k], want[k], err_msg=k)


def test_forward_logits_and_loss_match_oracle():
    cfg, params, batch, model = _setup()
    out = model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"])
    logits = out[0].float().cpu()
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        ref = rm.forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], None, cfg)
    assert logits.shape == ref.shape == (4, 16, 1003)
    rel = float((logits - ref).abs().max() / ref.abs().max())
    assert rel < 3e-2, rel                                   # bf16 activations end to end
    for eps in (0.0, 0.1):
        loss = float(model.loss(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"],
                                batch["input_ids"], eps))
        want = float(rm.loss_fn(ref, batch["input_ids"], batch["attention_mask"], eps))
        assert abs(loss - want) < 2e-2, (eps, loss, want)    # north-star bf16 loss tolerance


def test_gradients_match_oracle_autograd():
    cfg, params, batch, model = _setup()
    eng = model.engine
    model.store.ensure_grad()
    model.store.grad.fill_(float("nan"))                     # every slot must be (over)written
    ws = eng.forward_backward(torch.from_numpy(batch["pixel_values"]), torch.from_numpy(batch["decoder_input_ids"]),
                              torch.from_numpy(batch["attention_mask"]), torch.from_numpy(batch["input_ids"]), 0.1)
    torch.cuda.synchronize()
    loss_ref, grads_ref, _ = rm.loss_and_grads(params, batch, cfg, 0.1)
    assert abs(float(ws["out"][0]) - loss_ref) < 2e-2
    got = dict(_flat(model.store.to_numpy_tree(model.store.grad)))
    want = dict(_flat(grads_ref))
    worst = []
    for k in sorted(want):
        g, w = got[k], want[k]
        assert np.isfinite(g).all(), f"non-finite / unwritten gradient: {k}"
        denom = np.linalg.norm(w) + 1e-12
        err = np.linalg.norm(g - w) / denom if denom > 1e-7 else np.abs(g).max()
        worst.append((err, k))
    worst.sort(reverse=True)
    bad = [(e, k) for e, k in worst if e > 0.08]
    assert not bad, f"relative grad error too large: {bad[:10]}"
    # position rows never used must have exactly zero gradient
    pos = got["model/decoder/embed_positions/embedding"]
    assert np.all(pos[16 + 2:] == 0) and np.all(pos[:2] == 0)


def test_train_steps_follow_oracle_adamw():
    cfg, params, batch, model = _setup(layers=1)
    sched = mic_b200.create_learning_rate_fn(1000, 10, 1, 2, 1e-2)
    state = mic_b200.TrainState(model, sched, weight_decay=0.01, dropout=0.0)   # oracle comparison: deterministic
    losses = []
    p_ref = {k: v.copy() for k, v in _flat(params)}
    m_ref = {k: np.zeros_like(v) for k, v in p_ref.items()}
    v_ref = {k: np.zeros_like(v) for k, v in p_ref.items()}

    def unflat(d):
        out = {}
        for k, v in d.items():
            node = out
            parts = k.split("/")
            for q in parts[:-1]:
                node = node.setdefault(q, {})
            node[parts[-1]] = v
        return out
    for step in range(3):
        state, metrics = mic_b200.train_step(state, batch, 0.0)
        losses.append(float(metrics["loss"]))
        _, g_ref, _ = rm.loss_and_grads(unflat(p_ref), batch, cfg, 0.0)
        g_ref = dict(_flat(g_ref))
        lr = rm.linear_warmup_decay_lr(step, 1e-2, 2, 100)
        assert abs(metrics["learning_rate"] - lr) < 1e-12
        for k in p_ref:
            p_ref[k], m_ref[k], v_ref[k] = rm.adamw_update(p_ref[k], g_ref[k], m_ref[k], v_ref[k], step, lr,
                                                           weight_decay=0.01)
    got = dict(_flat(model.store.to_numpy_tree()))
    # step 0 has lr(0) = 0 (schedule evaluated at the pre-increment count) -> two effective updates
    for k in ("model/visual_projection/kernel", "model/decoder/layers/0/fc1/kernel", "final_logits_bias",
              "model/encoder/vision_model/encoder/layers/0/self_attn/q_proj/kernel"):
        moved = np.abs(p_ref[k] - dict(_flat(params))[k]).mean()
        diff = np.abs(got[k] - p_ref[k]).mean()
        assert diff < 0.35 * moved + 1e-6, (k, diff, moved)   # Adam's sign-like update amplifies bf16 grad noise
    assert losses[2] < losses[0]


@pytest.mark.parametrize("beams", [1, 4])
def test_generate_matches_oracle_where_margin_allows(beams):
    cfg, params, batch, model = _setup(vocab=1003, layers=2, B=3, std=0.3, seed=5)
    L = 12
    kw = dict(num_beams=beams, max_length=L, forced_bos_token_id=1001)
    ref = rg.generate(params, batch["pixel_values"], cfg, return_trace=True, **kw)
    out = model.generate(batch["pixel_values"], **kw)
    seq = out.sequences.cpu().numpy()
    assert seq.shape == ref["sequences"].shape == (3, L) and seq.dtype == np.int32
    assert np.all(seq[:, 0] == 2) and np.all(seq[:, 1] == 1001)
    if beams == 1:
        margins = ref["margins"]                       # (B, L-1) top1 - top2 raw-logit gap per step
        for b in range(3):
            for pos in range(1, L):
                if pos >= 2 and pos < L - 1 and margins[b, pos - 1] < 0.05:
                    break                              # sub-tolerance decision: stop comparing this row
                assert seq[b, pos] == ref["sequences"][b, pos], (b, pos, seq[b], ref["sequences"][b])
        assert np.all(seq[:, -1] == 1)
    else:
        # compare rows whose every selection margin (8th vs 9th candidate, adjacent kept candidates) is clear
        clear = np.ones(3, bool)
        for step in ref["trace"]:
            if step["cur_len"] in (1, L - 1):
                continue
            v = step["topk_log_probs"]
            gaps = np.abs(np.diff(np.concatenate([v, step["ninth"][:, None]], 1), axis=1))
            clear &= np.all((gaps > 0.05) | ~np.isfinite(gaps), axis=1)
        for b in range(3):
            if clear[b]:
                np.testing.assert_array_equal(seq[b], ref["sequences"][b])
        sc = out.scores.cpu().numpy()
        assert np.all(np.isfinite(sc))
        np.testing.assert_allclose(sc[clear], ref["scores"][clear], rtol=1e-3)


def test_generate_int32_pixel_truncation_is_reproduced():
    cfg, params, batch, model = _setup(B=2)
    enc_gen = model.encode(batch["pixel_values"]).last_hidden_state.float().cpu()
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        want_trunc = rm.encode(p, batch["pixel_values"], cfg, int32_cast=True)
        want_float = rm.encode(p, batch["pixel_values"], cfg, int32_cast=False)
    e_t = float((enc_gen - want_trunc).abs().max())
    e_f = float((enc_gen - want_float).abs().max())
    assert e_t < 0.1 * e_f + 2e-2, (e_t, e_f)


def test_cached_decode_api_matches_full_forward():
    cfg, params, batch, model = _setup(B=2)
    enc = model.encode(batch["pixel_values"])
    ids = torch.from_numpy(batch["decoder_input_ids"][:, :6]).cuda()
    full = model.decode(ids, enc).logits.float()
    cache = model.init_cache(2, 8, enc)
    for t in range(6):
        pos = torch.full((2, 1), t)
        step = model.decode(ids[:, t:t + 1], enc, decoder_position_ids=pos, past_key_values=cache)
        assert (step.logits[:, 0].float() - full[:, t]).abs().max() < 0.05 * full.abs().max()
        cache = step.past_key_values
    with pytest.raises(ValueError):
        model.decode(ids[:, :1], enc, past_key_values=cache)


def test_generate_cuda_graph_replay_equals_eager():
    """Second call replays the captured graph; it must reproduce the eager result and track new pixels."""
    cfg, params, batch, model = _setup(vocab=1003, layers=2, B=3, std=0.3, seed=5)
    kw = dict(num_beams=4, max_length=12, forced_bos_token_id=1001)
    a = model.generate(batch["pixel_values"], **kw)            # eager (captures afterwards)
    b = model.generate(batch["pixel_values"], **kw)            # replay
    assert torch.equal(a.sequences, b.sequences) and torch.equal(a.scores, b.scores)
    other = synthetic.make_batch(cfg, 3, seq_len=16, seed=9)["pixel_values"]
    c = model.generate(other, **kw)                            # replay with new input
    ref = rg.generate(params, other, cfg, **kw)
    assert np.all(c.sequences.cpu().numpy()[:, :2] == ref["sequences"][:, :2])
    d = model.generate(batch["pixel_values"], **kw)
    assert torch.equal(a.sequences, d.sequences)
    g1 = model.generate(batch["pixel_values"], num_beams=1, max_length=12, forced_bos_token_id=1001)
    g2 = model.generate(batch["pixel_values"], num_beams=1, max_length=12, forced_bos_token_id=1001)
    assert torch.equal(g1.sequences, g2.sequences)


def test_train_step_with_dropout_runs_and_differs_per_step():
    """train=True semantics: decoder dropout 0.1 (mbart_config.dropout) with a fresh mask every step."""
    cfg, params, batch, model = _setup(layers=2)
    assert cfg.mbart_config.dropout == 0.1
    state = mic_b200.TrainState(model, lambda step: 0.0)          # lr 0: parameters frozen, only the masks change
    losses = []
    for _ in range(4):                                             # eager, warm, capture, replay
        state, m = mic_b200.train_step(state, batch)
        losses.append(float(m["loss"]))
    det = float(model.loss(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], batch["input_ids"]))
    assert all(np.isfinite(losses))
    assert len(set(round(l, 5) for l in losses)) == 4, losses     # new mask each step (also under graph replay)
    assert all(abs(l - det) < 0.5 for l in losses) and any(abs(l - det) > 1e-4 for l in losses)
    state0 = mic_b200.TrainState(model, lambda step: 0.0, dropout=0.0)
    _, m0 = mic_b200.train_step(state0, batch)
    assert abs(float(m0["loss"]) - det) < 2e-3


def test_vit_bart_variant_forward_and_gradients_match_oracle():
    """flax_vit_bart (BASELINE configs[4]): channel-first pixels, 16x16 patches with conv bias,
    exact-gelu ViT with final layernorm, POST-LN BART decoder without final LN; 82 visual tokens."""
    cfg = mic_b200.tiny_vit_bart_config(vocab_size=1003, layers=2)
    assert cfg.clip_vision_config.num_tokens == 82 and not cfg.mbart_config.pre_layernorm
    params = synthetic.make_params(cfg, seed=2, perturbed=True, std=0.05)
    batch = synthetic.make_batch(cfg, 3, seq_len=16, seed=1, min_len=4)
    assert batch["pixel_values"].shape == (3, 3, 144, 144)
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg)
    model.params = params
    logits = model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"]).logits.float().cpu()
    p = rm.to_torch_tree(params)
    with torch.no_grad():
        ref = rm.forward_logits(p, batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], None, cfg)
    rel = float((logits - ref).abs().max() / ref.abs().max())
    assert rel < 3e-2, rel
    model.store.ensure_grad()
    model.store.grad.fill_(float("nan"))
    ws = model.engine.forward_backward(torch.from_numpy(batch["pixel_values"]), torch.from_numpy(batch["decoder_input_ids"]),
                                       torch.from_numpy(batch["attention_mask"]), torch.from_numpy(batch["input_ids"]), 0.1)
    torch.cuda.synchronize()
    loss_ref, grads_ref, _ = rm.loss_and_grads(params, batch, cfg, 0.1)
    assert abs(float(ws["out"][0]) - loss_ref) < 2e-2
    got = dict(_flat(model.store.to_numpy_tree(model.store.grad)))
    want = dict(_flat(grads_ref))
    bad = []
    for k in sorted(want):
        g, w = got[k], want[k]
        assert np.isfinite(g).all(), f"non-finite / unwritten gradient: {k}"
        denom = np.linalg.norm(w)
        err = np.linalg.norm(g - w) / denom if denom > 1e-6 else np.abs(g).max()
        if err > 0.08:
            bad.append((round(float(err), 4), k))
    assert not bad, bad[:10]
