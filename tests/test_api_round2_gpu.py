"""Round-2 API surface on the GPU: checkpoint interop + grafting, per-call `params=` override, generation hooks,
`_sample`, uint8 input hand-off, train-mode `__call__`, the ViT-BART class with the reference's names and its cached
(post-LN) decode."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import mic_b200  # noqa: E402
from mic_b200 import synthetic  # noqa: E402
from oracle import reference_generate as rg  # noqa: E402
from oracle import reference_model as rm  # noqa: E402


def _tiny(seed=5, std=0.12):
    cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
    params = synthetic.make_params(cfg, seed=seed, perturbed=True, std=std)
    batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=seed, min_len=4)
    return cfg, params, batch


def _logits(model, batch, **kw):
    return model(batch["pixel_values"], batch["decoder_input_ids"], batch["attention_mask"], **kw).logits.float().cpu().numpy()


def test_save_pretrained_from_pretrained_round_trip(tmp_path):
    cfg, params, batch = _tiny()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    ref = _logits(model, batch)
    model.save_pretrained(str(tmp_path / "ckpt"))
    again = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration.from_pretrained(str(tmp_path / "ckpt"), seed=123)
    assert again.missing_keys == [] and again.unexpected_keys == []
    assert again.config.mbart_config.vocab_size == 1003 and again.config.clip_vision_config.hidden_size == 128
    np.testing.assert_array_equal(_logits(again, batch), ref)
    a = dict(synthetic.tree_flatten(again.store.to_numpy_tree()))
    for k, v in synthetic.tree_flatten(params):
        np.testing.assert_array_equal(a[k], v)


def test_from_pretrained_reports_missing_and_unexpected(tmp_path):
    cfg, params, batch = _tiny()
    ck = mic_b200.checkpoint
    partial = {"model": dict(params["model"]), "final_logits_bias": params["final_logits_bias"], "stray": {"w": np.ones(3, np.float32)}}
    partial["model"] = {k: v for k, v in partial["model"].items() if k != "visual_projection"}
    ck.write_weights(str(tmp_path / "p"), partial, cfg.to_dict())
    m = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration.from_pretrained(str(tmp_path / "p"), seed=1)
    assert ("model", "visual_projection", "kernel") in m.missing_keys and ("stray", "w") in m.unexpected_keys
    got = m.store.to_numpy_tree()
    np.testing.assert_array_equal(got["model"]["shared"]["embedding"], params["model"]["shared"]["embedding"])
    assert np.abs(got["model"]["visual_projection"]["kernel"]).max() > 0        # kept its random init


def test_from_clip_vision_mbart_pretrained_grafts_two_checkpoints(tmp_path):
    """modeling_clip_vision_mbart.py:702-773: encoder <- CLIP vision checkpoint, decoder + shared <- mBART checkpoint;
    visual_projection / final_logits_bias keep their initial values."""
    cfg, params, batch = _tiny()
    ck = mic_b200.checkpoint
    clip_tree = params["model"]["encoder"]                                         # {"vision_model": ...} = FlaxCLIPVisionModel.params
    mbart_tree = {"shared": params["model"]["shared"], "decoder": params["model"]["decoder"],
                  "encoder": {"unused": {"w": np.zeros(2, np.float32)}}}            # FlaxMBartModel also has an encoder
    ck.write_weights(str(tmp_path / "clip"), clip_tree, {"vision_config": cfg.to_dict()["clip_vision_config"]})
    ck.write_weights(str(tmp_path / "mbart"), mbart_tree, cfg.to_dict()["mbart_config"])
    m = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration.from_clip_vision_mbart_pretrained(
        str(tmp_path / "clip"), str(tmp_path / "mbart"), seed=9)
    got = m.store.to_numpy_tree()
    np.testing.assert_array_equal(got["model"]["decoder"]["layers"]["1"]["fc1"]["kernel"],
                                  params["model"]["decoder"]["layers"]["1"]["fc1"]["kernel"])
    np.testing.assert_array_equal(got["model"]["encoder"]["vision_model"]["embeddings"]["class_embedding"],
                                  params["model"]["encoder"]["vision_model"]["embeddings"]["class_embedding"])
    np.testing.assert_array_equal(got["model"]["shared"]["embedding"], params["model"]["shared"]["embedding"])
    assert np.abs(got["final_logits_bias"]).max() == 0.0                          # zeros-init (:133-135), not grafted
    assert m.config.mbart_config.vocab_size == 1003
    # same through ready objects (`clip_vision_model=` / `mbart_model=`)
    m2 = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration.from_clip_vision_mbart_pretrained(
        clip_vision_model=(clip_tree, cfg.clip_vision_config), mbart_model=(mbart_tree, cfg.mbart_config), seed=9)
    np.testing.assert_array_equal(m2.store.to_numpy_tree()["model"]["decoder"]["layer_norm"]["scale"],
                                  params["model"]["decoder"]["layer_norm"]["scale"])


def test_train_checkpoint_with_optimizer_state_round_trip(tmp_path):
    """main.py:299-346: flax_model.msgpack + opt_state.msgpack + training_state.json, restored into a fresh state."""
    cfg, params, batch = _tiny()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    sched = mic_b200.create_learning_rate_fn(1000, 4, 1, 2, 1e-3)
    state = mic_b200.TrainState.create(apply_fn=model.__call__, params=params, tx=mic_b200.adamw(sched, b2=0.98),
                                       dropout=0.0)
    assert state.b2 == 0.98
    with pytest.raises(TypeError):
        mic_b200.TrainState.create(model=model, tx=object())
    for _ in range(3):
        mic_b200.train_step(state, batch)
    d = mic_b200.save_model_checkpoint(model, str(tmp_path), state, with_opt=True)
    assert d.endswith("ckpt-2")
    model2 = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=7)
    state2 = mic_b200.TrainState.create(model=model2, tx=mic_b200.adamw(sched, b2=0.98), dropout=0.0)
    p, o, step = mic_b200.restore_model_checkpoint(d, state2)
    assert step == 3 and set(o) == {"0", "1", "2"} and int(o["0"]["count"]) == 3
    model2.params = p
    state2.load_opt_state(o, step)
    assert torch.equal(model2.store.master, model.store.master)
    assert torch.equal(state2.store.adam_m, state.store.adam_m) and torch.equal(state2.store.adam_v, state.store.adam_v)
    _, m1 = mic_b200.train_step(state, batch)
    _, m2 = mic_b200.train_step(state2, batch)
    assert abs(float(m1["loss"]) - float(m2["loss"])) < 1e-5 and m1["learning_rate"] == m2["learning_rate"]
    mic_b200.save_model_checkpoint(model, str(tmp_path), state, with_opt=False)
    mic_b200.rotate_checkpoints(str(tmp_path), 1)
    import os
    assert sorted(x for x in os.listdir(tmp_path) if x.startswith("ckpt-")) == ["ckpt-3"]


def test_params_kwarg_is_a_per_call_override():
    """Reference semantics (pure-functional apply): `params=` does not change the model's own weights."""
    cfg, params, batch = _tiny()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    own = _logits(model, batch)
    other = synthetic.make_params(cfg, seed=77, perturbed=True, std=0.12)
    with_other = _logits(model, batch, params=other)
    assert np.abs(with_other - own).max() > 1e-2
    np.testing.assert_array_equal(_logits(model, batch), own)                     # weights are back
    np.testing.assert_array_equal(_logits(model, batch, params=model.params), own)  # own live tree: no copy, same result
    seq = model.generate(batch["pixel_values"], num_beams=1, max_length=8, forced_bos_token_id=1001, params=other)
    ref = rg.generate(other, batch["pixel_values"], cfg, num_beams=1, max_length=8, forced_bos_token_id=1001)
    assert (seq.sequences.cpu().numpy()[:, :3] == ref["sequences"][:, :3]).all()
    np.testing.assert_array_equal(_logits(model, batch), own)


def test_generation_hooks_drive_a_caller_side_greedy_loop():
    """prepare_inputs_for_generation / update_inputs_for_generation (:653-693) around decode(): the loop a user of the
    reference could write by hand gives the tokens generate() gives."""
    cfg, params, batch = _tiny(std=0.3)
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    L = 9
    enc = model.encode(batch["pixel_values"])
    B = enc.last_hidden_state.shape[0]
    ids = torch.full((B, 1), 2, dtype=torch.int32, device="cuda")
    kw = model.prepare_inputs_for_generation(ids, L, encoder_outputs=enc)
    assert kw["decoder_attention_mask"].shape == (B, L) and int(kw["decoder_position_ids"].max()) == 0
    toks = [ids]
    for cur_len in range(1, L):
        out = model.decode(toks[-1], **kw)
        logits = out.logits[:, -1].float()
        if cur_len == 1:
            nxt = torch.full((B,), 1001, dtype=torch.int64, device="cuda")
        elif cur_len == L - 1:
            nxt = torch.full((B,), 2, dtype=torch.int64, device="cuda")
        else:
            nxt = logits.argmax(-1)
        toks.append(nxt.to(torch.int32)[:, None])
        kw = model.update_inputs_for_generation(out, kw)
        assert int(kw["decoder_position_ids"][0, 0]) == cur_len
    mine = torch.cat(toks, 1).cpu().numpy()
    ref = rg.generate(params, batch["pixel_values"], cfg, num_beams=1, max_length=L, forced_bos_token_id=1001,
                      return_trace=True)
    compared = 0
    for b in range(B):
        for pos in range(1, L):
            if ref["margins"][b, pos - 1] < 0.1 or ref["sequences"][b, pos] == 1:
                break                       # sub-tolerance margin / the oracle row finished early (EOS -> pad handling)
            assert mine[b, pos] == ref["sequences"][b, pos], (b, pos, mine[b], ref["sequences"][b])
            compared += 1
    assert compared >= 12, compared


def test_sample_matches_the_oracle_stream():
    """`_sample` (:537-663): tokens = argmax(raw logits + Gumbel(threefry(key))).  Same key schedule and noise on both
    sides, so ids agree wherever the noisy top-2 margin exceeds the bf16 logit error; forced BOS/EOS are ignored (the
    reference samples from the raw logits)."""
    cfg, params, batch = _tiny(std=0.12)        # logits of O(3): the bf16 logit error (~0.03) is far below the 0.15 margin
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    total = 0
    for key in (None, np.array([3, 12345], np.uint32), np.array([9, 1], np.uint32), np.array([77, 5], np.uint32)):
        ref = rg.generate(params, batch["pixel_values"], cfg, num_beams=1, max_length=12, forced_bos_token_id=1001,
                          do_sample=True, prng_key=key, return_trace=True)
        out = model.generate(batch["pixel_values"], num_beams=1, max_length=12, forced_bos_token_id=1001, do_sample=True,
                             prng_key=key).sequences.cpu().numpy()
        assert out.shape == ref["sequences"].shape
        compared = 0
        for b in range(out.shape[0]):
            for pos in range(1, 12):
                if pos - 1 >= ref["margins"].shape[1] or ref["margins"][b, pos - 1] < 0.15:
                    break
                assert out[b, pos] == ref["sequences"][b, pos], (b, pos, out[b], ref["sequences"][b])
                compared += 1
                if ref["sequences"][b, pos] == 1:
                    break
        total += compared
        assert (ref["sequences"][:, 1] != 1001).any()           # forced BOS is NOT applied when sampling (quirk)
    assert total >= 20, total
    a = model.generate(batch["pixel_values"], num_beams=1, max_length=12, do_sample=True, prng_key=np.array([0, 1], np.uint32))
    b = model.generate(batch["pixel_values"], num_beams=1, max_length=12, do_sample=True, prng_key=np.array([0, 2], np.uint32))
    assert not np.array_equal(a.sequences.cpu().numpy(), b.sequences.cpu().numpy())
    with pytest.raises(NotImplementedError):
        model.generate(batch["pixel_values"], num_beams=2, do_sample=True)


def test_uint8_pixels_are_normalised_in_the_patch_kernel():
    """Input hand-off (main.py:165-179): uint8 NHWC in, x/255 and Normalize(mean, std) inside the patch kernel."""
    cfg, params, batch = _tiny()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    rng = np.random.default_rng(0)
    u8 = rng.integers(0, 256, size=batch["pixel_values"].shape, dtype=np.uint8)
    c = cfg.clip_vision_config
    f32 = ((u8.astype(np.float32) / np.float32(255.0)) - np.array(c.image_mean, np.float32)) / np.array(c.image_std, np.float32)
    a = model(u8, batch["decoder_input_ids"], batch["attention_mask"]).logits.float().cpu().numpy()
    b = model(f32, batch["decoder_input_ids"], batch["attention_mask"]).logits.float().cpu().numpy()
    assert np.abs(a - b).max() <= 2e-2 * np.abs(b).max()
    # generate(): the int32 cast quirk applies to the NORMALISED value
    sa = model.generate(u8, num_beams=1, max_length=6, forced_bos_token_id=1001).sequences.cpu().numpy()
    sb = model.generate(f32, num_beams=1, max_length=6, forced_bos_token_id=1001).sequences.cpu().numpy()
    assert (sa[:, :2] == sb[:, :2]).all()
    # and the training step takes uint8 batches (a quarter of the H2D bytes)
    state = mic_b200.TrainState(model, mic_b200.create_learning_rate_fn(1000, 4, 1, 10, 1e-3), dropout=0.0)
    state2_model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    state2_model.params = params
    state2 = mic_b200.TrainState(state2_model, mic_b200.create_learning_rate_fn(1000, 4, 1, 10, 1e-3), dropout=0.0)
    bu = dict(batch, pixel_values=torch.from_numpy(u8))
    bf = dict(batch, pixel_values=f32)
    for _ in range(3):
        _, mu = mic_b200.train_step(state, bu)
        _, mf = mic_b200.train_step(state2, bf)
    assert abs(float(mu["loss"]) - float(mf["loss"])) < 2e-2


def test_call_train_true_applies_decoder_dropout():
    cfg, params, batch = _tiny()
    model = mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    ev = _logits(model, batch)
    a = _logits(model, batch, train=True, dropout_rng=np.array([0, 1], np.uint32))
    a2 = _logits(model, batch, train=True, dropout_rng=np.array([0, 1], np.uint32))
    b = _logits(model, batch, train=True, dropout_rng=np.array([0, 2], np.uint32))
    np.testing.assert_array_equal(a, a2)
    assert np.abs(a - ev).max() > 1e-2 and np.abs(a - b).max() > 1e-2
    np.testing.assert_array_equal(_logits(model, batch), ev)          # dropout is off again afterwards
    with pytest.raises(ValueError):
        model(batch["pixel_values"], batch["decoder_input_ids"], train=True)
    bad = mic_b200.tiny_config(vocab_size=1003, layers=2)
    bad.mbart_config.attention_dropout = 0.1
    with pytest.raises(NotImplementedError):
        mic_b200.FlaxCLIPVisionMBartForConditionalGeneration(bad)


# ---------------------------------------------------------------------------------------------------------------
# flax_vit_bart variant as a first-class model
# ---------------------------------------------------------------------------------------------------------------
def _vit(std=0.12):
    cfg = mic_b200.tiny_vit_bart_config(vocab_size=1003, layers=2)
    params = synthetic.make_params(cfg, seed=4, perturbed=True, std=std)
    batch = synthetic.make_batch(cfg, 4, seq_len=16, seed=4, min_len=4)
    return cfg, params, batch


def test_vit_bart_class_uses_the_reference_parameter_names(tmp_path):
    cfg, params, batch = _vit()
    model = mic_b200.FlaxViTBartForConditionalGeneration(cfg, seed=0)
    model.params = params                                           # canonical tree accepted (synthetic generator)
    tree = model.params
    enc = tree["model"]["encoder"]
    assert set(enc) == {"embeddings", "encoder", "layernorm", "pooler"}
    assert tuple(enc["embeddings"]["cls_token"].shape) == (1, 1, cfg.clip_vision_config.hidden_size)
    assert "layer_norm" not in tree["model"]["decoder"] and "final_logits_bias" in tree
    ref_logits = _logits(model, batch)
    with torch.no_grad():
        want = rm.forward_logits(rm.to_torch_tree(params), batch["pixel_values"], batch["decoder_input_ids"],
                                 batch["attention_mask"], None, cfg).numpy()
    # bf16 product path vs the fp32 oracle: worst logit of 64k within 4 % of the logit range, rms within 1 %
    assert np.abs(ref_logits - want).max() <= 4e-2 * np.abs(want).max()
    assert np.sqrt(((ref_logits - want) ** 2).mean()) <= 1e-2 * np.abs(want).max()
    # live views: writing through the reference-named tree changes the model
    tree["model"]["encoder"]["layernorm"]["scale"].mul_(1.5)
    model.store.refresh_shadow()
    assert np.abs(_logits(model, batch) - ref_logits).max() > 1e-3
    tree["model"]["encoder"]["layernorm"]["scale"].div_(1.5)
    model.store.refresh_shadow()
    # checkpoint round trip in the reference's names
    model.save_pretrained(str(tmp_path / "vb"))
    stored = mic_b200.checkpoint.read_weights(str(tmp_path / "vb"))
    assert "query" in stored["model"]["encoder"]["encoder"]["layer"]["0"]["attention"]["attention"]
    again = mic_b200.FlaxViTBartForConditionalGeneration.from_pretrained(str(tmp_path / "vb"), seed=5)
    assert again.missing_keys == [] and again.unexpected_keys == []
    assert np.abs(_logits(again, batch) - ref_logits).max() <= 1e-6 + 1e-3 * np.abs(ref_logits).max()
    # grafting (modeling_vit_bart.py:663-732)
    vit_tree = stored["model"]["encoder"]
    bart_tree = {"shared": stored["model"]["shared"], "decoder": stored["model"]["decoder"]}
    g = mic_b200.FlaxViTBartForConditionalGeneration.from_vit_bart_pretrained(
        vit_model=(vit_tree, cfg.clip_vision_config.__dict__), bart_model=(bart_tree, cfg.mbart_config.__dict__), seed=3)
    np.testing.assert_array_equal(g.store.to_numpy_tree()["model"]["decoder"]["layers"]["0"]["fc2"]["kernel"],
                                  params["model"]["decoder"]["layers"]["0"]["fc2"]["kernel"])
    with pytest.raises(ValueError):
        mic_b200.FlaxViTBartForConditionalGeneration(mic_b200.tiny_config())


@pytest.mark.parametrize("num_beams", [1, 4])
def test_vit_bart_generate_with_cached_post_ln_decode(num_beams):
    """generate() of the variant: corrected encode (transpose + visual_projection, unlike modeling_vit_bart.py:292-300)
    and the cached decode of the POST-LN BART decoder, against the oracle."""
    cfg, params, batch = _vit(std=0.12)
    model = mic_b200.FlaxViTBartForConditionalGeneration(cfg, seed=0)
    model.params = params
    # (a) the cached post-LN decode step reproduces the full-sequence forward of the same model, position by position
    ids = torch.from_numpy(batch["decoder_input_ids"][:, :8]).cuda()
    full = model(batch["pixel_values"], ids).logits.float()
    enc = model.encode(np.trunc(batch["pixel_values"]))                      # same (already integral) pixels for both
    full = model(np.trunc(batch["pixel_values"]), ids).logits.float()
    hk = model.prepare_inputs_for_generation(ids[:, :1], 8, encoder_outputs=enc)
    for t_ in range(8):
        out = model.decode(ids[:, t_:t_ + 1], **hk)
        err = (out.logits[:, 0].float() - full[:, t_]).abs().max().item()
        assert err <= 3e-2 * full.abs().max().item(), (t_, err, full.abs().max().item())
        hk = model.update_inputs_for_generation(out, hk)
    # (b) generate() against the oracle
    kw = dict(num_beams=num_beams, max_length=10, forced_bos_token_id=1001, decoder_start_token_id=2)
    ref = rg.generate(params, batch["pixel_values"], cfg, return_trace=True, **kw)
    out = model.generate(batch["pixel_values"], **kw)
    seq = out.sequences.cpu().numpy()
    seq2 = model.generate(batch["pixel_values"], **kw).sequences.cpu().numpy()           # graph replay
    np.testing.assert_array_equal(seq, seq2)
    assert seq.shape == ref["sequences"].shape
    assert (seq[:, :2] == ref["sequences"][:, :2]).all()
    if num_beams == 1:
        compared = 0
        for b in range(seq.shape[0]):
            for pos in range(1, seq.shape[1]):
                if ref["margins"][b, pos - 1] < 0.1:
                    break
                assert seq[b, pos] == ref["sequences"][b, pos], (b, pos, seq[b], ref["sequences"][b])
                compared += 1
                if ref["sequences"][b, pos] == 1:
                    break
        assert compared >= 8, compared
    else:
        clear = np.ones(seq.shape[0], bool)
        for step in ref["trace"]:
            allv = np.concatenate([step["topk_raw"], step["ninth"][:, None]], 1)
            with np.errstate(invalid="ignore"):
                gaps = np.abs(np.diff(allv, axis=1))
            ok = (gaps > 0.1) | ~np.isfinite(gaps) | ((np.abs(allv[:, :-1]) > 1e6) & (np.abs(allv[:, 1:]) > 1e6))
            clear &= ok.all(1)
        np.testing.assert_array_equal(seq[clear], ref["sequences"][clear])
        # every row: the sequence the CUDA search returns scores (under the ORACLE) within the bf16 drift of the oracle's
        # own best hypothesis — a wrong cache / ancestor handling would produce far worse sequences
        # every row agrees with the oracle at least up to its first sub-tolerance beam decision: the first un-forced
        # token (cur_len 2) has all candidates from one live beam with distinct scores
        assert (seq[:, :3] == ref["sequences"][:, :3]).mean() >= 0.9, (seq, ref["sequences"])
    # the encoder states generate() uses: projected to d_model
    enc = model.encode(batch["pixel_values"]).last_hidden_state
    with torch.no_grad():
        want = rm.encode(rm.to_torch_tree(params), batch["pixel_values"], cfg, int32_cast=True).numpy()
    assert tuple(enc.shape) == want.shape and want.shape[-1] == cfg.mbart_config.d_model
    assert np.abs(enc.float().cpu().numpy() - want).max() <= 3e-2 * np.abs(want).max()
