"""Checkpoint container (flax.serialization msgpack, checkpoint.py), the ViT-BART name mapping and the jax PRNG
restatement — everything that needs no GPU."""
import os

import msgpack
import numpy as np
import pytest

import mic_b200
from mic_b200 import synthetic
from oracle import reference_generate as rg

ck = mic_b200.checkpoint


def test_msgpack_container_round_trip_and_wire_format(tmp_path):
    tree = {"model": {"dense": {"kernel": np.arange(12, dtype=np.float32).reshape(3, 4), "bias": np.zeros(4, np.float32)}},
            "final_logits_bias": np.ones((1, 7), np.float32), "count": np.int32(5), "chain": [{"a": np.float32(2.5)}, {}]}
    data = ck.to_bytes(tree)
    # wire format: every array is ExtType(1, packb((shape, dtype name, raw bytes))) — what flax.serialization writes
    raw = msgpack.unpackb(data, raw=False, strict_map_key=False)
    ext = raw["model"]["dense"]["kernel"]
    assert isinstance(ext, msgpack.ExtType) and ext.code == 1
    shape, dtype, buf = msgpack.unpackb(ext.data, raw=False)
    assert list(shape) == [3, 4] and dtype == "float32" and buf == tree["model"]["dense"]["kernel"].tobytes()
    assert raw["count"].code == 3                                  # numpy scalar
    assert set(raw["chain"]) == {"0", "1"} and raw["chain"]["1"] == {}        # tuples / lists -> "0", "1", ...
    back = ck.msgpack_restore(data)
    np.testing.assert_array_equal(back["model"]["dense"]["kernel"], tree["model"]["dense"]["kernel"])
    assert back["count"] == 5 and float(back["chain"]["0"]["a"]) == 2.5
    # from_bytes restores INTO a target structure and refuses a key mismatch (flax semantics)
    got = ck.from_bytes({"model": {"dense": {"kernel": 0, "bias": 0}}, "final_logits_bias": 0, "count": 0,
                         "chain": [{"a": 0}, {}]}, data)
    np.testing.assert_array_equal(got["final_logits_bias"], tree["final_logits_bias"])
    with pytest.raises(ValueError):
        ck.from_bytes({"model": {"dense": {"kernel": 0}}}, data)


def test_large_arrays_are_chunked_like_flax(monkeypatch):
    monkeypatch.setattr(ck, "MAX_CHUNK_BYTES", 1024)
    a = np.arange(1000, dtype=np.float32).reshape(10, 100)          # 4000 bytes -> 4 chunks of <= 256 elements
    raw = msgpack.unpackb(ck.to_bytes({"big": a}), raw=False, strict_map_key=False)
    assert raw["big"]["__msgpack_chunked_array__"] is True and len(raw["big"]["chunks"]) == 4
    assert raw["big"]["shape"] == {"0": 10, "1": 100}
    np.testing.assert_array_equal(ck.msgpack_restore(ck.to_bytes({"big": a}))["big"], a)


def test_bfloat16_leaves_are_widened():
    f = np.array([1.0, -2.5, 3.140625], np.float32)
    u16 = (f.view(np.uint32) >> 16).astype(np.uint16)
    payload = msgpack.packb(([3], "bfloat16", u16.tobytes()), use_bin_type=True)
    data = msgpack.packb({"w": msgpack.ExtType(1, payload)}, use_bin_type=True)
    np.testing.assert_array_equal(ck.msgpack_restore(data)["w"], f)


def test_full_parameter_tree_round_trips_through_a_directory(tmp_path):
    cfg = mic_b200.tiny_config(vocab_size=1003, layers=2)
    params = synthetic.make_params(cfg, seed=3, perturbed=True)
    ck.write_weights(str(tmp_path / "m"), params, cfg.to_dict())
    assert sorted(os.listdir(tmp_path / "m")) == ["config.json", "flax_model.msgpack"]
    back = ck.read_weights(str(tmp_path / "m"))
    a, b = dict(synthetic.tree_flatten(params)), dict(synthetic.tree_flatten(back))
    assert a.keys() == b.keys()
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert ck.read_config_dict(str(tmp_path / "m"))["mbart_config"]["vocab_size"] == 1003
    with pytest.raises(OSError):
        ck.resolve_local_dir("openai/clip-vit-base-patch32")        # hub names cannot be resolved offline


def test_merge_reports_missing_and_unexpected_keys():
    want = {"a": {"w": np.zeros((2, 2), np.float32), "b": np.zeros(2, np.float32)}, "flb": np.zeros((1, 3), np.float32)}
    have = {"a": {"w": np.ones((2, 2), np.float32)}, "extra": {"z": np.ones(1, np.float32)}, "flb": np.ones((1, 3), np.float32)}
    merged, missing, unexpected = ck.merge_into(want, have)
    assert missing == [("a", "b")] and unexpected == [("extra", "z")]
    assert merged["a"]["w"].sum() == 4 and merged["a"]["b"].sum() == 0 and merged["flb"].sum() == 3
    with pytest.raises(ValueError):
        ck.merge_into(want, {"a": {"w": np.ones((3, 2), np.float32)}})


def test_vit_bart_names_are_a_bijection_of_the_canonical_tree():
    from mic_b200.modeling_vit_bart import canonical_to_vit_bart, vit_bart_to_canonical
    cfg = mic_b200.tiny_vit_bart_config(vocab_size=1003, layers=2)
    can = synthetic.make_params(cfg, seed=2, perturbed=True)
    dv = cfg.clip_vision_config.hidden_size
    pooler = {"kernel": np.ones((dv, dv), np.float32), "bias": np.zeros(dv, np.float32)}
    ref = canonical_to_vit_bart(can, pooler)
    enc = ref["model"]["encoder"]
    # the reference's FlaxViTModule / FlaxBartDecoder names and shapes (modeling_vit_bart.py:33-50)
    S = cfg.clip_vision_config.num_tokens
    assert enc["embeddings"]["cls_token"].shape == (1, 1, dv)
    assert enc["embeddings"]["position_embeddings"].shape == (1, S, dv)
    assert enc["embeddings"]["patch_embeddings"]["projection"]["kernel"].shape == (16, 16, 3, dv)
    assert set(enc["encoder"]["layer"]["0"]) == {"attention", "intermediate", "output", "layernorm_before", "layernorm_after"}
    assert set(enc["encoder"]["layer"]["0"]["attention"]["attention"]) == {"query", "key", "value"}
    assert "layernorm" in enc and "pooler" in enc and "vision_model" not in enc
    assert "layer_norm" not in ref["model"]["decoder"]               # BART has no final decoder LayerNorm
    back = vit_bart_to_canonical(ref, can)
    a, b = dict(synthetic.tree_flatten(can)), dict(synthetic.tree_flatten(back))
    assert a.keys() == b.keys()
    for k in a:
        np.testing.assert_array_equal(np.asarray(a[k]).reshape(-1), np.asarray(b[k]).reshape(-1))


def test_jax_prng_known_answers():
    """jax.random.split(jax.random.PRNGKey(0)) — the pair printed throughout the JAX documentation — pins the
    threefry2x32 restatement used by `_sample` (oracle and the host-side key schedule of the product)."""
    a, b = rg.prng_split(rg.prng_key(0))
    assert a.tolist() == [4146024105, 967050713] and b.tolist() == [2718843009, 1272950319]
    gen = mic_b200.generation
    assert gen.prng_split((0, 0)) == ((4146024105, 967050713), (2718843009, 1272950319))
    assert gen.prng_key_pair(None) == (0, 0) and gen.prng_key_pair(np.array([7, 9], np.uint32)) == (7, 9)
    # odd-sized count arrays are padded with one zero count
    odd = rg.threefry_2x32(rg.prng_key(1), np.arange(5, dtype=np.uint32))
    even = rg.threefry_2x32(rg.prng_key(1), np.array([0, 1, 2, 3, 4, 0], dtype=np.uint32))
    assert odd.tolist() == even[:5].tolist()
    # Gumbel-max sampling follows the softmax distribution
    logits = np.log(np.array([[0.7, 0.2, 0.1]], np.float32))
    key, hits = rg.prng_key(42), np.zeros(3)
    for _ in range(600):
        k, key = rg.prng_split(key)
        hits[rg.categorical(k, logits)[0]] += 1
    assert abs(hits[0] / 600 - 0.7) < 0.07 and abs(hits[2] / 600 - 0.1) < 0.05


def test_min_length_processor_formula():
    """FlaxMinLengthLogitsProcessor: apply_penalty = 1 - clip(cur_len - min_length, 0, 1): EOS is masked
    while cur_len <= min_length — checked against the literal formula, not through the oracle."""
    gen = mic_b200.generation
    for min_length in (0, 3, 6):
        for cur_len in range(1, 10):
            literal = 1 - int(np.clip(cur_len - min_length, 0, 1))
            assert gen._min_length_applies(cur_len, min_length) == bool(literal)
            s = rg.apply_processors(np.zeros((1, 5), np.float32), cur_len, min_length=min_length, eos_token_id=2,
                                    forced_bos_token_id=None, forced_eos_token_id=None, max_length=64)
            assert (s[0, 2] == -np.inf) == bool(literal)
    assert not gen._min_length_applies(1, 0)            # the default min_length = 0 never masks
