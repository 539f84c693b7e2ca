"""Checkpoint interop with the reference: the msgpack container that
`flax.serialization.to_bytes / from_bytes` writes, without flax or jax.

Replaces, for LOCAL directories (there is no hub access here):
  * `FlaxPreTrainedModel.save_pretrained`   models/flax_clip_vision_mbart/modeling_clip_vision_utils.py:398-451
      -> <dir>/config.json + <dir>/flax_model.msgpack
  * `FlaxPreTrainedModel.from_pretrained`   modeling_clip_vision_utils.py:120-396 (missing keys keep their initial
      value, unexpected keys are dropped, both are reported — :340-396)
  * `save_model_checkpoint` / `restore_model_checkpoint`   main.py:299-346
      -> ckpt-<step>/{flax_model.msgpack, opt_state.msgpack, training_state.json}

Container format (flax/serialization.py, published; flax==0.3.4 pinned by requirements.txt:10): a msgpack map of
maps whose keys are strings; every array leaf is ExtType(1, packb((shape, dtype.name, raw C-order bytes)));
numpy scalars are ExtType(3, same triple); arrays above 2**30 bytes are stored as
{"__msgpack_chunked_array__": True, "shape": {"0": d0, ...}, "chunks": {"0": flat chunk, ...}}; lists / tuples
(e.g. the optax state chain) become maps keyed "0", "1", ...  [MEMORY for the ext codes; round-trip tested here].
"""
from __future__ import annotations

import json
import os

import msgpack
import numpy as np

FLAX_WEIGHTS_NAME = "flax_model.msgpack"
CONFIG_NAME = "config.json"
_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
MAX_CHUNK_BYTES = 2 ** 30


# ----------------------------------------------------------------------------------------------
# flax.serialization-compatible msgpack
# ----------------------------------------------------------------------------------------------
def _array_payload(a: np.ndarray) -> bytes:
    a = np.asarray(a)
    if a.dtype.hasobject:
        raise ValueError("object arrays cannot be serialised")
    return msgpack.packb((list(a.shape), a.dtype.name, a.tobytes("C")), use_bin_type=True)


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _array_payload(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _array_payload(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(_EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
    raise TypeError(f"cannot serialise leaf of type {type(x)}")


def _dtype_of(name: str):
    if name == "bfloat16":          # numpy has no bfloat16: widen to float32 (the master copy is fp32 anyway)
        return None
    return np.dtype(name)


def _ext_unpack(code, data):
    if code in (_EXT_NDARRAY, _EXT_NPSCALAR):
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        dt = _dtype_of(dtype_name)
        if dt is None:
            u16 = np.frombuffer(buf, dtype=np.uint16)
            arr = (u16.astype(np.uint32) << 16).view(np.float32).reshape(shape)
        else:
            arr = np.frombuffer(buf, dtype=dt).reshape(shape)
        return arr[()] if code == _EXT_NPSCALAR else arr
    if code == _EXT_COMPLEX:
        re, im = msgpack.unpackb(data)
        return complex(re, im)
    return msgpack.ExtType(code, data)


def _to_state_dict(x):
    """Tensors / arrays -> numpy leaves, lists and tuples -> {"0": ...} maps, torch tensors -> host copies."""
    if isinstance(x, dict):
        return {str(k): _to_state_dict(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return {str(i): _to_state_dict(v) for i, v in enumerate(x)}
    if hasattr(x, "detach") and hasattr(x, "cpu"):      # torch tensor (possibly a strided view of the flat buffer)
        x = x.detach().cpu().contiguous().numpy()
    if isinstance(x, (np.ndarray, np.generic)):
        a = np.asarray(x)
        if a.ndim and a.size * a.dtype.itemsize > MAX_CHUNK_BYTES:
            per = max(1, MAX_CHUNK_BYTES // a.dtype.itemsize)
            flat = np.ascontiguousarray(a).reshape(-1)
            return {"__msgpack_chunked_array__": True, "shape": {str(i): int(d) for i, d in enumerate(a.shape)},
                    "chunks": {str(i): flat[lo:lo + per] for i, lo in enumerate(range(0, flat.size, per))}}
        return np.ascontiguousarray(a) if a.ndim else a[()]
    if isinstance(x, (bool, int, float, str)) or x is None:
        return x
    raise TypeError(f"cannot serialise {type(x)}")


def _unchunk(x):
    if isinstance(x, dict):
        if x.get("__msgpack_chunked_array__"):
            shape = tuple(x["shape"][str(i)] for i in range(len(x["shape"])))
            parts = [x["chunks"][str(i)] for i in range(len(x["chunks"]))]
            return np.concatenate([np.asarray(p).reshape(-1) for p in parts]).reshape(shape)
        return {k: _unchunk(v) for k, v in x.items()}
    return x


def to_bytes(tree) -> bytes:
    """flax.serialization.to_bytes for nested dicts / lists of arrays."""
    return msgpack.packb(_to_state_dict(tree), default=_ext_pack, strict_types=True, use_bin_type=True)


def msgpack_restore(data: bytes):
    """flax.serialization.msgpack_restore: bytes -> nested dict of numpy arrays."""
    return _unchunk(msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


def from_bytes(target, data: bytes):
    """flax.serialization.from_bytes: restore INTO the structure of `target` — the stored tree must have exactly
    target's keys (flax raises on a mismatch); returns a tree of numpy arrays shaped like `target`."""
    state = msgpack_restore(data)

    def rec(t, s, path):
        if isinstance(t, (list, tuple)):
            t = {str(i): v for i, v in enumerate(t)}
        if isinstance(t, dict):
            if not isinstance(s, dict) or set(map(str, t)) != set(s):
                raise ValueError(f"The target dict keys and state dict keys do not match at {'/'.join(path) or '<root>'}: "
                                 f"{sorted(map(str, t))[:8]} vs {sorted(s)[:8] if isinstance(s, dict) else type(s)}")
            return {k: rec(v, s[str(k)], path + (str(k),)) for k, v in t.items()}
        return s
    return rec(target, state, ())


# ----------------------------------------------------------------------------------------------
# tree helpers
# ----------------------------------------------------------------------------------------------
def flatten(tree, prefix=()):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flatten(v, prefix + (k,)))
        else:
            out[prefix + (k,)] = v
    return out


def unflatten(flat):
    out = {}
    for path, v in flat.items():
        node = out
        for k in path[:-1]:
            node = node.setdefault(k, {})
        node[path[-1]] = v
    return out


def merge_into(model_tree, loaded_tree):
    """`from_pretrained` key reconciliation (modeling_clip_vision_utils.py:340-396): returns (merged tree, missing keys,
    unexpected keys).  Missing keys keep the model's current (initial) values; unexpected keys are dropped."""
    want, have = flatten(model_tree), flatten(loaded_tree)
    missing = sorted(set(want) - set(have))
    unexpected = sorted(set(have) - set(want))
    merged = dict(want)
    for k in want:
        if k in have:
            if tuple(np.shape(have[k])) != tuple(want[k].shape):
                raise ValueError(f"checkpoint tensor {'/'.join(k)} has shape {tuple(np.shape(have[k]))}, the model "
                                 f"expects {tuple(want[k].shape)}")
            merged[k] = have[k]
    return unflatten(merged), missing, unexpected


# ----------------------------------------------------------------------------------------------
# directories
# ----------------------------------------------------------------------------------------------
def resolve_local_dir(name_or_path: str) -> str:
    if not os.path.isdir(str(name_or_path)):
        raise OSError(f"'{name_or_path}' is not a local directory. This build has no network / hub access: pass a "
                      f"directory that holds {CONFIG_NAME} and {FLAX_WEIGHTS_NAME}.")
    return os.path.abspath(str(name_or_path))


def read_config_dict(directory: str) -> dict:
    p = os.path.join(directory, CONFIG_NAME)
    if not os.path.isfile(p):
        raise OSError(f"{p} not found")
    with open(p) as f:
        return json.load(f)


def read_weights(directory: str):
    p = os.path.join(directory, FLAX_WEIGHTS_NAME)
    if not os.path.isfile(p):
        raise OSError(f"Error no file named {FLAX_WEIGHTS_NAME} found in directory {directory}")
    with open(p, "rb") as f:
        return msgpack_restore(f.read())


def write_weights(directory: str, tree, config_dict: dict | None = None):
    if os.path.isfile(directory):
        raise OSError(f"Provided path ({directory}) should be a directory, not a file")
    os.makedirs(directory, exist_ok=True)
    if config_dict is not None:
        with open(os.path.join(directory, CONFIG_NAME), "w") as f:
            json.dump(config_dict, f, indent=2, sort_keys=True)
    with open(os.path.join(directory, FLAX_WEIGHTS_NAME), "wb") as f:
        f.write(to_bytes(tree))
    return os.path.join(directory, FLAX_WEIGHTS_NAME)
