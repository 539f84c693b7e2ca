"""`FlaxViTBartForConditionalGeneration` — the `flax_vit_bart` variant (`models/flax_vit_bart/modeling_vit_bart.py`):
ViT-B/16 encoder (197 tokens, channel-first pixels transposed at :445, exact gelu, final `layernorm` on the sequence)
+ `visual_projection` + post-LN BART decoder, on the same sm_100a engine as the CLIP-mBART model.

What differs from `FlaxCLIPVisionMBartForConditionalGeneration` is configuration (engine switches) and the PARAMETER
NAMES: the reference's tree is that of `FlaxViTModule` / `FlaxBartDecoder` (:33-50), i.e. the names of the HF
modules [MEMORY — cross-checked against the PyTorch twins' attribute names, $T/models/vit/modeling_vit.py,
$T/models/bart/modeling_bart.py]:

  model/encoder/embeddings/{cls_token (1,1,d), position_embeddings (1,197,d), patch_embeddings/projection/{kernel,bias}}
  model/encoder/encoder/layer/<i>/{attention/attention/{query,key,value}, attention/output/dense, intermediate/dense,
                                   output/dense, layernorm_before, layernorm_after}
  model/encoder/layernorm, model/encoder/pooler/dense           (pooler: present in the checkpoint, unused -> zero grad)
  model/decoder/{embed_positions, layernorm_embedding, layers/<i>/...}      (no final `layer_norm`: BART has none)
  model/shared/embedding, model/visual_projection/{kernel,bias}, final_logits_bias

`.params` exposes exactly this tree as live views of the flat buffers (a pure renaming / reshaping of the engine's
canonical tree); checkpoints written by `save_pretrained` therefore load in the reference and vice versa.

`encode()` is done CORRECTLY here: the reference's (:292-300) forgets both the NCHW->NHWC transpose and
`visual_projection`, so its `generate()` cannot run with d_enc != d_dec; ours matches what `__call__` computes.
"""
from __future__ import annotations

import torch

from . import checkpoint as ck
from .configuration import CLIPVisionConfig, CLIPVisionMBartConfig, MBartConfig, vit_bart_config
from .modeling_clip_vision_mbart import (FlaxCLIPVisionMBartForConditionalGeneration, _config_from_dict)

_VIT_LAYER = {          # reference name -> canonical (CLIP-style) name inside one encoder layer
    ("attention", "attention", "query"): ("self_attn", "q_proj"),
    ("attention", "attention", "key"): ("self_attn", "k_proj"),
    ("attention", "attention", "value"): ("self_attn", "v_proj"),
    ("attention", "output", "dense"): ("self_attn", "out_proj"),
    ("intermediate", "dense"): ("mlp", "fc1"),
    ("output", "dense"): ("mlp", "fc2"),
    ("layernorm_before",): ("layer_norm1",),
    ("layernorm_after",): ("layer_norm2",),
}


def _get(tree, path):
    for k in path:
        tree = tree[k]
    return tree


def _put(tree, path, v):
    for k in path[:-1]:
        tree = tree.setdefault(k, {})
    tree[path[-1]] = v


def _view(x, shape):
    return x.reshape(shape) if not isinstance(x, torch.Tensor) else x.view(shape)


def canonical_to_vit_bart(tree, pooler):
    """Engine tree (CLIP-style names) -> the reference's ViT-BART names.  Leaves are re-used (views), not copied."""
    vm = tree["model"]["encoder"]["vision_model"]
    dv = vm["embeddings"]["class_embedding"].shape[0]
    S = vm["embeddings"]["position_embedding"]["embedding"].shape[0]
    enc = {"embeddings": {"cls_token": _view(vm["embeddings"]["class_embedding"], (1, 1, dv)),
                          "position_embeddings": _view(vm["embeddings"]["position_embedding"]["embedding"], (1, S, dv)),
                          "patch_embeddings": {"projection": dict(vm["embeddings"]["patch_embedding"])}},
           "encoder": {"layer": {}}, "layernorm": dict(vm["post_layernorm"]), "pooler": {"dense": dict(pooler)}}
    for i, lp in vm["encoder"]["layers"].items():
        out = {}
        for ref_path, can_path in _VIT_LAYER.items():
            _put(out, ref_path, dict(_get(lp, can_path)))
        enc["encoder"]["layer"][i] = out
    dec = {k: v for k, v in tree["model"]["decoder"].items() if k != "layer_norm"}
    return {"model": {"encoder": enc, "decoder": dec, "shared": tree["model"]["shared"],
                      "visual_projection": tree["model"]["visual_projection"]},
            "final_logits_bias": tree["final_logits_bias"]}


def vit_bart_to_canonical(tree, canonical_template):
    """The reference's ViT-BART names -> engine tree.  Entries the variant does not have (`pre_layrnorm`, the
    decoder's final `layer_norm`) are taken from `canonical_template` (unused identity parameters)."""
    enc = tree["model"]["encoder"]
    tv = canonical_template["model"]["encoder"]["vision_model"]
    dv = tv["embeddings"]["class_embedding"].shape[0]
    S = tv["embeddings"]["position_embedding"]["embedding"].shape[0]
    vm = {"embeddings": {"class_embedding": _view(enc["embeddings"]["cls_token"], (dv,)),
                         "position_embedding": {"embedding": _view(enc["embeddings"]["position_embeddings"], (S, dv))},
                         "patch_embedding": dict(enc["embeddings"]["patch_embeddings"]["projection"])},
          "pre_layrnorm": tv["pre_layrnorm"], "post_layernorm": dict(enc["layernorm"]), "encoder": {"layers": {}}}
    for i, lp in enc["encoder"]["layer"].items():
        out = {}
        for ref_path, can_path in _VIT_LAYER.items():
            _put(out, can_path, dict(_get(lp, ref_path)))
        vm["encoder"]["layers"][i] = out
    dec = dict(tree["model"]["decoder"])
    dec["layer_norm"] = canonical_template["model"]["decoder"]["layer_norm"]
    return {"model": {"encoder": {"vision_model": vm}, "decoder": dec, "shared": tree["model"]["shared"],
                      "visual_projection": tree["model"]["visual_projection"]},
            "final_logits_bias": tree["final_logits_bias"]}


class FlaxViTBartForConditionalGeneration(FlaxCLIPVisionMBartForConditionalGeneration):
    """Same call surface as the CLIP-mBART class (`__call__` :418-481 takes CHANNEL-FIRST pixels, `generate`,
    `encode`, `decode`, `init_cache`, `.params`, `save_pretrained`, `from_pretrained`, `from_vit_bart_pretrained`)."""

    def __init__(self, config: CLIPVisionMBartConfig = None, input_shape=None, seed: int = 0, dtype="bfloat16",
                 device="cuda", _do_init: bool = True):
        config = vit_bart_config() if config is None else config
        v, t = config.clip_vision_config, config.mbart_config
        if v.pre_layernorm or not v.final_layernorm or not v.channel_first_input or t.pre_layernorm or t.final_layer_norm:
            raise ValueError("FlaxViTBartForConditionalGeneration needs a ViT-BART config (mic_b200.vit_bart_config()): "
                             "ViT without pre-LN and with the final layernorm, channel-first pixels, post-LN BART decoder")
        super().__init__(config, input_shape, seed, dtype, device, _do_init)
        dv = v.hidden_size
        g = torch.Generator(device=self.device).manual_seed(int(seed) + 1)
        # FlaxViTPooler (add_pooling_layer=True, :40): lives in the checkpoint, never reaches the loss
        self._pooler = {"kernel": torch.empty((dv, dv), device=self.device).normal_(0.0, v.initializer_range, generator=g),
                        "bias": torch.zeros((dv,), device=self.device)}

    # the reference's names in, the reference's names out
    @property
    def params(self):
        return canonical_to_vit_bart(self.store.tree(), self._pooler)

    @params.setter
    def params(self, tree):
        if "vision_model" in tree.get("model", {}).get("encoder", {}):      # canonical tree (tests, synthetic.make_params)
            self.store.load_tree(tree)
            return
        pool = tree["model"]["encoder"].get("pooler", {}).get("dense")
        if pool is not None:
            for k in ("kernel", "bias"):
                src = pool[k]
                src = torch.as_tensor(src) if not isinstance(src, torch.Tensor) else src
                self._pooler[k].copy_(src.to(self.device, torch.float32))
        self.store.load_tree(vit_bart_to_canonical(tree, self.store.tree()))

    def _is_own_tree(self, params):
        flb = params.get("final_logits_bias") if isinstance(params, dict) else None
        return isinstance(flb, torch.Tensor) and flb.data_ptr() == self.store.tree()["final_logits_bias"].data_ptr()

    @property
    def grads(self):
        """Gradient tree with the reference's names (pooler: zeros — it is not on the path to the loss)."""
        zero = {k: torch.zeros_like(v) for k, v in self._pooler.items()}
        return canonical_to_vit_bart(self.store.tree(self.store.grad), zero)

    @classmethod
    def _config_from_json(cls, d):
        base = vit_bart_config()
        v = _config_from_dict(CLIPVisionConfig, {**base.clip_vision_config.__dict__, **d.get("vit_config", d.get("clip_vision_config", {}))})
        t = _config_from_dict(MBartConfig, {**base.mbart_config.__dict__, **d.get("bart_config", d.get("mbart_config", {}))})
        return CLIPVisionMBartConfig(v, t, model_type="vit-bart")

    def save_pretrained(self, save_directory, params=None, push_to_hub=False, **kwargs):
        if push_to_hub:
            raise NotImplementedError("no hub access in this build")
        cfg = self.config.to_dict()
        cfg["vit_config"], cfg["bart_config"] = cfg.pop("clip_vision_config"), cfg.pop("mbart_config")
        cfg["architectures"] = [type(self).__name__[4:]]
        return ck.write_weights(str(save_directory), params if params is not None else self.params, cfg)

    @classmethod
    def from_vit_bart_pretrained(cls, vit_model_name_or_path=None, bart_model_name_or_path=None, *model_args, **kwargs):
        """`from_vit_bart_pretrained` (modeling_vit_bart.py:650-732): graft a FlaxViTModel checkpoint
        (params = the FlaxViTModule tree) and a FlaxBartModel checkpoint (params["decoder"], params["shared"])."""
        kw_vit = {k[len("vit_"):]: v for k, v in kwargs.items() if k.startswith("vit_")}
        kw_bart = {k[len("bart_"):]: v for k, v in kwargs.items() if k.startswith("bart_")}
        for k in kw_vit:
            del kwargs["vit_" + k]
        for k in kw_bart:
            del kwargs["bart_" + k]

        def load(kw, path, what):
            obj = kw.pop("model", None)
            if obj is not None:
                return obj if isinstance(obj, tuple) else (obj.params, obj.config)
            assert path is not None, f"If `model` is not defined as an argument, a `{what}_model_name_or_path` has to be defined"
            d = ck.resolve_local_dir(path)
            return ck.read_weights(d), kw.pop("config", None) or ck.read_config_dict(d)

        vit_params, vit_cfg = load(kw_vit, vit_model_name_or_path, "vit")
        bart_params, bart_cfg = load(kw_bart, bart_model_name_or_path, "bart")
        config = cls._config_from_json({"vit_config": vit_cfg if isinstance(vit_cfg, dict) else vit_cfg.__dict__,
                                        "bart_config": bart_cfg if isinstance(bart_cfg, dict) else bart_cfg.__dict__})
        dtype = kwargs.pop("dtype", "bfloat16")
        model = cls(config, *model_args, dtype=dtype, **kwargs)
        tree = canonical_to_vit_bart(model.store.to_numpy_tree(), {k: v.cpu().numpy() for k, v in model._pooler.items()})
        tree["model"]["encoder"] = vit_params
        tree["model"]["decoder"] = bart_params["decoder"]
        tree["model"]["shared"] = bart_params["shared"]
        model.params = tree
        return model
