"""Drop-in model class: same public surface as the reference's
`FlaxCLIPVisionMBartForConditionalGeneration` (`models/flax_clip_vision_mbart/modeling_clip_vision_mbart.py`)
— `__call__` (:447-510), `encode` (:284-337), `decode` (:519-651), `init_cache` (:249-282), `generate`
(`generation_clip_vision_utils.py:128-336`), `.params` (`modeling_clip_vision_utils.py:99-117`) —
executing on hand-written sm_100a kernels.  Arrays in / out are torch CUDA tensors (numpy accepted).

There is no CPU path: constructing the model without a CUDA device or without libmic_b200.so raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch

import contextlib
import dataclasses
import logging

from . import checkpoint as ck
from . import generation as gen
from . import ops
from ._lib import lib
from .configuration import CLIPVisionConfig, CLIPVisionMBartConfig, MBartConfig
from .engine import CaptionEngine
from .params import ParamStore

F32, I32 = torch.float32, torch.int32
logger = logging.getLogger(__name__)


@dataclass
class Seq2SeqLMOutput:
    """Shape of transformers' FlaxSeq2SeqLMOutput as consumed by the reference (`[0]` and `.logits`)."""
    logits: Any = None
    past_key_values: Any = None
    encoder_last_hidden_state: Any = None

    def __getitem__(self, i):
        return (self.logits, self.past_key_values, self.encoder_last_hidden_state)[i]


@dataclass
class BaseModelOutput:
    last_hidden_state: Any = None

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


@dataclass
class SearchOutput:
    """FlaxGreedySearchOutput / FlaxBeamSearchOutput."""
    sequences: Any = None
    scores: Any = None


def _pixels(x, device):
    """Pixels keep uint8 (input hand-off: normalised inside the patch kernel); everything else becomes f32, the
    reference's cast (modeling_clip_vision_mbart.py:501)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    return x.to(device) if x.dtype == torch.uint8 else x.to(device=device, dtype=F32)


def _config_from_dict(cls, d):
    """Build a config dataclass from an HF-style config.json dict (unknown keys are ignored)."""
    names = {f.name for f in dataclasses.fields(cls)}
    kw = {k: v for k, v in d.items() if k in names}
    for k in ("image_mean", "image_std"):
        if k in kw:
            kw[k] = tuple(kw[k])
    return cls(**kw)


def _as_tensor(x, device, dtype=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    return x.to(device=device, dtype=dtype) if dtype is not None else x.to(device)


class FlaxCLIPVisionMBartForConditionalGeneration:
    config_class = CLIPVisionMBartConfig
    base_model_prefix = "model"

    def __init__(self, config: CLIPVisionMBartConfig, input_shape=None, seed: int = 0, dtype="bfloat16",
                 device="cuda", _do_init: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("mic_b200 runs on an NVIDIA B200 (sm_100a) only; no CUDA device is visible")
        lib()  # fail loudly if the extension is missing
        if str(dtype) in ("bfloat16", "torch.bfloat16", "bf16"):
            self.dtype = "bfloat16"
        elif str(dtype) in ("float32", "torch.float32", "fp32", "f32"):
            # fp32 VERIFICATION mode (engine_fp32.py): __call__ / loss run with fp32 storage and fp32 SIMT kernels
            # (BASELINE configs[0]: logits within 1e-3, loss within 1e-4 of the oracle); training and generate()
            # exist in bf16 only
            self.dtype = "float32"
        else:
            raise NotImplementedError(f"compute dtype {dtype}: bfloat16 (product path) or float32 (verification path)")
        t = config.mbart_config
        if t.attention_dropout or t.activation_dropout:
            raise NotImplementedError("attention_dropout / activation_dropout are not applied by the fused training tape "
                                      "(the reference's hub configs use 0.0 for both); set them to 0")
        self.config = config
        self.device = torch.device(device)
        self._param_backup = None
        self.store = ParamStore(config, self.device)
        self.engine = CaptionEngine(config, self.store)
        if _do_init:
            self.init_weights(seed)

    # ---- parameters ---------------------------------------------------------------------------
    def init_weights(self, seed: int = 0):
        """Random init as `init_weights` (:224-247) would give: N(0, init_std) kernels/embeddings, zero
        biases, unit LayerNorm scales, zero final_logits_bias (distribution-equivalent, not bit-equal to
        jax.random)."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        ps = self.store
        std_t, std_v = self.config.mbart_config.init_std, self.config.clip_vision_config.initializer_range
        for name in ps.layout.order:
            v = ps.f(name)
            if name.endswith(".scale"):
                v.fill_(1.0)
            elif name.endswith(".bias") or name.endswith(".b") or name == "flb":
                v.zero_()
            else:
                v.normal_(0.0, std_v if name.startswith("v.") else std_t, generator=g)
        ps.refresh_shadow()

    @property
    def params(self):
        """Nested dict with the reference's Flax names; leaves are live views of the fp32 master buffer."""
        return self.store.tree()

    @params.setter
    def params(self, tree):
        self.store.load_tree(tree)

    def _is_own_tree(self, params):
        flb = params.get("final_logits_bias") if isinstance(params, dict) else None
        return isinstance(flb, torch.Tensor) and flb.data_ptr() == self.store.tree()["final_logits_bias"].data_ptr()

    def _use_params_permanently(self, params):
        if params is not None and not self._is_own_tree(params):
            self.params = params

    @contextlib.contextmanager
    def _params_scope(self, params):
        """`params=` is a PER-CALL override in the reference (pure-functional apply): the model's own parameters are
        untouched afterwards.  Passing the model's own live tree (state.params) costs nothing; a foreign tree is
        loaded for the duration of the call and the previous weights are put back (two device copies of 3.3 GB)."""
        if params is None or self._is_own_tree(params):
            yield
            return
        ps = self.store
        if self._param_backup is None:
            self._param_backup = (torch.empty_like(ps.master), torch.empty_like(ps.shadow))
        self._param_backup[0].copy_(ps.master)
        self._param_backup[1].copy_(ps.shadow)
        try:
            self.params = params
            yield
        finally:
            ps.master.copy_(self._param_backup[0])
            ps.shadow.copy_(self._param_backup[1])

    # ---- forward ------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, pixel_values, decoder_input_ids=None, decoder_attention_mask=None,
                 decoder_position_ids=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                 train: bool = False, params: Optional[dict] = None, dropout_rng=None):
        """:447-510.  train=True applies the decoder's dropout (mbart_config.dropout) with a mask seeded from
        `dropout_rng` — a counter-hash mask, NOT jax's threefry stream, so only the distribution matches."""
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are not materialised on the fused path")
        with self._params_scope(params):
            eng = self.engine
            px = self._prepare_pixels(pixel_values)
            ids = _as_tensor(decoder_input_ids, self.device, I32)
            B, T = ids.shape
            mask = torch.ones((B, T), dtype=I32, device=self.device) if decoder_attention_mask is None else \
                _as_tensor(decoder_attention_mask, self.device, I32)
            if self.dtype == "float32":
                if decoder_position_ids is not None or train or px.dtype == torch.uint8:
                    raise NotImplementedError("fp32 verification mode: default positions, float pixels, inference only")
                from .engine_fp32 import Fp32Forward
                return Seq2SeqLMOutput(logits=Fp32Forward(eng).logits(px, ids, mask), encoder_last_hidden_state=None)
            pos = None if decoder_position_ids is None else \
                _as_tensor(decoder_position_ids, self.device, I32).contiguous().view(-1)
            drop = float(self.config.mbart_config.dropout) if train else 0.0
            if drop > 0.0:
                if dropout_rng is None:
                    raise ValueError("train=True needs a `dropout_rng` (main.py:686-692)")
                from .training import _rng_to_int
                eng.dropout_p = drop
                eng.drop_seed.fill_(_rng_to_int(dropout_rng))
            try:
                enc = eng.encode(px, trunc_int=False, save=False, tag="fw.enc")
                enc_kv = eng.cross_kv(enc, tag="fw.enc")
                hf = eng.decoder_forward(ids.contiguous().view(-1), mask.contiguous(), pos, enc_kv, B, T,
                                         self.config.clip_vision_config.num_tokens, save=False, tag="fw.dec",
                                         train=drop > 0.0)
            finally:
                eng.dropout_p = 0.0
            logits = eng.logits(hf).view(B, T, -1)
            return Seq2SeqLMOutput(logits=logits, encoder_last_hidden_state=enc.view(B, -1, enc.shape[-1]))

    def _prepare_pixels(self, pixel_values):
        return _pixels(pixel_values, self.device)

    @torch.no_grad()
    def loss(self, pixel_values, decoder_input_ids, attention_mask, labels, label_smoothing_factor=0.0, params=None):
        """eval_step (main.py:710-721): forward + loss_fn without materialising logits. Returns a 0-d tensor."""
        with self._params_scope(params):
            eng = self.engine
            px = self._prepare_pixels(pixel_values)
            ids = _as_tensor(decoder_input_ids, self.device, I32)
            B, T = ids.shape
            mask = _as_tensor(attention_mask, self.device, I32).contiguous()
            if self.dtype == "float32":
                from .engine_fp32 import Fp32Forward
                return Fp32Forward(eng).loss(px, ids, mask, _as_tensor(labels, self.device, I32),
                                             label_smoothing_factor)[0].float()
            lab = _as_tensor(labels, self.device, I32).contiguous().view(-1)
            enc = eng.encode(px, trunc_int=False, save=False, tag="fw.enc")
            enc_kv = eng.cross_kv(enc, tag="fw.enc")
            hf = eng.decoder_forward(ids.contiguous().view(-1), mask, None, enc_kv, B, T,
                                     self.config.clip_vision_config.num_tokens, save=False, tag="fw.dec")
            ws = eng.loss_forward(hf, lab, mask.view(-1), label_smoothing_factor)
            return ws["out"][0].clone()

    @torch.no_grad()
    def encode(self, pixel_values, output_attentions=None, output_hidden_states=None, return_dict=None,
               train: bool = False, params=None, dropout_rng=None):
        """:284-337 — note the int32 cast of the pixels at :330 is reproduced."""
        with self._params_scope(params):
            px = self._prepare_pixels(pixel_values)
            enc = self.engine.encode(px, trunc_int=True, save=False, tag="gen.enc")
            return BaseModelOutput(last_hidden_state=enc.view(px.shape[0], -1, enc.shape[-1]).clone())

    @torch.no_grad()
    def init_cache(self, batch_size, max_length, encoder_outputs):
        """:249-282 — zero self-attention cache + cross K/V of the given encoder states."""
        enc = encoder_outputs[0]
        enc2d = enc.reshape(-1, enc.shape[-1]).contiguous()
        enc_kv = self.engine.cross_kv(enc2d, tag="gen.enc")
        rows_per_image = max(batch_size // enc.shape[0], 1)
        return gen.DecodeCache(self.engine, batch_size, max_length, enc_kv, rows_per_image, use_ancestors=False)

    @torch.no_grad()
    def decode(self, decoder_input_ids, encoder_outputs, encoder_attention_mask=None, decoder_attention_mask=None,
               decoder_position_ids=None, past_key_values=None, output_attentions=None, output_hidden_states=None,
               return_dict=None, train: bool = False, params=None, dropout_rng=None):
        """:519-651 — with `past_key_values` (a DecodeCache from init_cache) runs the cached 1-token step."""
        with self._params_scope(params):
            ids = _as_tensor(decoder_input_ids, self.device, I32)
            eng = self.engine
            if past_key_values is None:
                B, T = ids.shape
                enc = encoder_outputs[0]
                enc_kv = eng.cross_kv(enc.reshape(-1, enc.shape[-1]).contiguous(), tag="fw.enc")
                mask = torch.ones((B, T), dtype=I32, device=self.device) if decoder_attention_mask is None else \
                    _as_tensor(decoder_attention_mask, self.device, I32).contiguous()
                pos = None if decoder_position_ids is None else \
                    _as_tensor(decoder_position_ids, self.device, I32).contiguous().view(-1)
                hf = eng.decoder_forward(ids.contiguous().view(-1), mask, pos, enc_kv, B, T, enc.shape[1], save=False,
                                         tag="fw.dec")
                return Seq2SeqLMOutput(logits=eng.logits(hf).view(B, T, -1))
            if decoder_position_ids is None:
                raise ValueError("Make sure to provide `decoder_position_ids` when passing `past_key_values`.")
            if ids.shape[1] != 1:
                raise NotImplementedError("cached decode handles one token per row")
            pos = int(_as_tensor(decoder_position_ids, "cpu").reshape(-1)[0])
            hf = gen.decode_step(eng, past_key_values, ids.contiguous().view(-1), pos)
            past_key_values.index = pos + 1
            return Seq2SeqLMOutput(logits=eng.logits(hf).view(ids.shape[0], 1, -1), past_key_values=past_key_values)

    # ---- generation hooks (:653-693): what a caller-side search loop over decode() uses ----------
    def prepare_inputs_for_generation(self, decoder_input_ids, max_length, attention_mask=None,
                                      decoder_attention_mask=None, encoder_outputs=None, **kwargs):
        """:653-686 — cache from init_cache, ONE static all-ones mask of max_length (the causal mask hides the
        future), positions = cumsum(mask) - 1 or arange."""
        ids = _as_tensor(decoder_input_ids, self.device, I32)
        batch_size, seq_length = ids.shape
        past_key_values = self.init_cache(batch_size, max_length, encoder_outputs)
        extended = torch.ones((batch_size, max_length), dtype=I32, device=self.device)
        if decoder_attention_mask is not None:
            dam = _as_tensor(decoder_attention_mask, self.device, I32)
            position_ids = dam.cumsum(dim=-1, dtype=I32) - 1
            extended[:, :dam.shape[1]] = dam
        else:
            position_ids = torch.arange(seq_length, dtype=I32, device=self.device)[None, :].expand(batch_size, seq_length)
        return {"past_key_values": past_key_values, "encoder_outputs": encoder_outputs,
                "encoder_attention_mask": attention_mask, "decoder_attention_mask": extended,
                "decoder_position_ids": position_ids}

    def update_inputs_for_generation(self, model_outputs, model_kwargs):
        """:688-693."""
        model_kwargs["past_key_values"] = model_outputs.past_key_values
        model_kwargs["decoder_position_ids"] = model_kwargs["decoder_position_ids"][:, -1:] + 1
        return model_kwargs

    # ---- generation ---------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids, max_length=None, pad_token_id=None, bos_token_id=None, eos_token_id=None,
                 decoder_start_token_id=None, do_sample=None, prng_key=None, top_k=None, top_p=None, temperature=None,
                 num_beams=None, no_repeat_ngram_size=None, min_length=None, forced_bos_token_id=None,
                 forced_eos_token_id=None, length_penalty=None, early_stopping=None, trace: bool = True, params=None,
                 **model_kwargs):
        """`generate(input_ids=<pixel_values>, ...)` — defaults resolved from config.mbart_config exactly as
        generation_clip_vision_utils.py:196-229,386-409 does; dispatch greedy :254 / sample :273 / beam :297."""
        t = self.config.mbart_config
        max_length = max_length if max_length is not None else t.max_length
        pad_token_id = pad_token_id if pad_token_id is not None else t.pad_token_id
        eos_token_id = eos_token_id if eos_token_id is not None else t.eos_token_id
        decoder_start_token_id = decoder_start_token_id if decoder_start_token_id else t.decoder_start_token_id
        if decoder_start_token_id is None and self.config.is_encoder_decoder:
            raise ValueError("`decoder_start_token_id` has to be defined for encoder-decoder generation.")
        do_sample = do_sample if do_sample is not None else t.do_sample
        num_beams = num_beams if num_beams is not None else t.num_beams
        min_length = min_length if min_length is not None else t.min_length
        forced_bos_token_id = forced_bos_token_id if forced_bos_token_id is not None else t.forced_bos_token_id
        forced_eos_token_id = forced_eos_token_id if forced_eos_token_id is not None else t.forced_eos_token_id
        length_penalty = length_penalty if length_penalty is not None else t.length_penalty
        early_stopping = early_stopping if early_stopping is not None else t.early_stopping
        if do_sample and num_beams != 1:
            raise NotImplementedError("`Beam sampling is currently not implemented.")          # :336
        with self._params_scope(params):
            px = self._prepare_pixels(input_ids)
            if do_sample:
                out = gen.generate(self.engine, px, max_length=max_length, pad_token_id=pad_token_id,
                                   eos_token_id=eos_token_id, decoder_start_token_id=decoder_start_token_id,
                                   num_beams=1, min_length=None, forced_bos_token_id=None, forced_eos_token_id=None,
                                   length_penalty=1.0, early_stopping=False,
                                   sample_key=gen.prng_key_pair(prng_key))
                return SearchOutput(sequences=out["sequences"])
            out = gen.generate(self.engine, px, max_length=max_length, pad_token_id=pad_token_id,
                               eos_token_id=eos_token_id, decoder_start_token_id=decoder_start_token_id,
                               num_beams=num_beams, min_length=min_length, forced_bos_token_id=forced_bos_token_id,
                               forced_eos_token_id=forced_eos_token_id, length_penalty=length_penalty,
                               early_stopping=early_stopping)
            return SearchOutput(sequences=out["sequences"], scores=out.get("scores"))

    # ---- checkpoints (modeling_clip_vision_utils.py:120-451, local directories only) ------------
    def _checkpoint_tree(self, params=None):
        """The tree written to flax_model.msgpack: the reference's own names."""
        return params if params is not None else self.params

    def save_pretrained(self, save_directory, params=None, push_to_hub=False, **kwargs):
        """:398-451 — config.json + flax_model.msgpack (flax.serialization container, checkpoint.py)."""
        if push_to_hub:
            raise NotImplementedError("no hub access in this build")
        cfg = self.config.to_dict()
        cfg["architectures"] = [type(self).__name__[4:]]
        return ck.write_weights(str(save_directory), self._checkpoint_tree(params), cfg)

    @classmethod
    def _config_from_json(cls, d):
        return CLIPVisionMBartConfig(
            _config_from_dict(CLIPVisionConfig, d.get("clip_vision_config", d.get("vit_config", {}))),
            _config_from_dict(MBartConfig, d.get("mbart_config", d.get("bart_config", {}))),
            model_type=d.get("model_type", "clip-vision-mbart"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, dtype="bfloat16", *model_args, config=None, seed: int = 0,
                        **kwargs):
        """:120-396 for a LOCAL directory: config.json -> config, random-init model, then every checkpoint tensor whose
        name the model knows replaces the initial value; missing and unexpected names are reported (:340-396)."""
        d = ck.resolve_local_dir(pretrained_model_name_or_path)
        if config is None:
            config = cls._config_from_json(ck.read_config_dict(d))
        model = cls(config, dtype=dtype, seed=seed, **kwargs)
        model._load_checkpoint_tree(ck.read_weights(d), d)
        return model

    def _load_checkpoint_tree(self, loaded, where=""):
        merged, missing, unexpected = ck.merge_into(self.params, loaded)
        if unexpected:
            logger.warning("Some weights of the model checkpoint at %s were not used when initializing %s: %s", where,
                           type(self).__name__, ["/".join(k) for k in unexpected][:20])
        if missing:
            logger.warning("Some weights of %s were not initialized from the model checkpoint at %s and are newly "
                           "initialized: %s", type(self).__name__, where, ["/".join(k) for k in missing][:20])
        self.params = merged
        self.missing_keys, self.unexpected_keys = missing, unexpected

    @classmethod
    def from_clip_vision_mbart_pretrained(cls, clip_vision_model_name_or_path=None, mbart_model_name_or_path=None,
                                          *model_args, **kwargs):
        """:702-773 — build the composite from a FlaxCLIPVisionModel checkpoint and a FlaxMBartModel checkpoint and
        GRAFT their sub-trees (:768-770): params["model"]["encoder"] <- clip.params, ["decoder"] <- mbart.params
        ["decoder"], ["shared"] <- mbart.params["shared"]; `visual_projection` and `final_logits_bias` keep their
        initial values.  Sources are local directories (config.json + flax_model.msgpack) or, as in the reference,
        ready objects passed as `clip_vision_model=` / `mbart_model=` (anything with `.params` and `.config`, or a
        (params, config) pair)."""
        kwargs_mbart = {k[len("mbart_"):]: v for k, v in kwargs.items() if k.startswith("mbart_")}
        kwargs_clip = {k[len("clip_vision_"):]: v for k, v in kwargs.items() if k.startswith("clip_vision_")}
        for k in kwargs_mbart:
            del kwargs["mbart_" + k]
        for k in kwargs_clip:
            del kwargs["clip_vision_" + k]

        def load(kw, path, what, config_cls, sub):
            obj = kw.pop("model", None)
            if obj is not None:
                params, cfg = (obj if isinstance(obj, tuple) else (obj.params, obj.config))
            else:
                assert path is not None, f"If `model` is not defined as an argument, a `{what}_model_name_or_path` " \
                                         "has to be defined"
                d = ck.resolve_local_dir(path)
                cfg = kw.pop("config", None) or ck.read_config_dict(d)
                params = ck.read_weights(d)
            if isinstance(cfg, dict):
                cfg = _config_from_dict(config_cls, cfg.get(sub, cfg))
            return params, cfg

        mbart_params, mbart_config = load(kwargs_mbart, mbart_model_name_or_path, "mbart", MBartConfig, "mbart_config")
        clip_params, clip_config = load(kwargs_clip, clip_vision_model_name_or_path, "clip_vision", CLIPVisionConfig,
                                        "vision_config")
        dtype = kwargs.pop("dtype", "bfloat16")
        config = CLIPVisionMBartConfig.from_clip_vision_mbart_configs(clip_config, mbart_config)
        model = cls(config, *model_args, dtype=dtype, **kwargs)
        tree = model.store.to_numpy_tree()
        tree["model"]["encoder"] = clip_params
        tree["model"]["decoder"] = mbart_params["decoder"]
        tree["model"]["shared"] = mbart_params["shared"]
        model.params = tree
        return model
