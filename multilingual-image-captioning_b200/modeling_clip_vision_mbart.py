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

from . import generation as gen
from . import ops
from ._lib import lib
from .configuration import CLIPVisionMBartConfig
from .engine import CaptionEngine
from .params import ParamStore

F32, I32 = torch.float32, torch.int32


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
        self.config = config
        self.device = torch.device(device)
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

    def _use_params(self, params):
        if params is None:
            return
        mine = self.store.tree()
        if params.get("final_logits_bias") is not None and isinstance(params["final_logits_bias"], torch.Tensor) and \
                params["final_logits_bias"].data_ptr() == mine["final_logits_bias"].data_ptr():
            return  # the caller passed our own live tree back (state.params) — nothing to copy
        self.store.load_tree(params)

    # ---- forward ------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, pixel_values, decoder_input_ids=None, decoder_attention_mask=None,
                 decoder_position_ids=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                 train: bool = False, params: Optional[dict] = None, dropout_rng=None):
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are not materialised on the fused path")
        self._use_params(params)
        eng = self.engine
        px = _as_tensor(pixel_values, self.device, F32)
        ids = _as_tensor(decoder_input_ids, self.device, I32)
        B, T = ids.shape
        mask = torch.ones((B, T), dtype=I32, device=self.device) if decoder_attention_mask is None else \
            _as_tensor(decoder_attention_mask, self.device, I32)
        if self.dtype == "float32":
            if decoder_position_ids is not None or train:
                raise NotImplementedError("fp32 verification mode: default positions, inference only")
            from .engine_fp32 import Fp32Forward
            return Seq2SeqLMOutput(logits=Fp32Forward(eng).logits(px, ids, mask), encoder_last_hidden_state=None)
        pos = None if decoder_position_ids is None else _as_tensor(decoder_position_ids, self.device, I32).contiguous().view(-1)
        enc = eng.encode(px, trunc_int=False, save=False, tag="fw.enc")
        enc_kv = eng.cross_kv(enc, tag="fw.enc")
        hf = eng.decoder_forward(ids.contiguous().view(-1), mask.contiguous(), pos, enc_kv, B, T,
                                 self.config.clip_vision_config.num_tokens, save=False, tag="fw.dec")
        logits = eng.logits(hf).view(B, T, -1)
        return Seq2SeqLMOutput(logits=logits, encoder_last_hidden_state=enc.view(B, -1, enc.shape[-1]))

    @torch.no_grad()
    def loss(self, pixel_values, decoder_input_ids, attention_mask, labels, label_smoothing_factor=0.0, params=None):
        """eval_step (main.py:710-721): forward + loss_fn without materialising logits. Returns a 0-d tensor."""
        self._use_params(params)
        eng = self.engine
        px = _as_tensor(pixel_values, self.device, F32)
        ids = _as_tensor(decoder_input_ids, self.device, I32)
        B, T = ids.shape
        mask = _as_tensor(attention_mask, self.device, I32).contiguous()
        if self.dtype == "float32":
            from .engine_fp32 import Fp32Forward
            return Fp32Forward(eng).loss(px, ids, mask, _as_tensor(labels, self.device, I32), label_smoothing_factor)[0].float()
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
        self._use_params(params)
        px = _as_tensor(pixel_values, self.device, F32)
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
        self._use_params(params)
        ids = _as_tensor(decoder_input_ids, self.device, I32)
        eng = self.engine
        if past_key_values is None:
            B, T = ids.shape
            enc = encoder_outputs[0]
            enc_kv = eng.cross_kv(enc.reshape(-1, enc.shape[-1]).contiguous(), tag="fw.enc")
            mask = torch.ones((B, T), dtype=I32, device=self.device) if decoder_attention_mask is None else \
                _as_tensor(decoder_attention_mask, self.device, I32).contiguous()
            pos = None if decoder_position_ids is None else _as_tensor(decoder_position_ids, self.device, I32).contiguous().view(-1)
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

    # ---- generation ---------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids, max_length=None, pad_token_id=None, bos_token_id=None, eos_token_id=None,
                 decoder_start_token_id=None, do_sample=None, prng_key=None, top_k=None, top_p=None, temperature=None,
                 num_beams=None, no_repeat_ngram_size=None, min_length=None, forced_bos_token_id=None,
                 forced_eos_token_id=None, length_penalty=None, early_stopping=None, trace: bool = True, params=None,
                 **model_kwargs):
        """`generate(input_ids=<pixel_values>, ...)` — defaults resolved from config.mbart_config exactly as
        generation_clip_vision_utils.py:196-229,386-409 does."""
        t = self.config.mbart_config
        self._use_params(params)
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
        if do_sample and num_beams == 1:
            raise NotImplementedError("sampling (`_sample` :537-663) is outside the hot path built here")
        if do_sample:
            raise NotImplementedError("`Beam sampling is currently not implemented.")
        px = _as_tensor(input_ids, self.device, F32)
        out = gen.generate(self.engine, px, max_length=max_length, pad_token_id=pad_token_id,
                           eos_token_id=eos_token_id, decoder_start_token_id=decoder_start_token_id,
                           num_beams=num_beams, min_length=min_length, forced_bos_token_id=forced_bos_token_id,
                           forced_eos_token_id=forced_eos_token_id, length_penalty=length_penalty,
                           early_stopping=early_stopping)
        return SearchOutput(sequences=out["sequences"], scores=out.get("scores"))

    @classmethod
    def from_clip_vision_mbart_pretrained(cls, clip_vision_model_name_or_path=None, mbart_model_name_or_path=None,
                                          *model_args, **kwargs):
        """:702-773 grafts two hub checkpoints; checkpoint I/O is outside the hot path and the
        hub is unreachable here.  Build from a config and assign `.params` instead."""
        raise NotImplementedError("checkpoint loading is out of scope; construct from a CLIPVisionMBartConfig and set "
                                  "`model.params = <tree with the Flax names>`")
