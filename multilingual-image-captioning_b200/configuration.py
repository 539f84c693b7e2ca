"""Hyper-parameter containers for the captioning hot path.

Mirrors the composite config of the reference
(`models/flax_clip_vision_mbart/configuration_clip_vision_mbart.py:10-51`,
`models/flax_vit_bart/configuration_vit_bart.py:10-43`): a vision config and a text-decoder
config held side by side, reachable as `config.clip_vision_config` / `config.mbart_config`.
Default values are those of `openai/clip-vit-base-patch32` and `facebook/mbart-large-50`
(SURVEY.md §8a-0).
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict


@dataclass
class CLIPVisionConfig:
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    image_size: int = 224
    patch_size: int = 32
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5
    initializer_range: float = 0.02
    # ViT (flax_vit_bart) differences: conv bias, exact gelu, final layernorm on the sequence
    patch_bias: bool = False
    pre_layernorm: bool = True      # CLIP `pre_layrnorm`
    final_layernorm: bool = False   # ViT `layernorm` applied to last_hidden_state
    channel_first_input: bool = False  # modeling_vit_bart.py:445 transposes NCHW -> NHWC
    # Normalize(mean, std) of the reference's Transform (main.py:174); applied in-kernel when uint8 pixels are handed in
    image_mean: tuple = (0.48145466, 0.4578275, 0.40821073)
    image_std: tuple = (0.26862954, 0.26130258, 0.27577711)

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def num_tokens(self) -> int:
        return self.num_patches + 1

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class MBartConfig:
    vocab_size: int = 250054
    d_model: int = 1024
    decoder_layers: int = 12
    decoder_attention_heads: int = 16
    decoder_ffn_dim: int = 4096
    activation_function: str = "gelu"
    dropout: float = 0.1
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    scale_embedding: bool = True
    max_position_embeddings: int = 1024
    position_offset: int = 2
    init_std: float = 0.02
    # Flax nn.LayerNorm default at the pinned transformers commit (SURVEY.md risk U1)
    layer_norm_eps: float = 1e-6
    pre_layernorm: bool = True       # mBART: pre-LN; BART: post-LN
    final_layer_norm: bool = True    # mBART only
    pad_token_id: int = 1
    bos_token_id: int = 0
    eos_token_id: int = 2
    decoder_start_token_id: int = 2
    forced_bos_token_id: int | None = None
    forced_eos_token_id: int | None = 2
    # generation defaults (generation_clip_vision_utils.py:196-229 reads them from here)
    max_length: int = 200
    min_length: int = 0
    num_beams: int = 5
    do_sample: bool = False
    early_stopping: bool = True
    length_penalty: float = 1.0
    no_repeat_ngram_size: int = 0

    @property
    def hidden_size(self) -> int:
        return self.d_model

    @property
    def head_dim(self) -> int:
        return self.d_model // self.decoder_attention_heads


@dataclass
class CLIPVisionMBartConfig:
    """Composite config (`configuration_clip_vision_mbart.py:10-51`)."""

    clip_vision_config: CLIPVisionConfig = field(default_factory=CLIPVisionConfig)
    mbart_config: MBartConfig = field(default_factory=MBartConfig)
    tie_word_embeddings: bool = True
    is_encoder_decoder: bool = True
    model_type: str = "clip-vision-mbart"
    output_attentions: bool = False
    output_hidden_states: bool = False
    return_dict: bool = True

    @classmethod
    def from_clip_vision_mbart_configs(cls, clip_vision_config, mbart_config, **kw):
        return cls(clip_vision_config=clip_vision_config, mbart_config=mbart_config, **kw)

    def to_dict(self):
        return asdict(self)

    # the ViT-BART variant exposes the same two sub-configs under other names
    @property
    def vit_config(self):
        return self.clip_vision_config

    @property
    def bart_config(self):
        return self.mbart_config


def clip_mbart_config(**mbart_overrides) -> CLIPVisionMBartConfig:
    """BASELINE configs 1-4: CLIP-ViT-B/32 + mBART-50 decoder."""
    return CLIPVisionMBartConfig(CLIPVisionConfig(), MBartConfig(**mbart_overrides))


def vit_bart_config() -> CLIPVisionMBartConfig:
    """BASELINE config 5: ViT-B/16 + BART-large decoder (`modeling_vit_bart.py`)."""
    v = CLIPVisionConfig(patch_size=16, hidden_act="gelu", layer_norm_eps=1e-12, patch_bias=True,
                         pre_layernorm=False, final_layernorm=True, channel_first_input=True)
    t = MBartConfig(vocab_size=50265, scale_embedding=False, pre_layernorm=False, final_layer_norm=False,
                    layer_norm_eps=1e-6, forced_eos_token_id=2, num_beams=4, max_length=20,
                    decoder_start_token_id=2)
    return CLIPVisionMBartConfig(v, t, model_type="vit-bart")


def tiny_config(vocab_size: int = 1003, layers: int = 2) -> CLIPVisionMBartConfig:
    """Small shapes with the same structure (head_dim 64, odd vocab) for fast parity tests."""
    v = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=layers,
                         num_attention_heads=2, image_size=64, patch_size=32)
    t = MBartConfig(vocab_size=vocab_size, d_model=128, decoder_layers=layers, decoder_attention_heads=2,
                    decoder_ffn_dim=256, max_position_embeddings=128)
    return CLIPVisionMBartConfig(v, t)


def tiny_vit_bart_config(vocab_size: int = 1003, layers: int = 2, image_size: int = 144) -> CLIPVisionMBartConfig:
    """flax_vit_bart structure at test size: 16x16 patches (81+1 = 82 visual tokens > 64 exercises the general
    attention kernels), conv bias, exact gelu, final ViT layernorm, post-LN BART decoder without final LN."""
    v = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=layers, num_attention_heads=2,
                         image_size=image_size, patch_size=16, hidden_act="gelu", layer_norm_eps=1e-12, patch_bias=True,
                         pre_layernorm=False, final_layernorm=True, channel_first_input=True)
    t = MBartConfig(vocab_size=vocab_size, d_model=128, decoder_layers=layers, decoder_attention_heads=2,
                    decoder_ffn_dim=256, max_position_embeddings=128, scale_embedding=False, pre_layernorm=False,
                    final_layer_norm=False, layer_norm_eps=1e-6)
    return CLIPVisionMBartConfig(v, t, model_type="vit-bart")
