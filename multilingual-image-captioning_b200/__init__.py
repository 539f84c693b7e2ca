"""B200-native captioning hot path (CLIP-ViT -> mBART-50), drop-in for the reference's
`FlaxCLIPVisionMBartForConditionalGeneration` API.  Import as `import mic_b200`.

Only configuration / synthetic data import eagerly (usable on a CPU-only box); the model, engine and
training modules need CUDA + libmic_b200.so and are imported on first attribute access.
"""
import importlib

from .configuration import (CLIPVisionConfig, MBartConfig, CLIPVisionMBartConfig, clip_mbart_config,
                            vit_bart_config, tiny_config, tiny_vit_bart_config)
from . import synthetic

_LAZY = {
    "FlaxCLIPVisionMBartForConditionalGeneration": ".modeling_clip_vision_mbart",
    "Seq2SeqLMOutput": ".modeling_clip_vision_mbart",
    "FlaxViTBartForConditionalGeneration": ".modeling_vit_bart",
    "TrainState": ".training", "train_step": ".training", "eval_step": ".training",
    "create_learning_rate_fn": ".training", "adamw": ".training", "AdamWConfig": ".training",
    "save_model_checkpoint": ".training", "restore_model_checkpoint": ".training", "rotate_checkpoints": ".training",
    "Transform": ".transforms", "BatchTransform": ".transforms",
}
_LAZY_MODULES = ("ops", "engine", "generation", "training", "params", "modeling_clip_vision_mbart", "modeling_vit_bart",
                 "checkpoint", "transforms", "_lib")


def __getattr__(name):
    if name in _LAZY:
        return getattr(importlib.import_module(_LAZY[name], __name__), name)
    if name in _LAZY_MODULES:
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)


__all__ = ["CLIPVisionConfig", "MBartConfig", "CLIPVisionMBartConfig", "clip_mbart_config", "vit_bart_config",
           "tiny_config", "tiny_vit_bart_config", "synthetic"] + list(_LAZY)
