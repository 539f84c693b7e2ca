"""B200-native captioning hot path (CLIP-ViT -> mBART-50), drop-in for the reference's
`FlaxCLIPVisionMBartForConditionalGeneration` API.  Import as `import mic_b200`."""
from .configuration import (CLIPVisionConfig, MBartConfig, CLIPVisionMBartConfig, clip_mbart_config,
                            vit_bart_config, tiny_config)
from . import synthetic

__all__ = ["CLIPVisionConfig", "MBartConfig", "CLIPVisionMBartConfig", "clip_mbart_config",
           "vit_bart_config", "tiny_config", "synthetic"]
