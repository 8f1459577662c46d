"""diffusers / CLIP stand-ins (test and bench infrastructure; diffusers is not installable offline)."""
from .unet import (Attention, AttnProcessor, BasicTransformerBlock, Transformer2DModel, ResnetBlock2D, Upsample2D, Downsample2D,
                   UNet2DConditionModel, UNetConfig, sd15_config, sd21_config, sdxl_config, tiny_config, attention_geometry)
from .pipeline import DDIMScheduler, WordPieceTokenizer, TextEncoder, VAE, StableDiffusionPipeline, make_pipeline
