"""Shared pieces of the pipeline-level drivers that mirror the reference's `*/model/sd_utils.py` classes.

Every method directory of the reference carries its own copy of the same three things around its 50-step loop: the text
conditioning ([uncond * n, cond * n], or SDXL's encode_prompt + pooled embedding + size ids), a classifier-free-guidance UNet
forward followed by `scheduler.step`, and the VAE decode to uint8. Here they exist once; the step update is the fused
ief_cfg_ddim_step launch (ddim.FusedDDIM) instead of the scheduler's ~10 elementwise kernels.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _cabi
from .ddim import FusedDDIM
from .graphs import GraphedUNet


def text_context(model, prompts: List[str], negative: str = "", with_uncond: bool = True, truncation: bool = True):
    """(uncond [n, 77, C] or None, cond [n, 77, C]) from the pipeline's tokenizer + text encoder."""
    tk = model.tokenizer
    extra = {"truncation": True} if truncation else {}
    ids = tk(prompts, padding="max_length", max_length=tk.model_max_length, return_tensors="pt", **extra).input_ids
    device = model.unet.device
    cond = model.text_encoder(ids.to(device))[0]
    uncond = None
    if with_uncond:
        neg = tk([negative] * len(prompts), padding="max_length", max_length=ids.shape[-1], return_tensors="pt").input_ids
        uncond = model.text_encoder(neg.to(device))[0]
    return uncond, cond


def sdxl_conditioning(model, prompt, device, do_classifier_free_guidance: bool, height: int, width: int,
                      batch_size: int) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """SDXL context + added_cond_kwargs as the reference's `encode_prompt_xl` builds them (p2p/model/sd_utils.py:184-221):
    [negative; positive] prompt embeddings, the pooled embeddings the same way, and one (size, crop, target) id row per UNet row."""
    pos, neg, pooled, neg_pooled = model.encode_prompt(
        prompt=prompt, prompt_2=None, device=device, num_images_per_prompt=1, do_classifier_free_guidance=do_classifier_free_guidance,
        negative_prompt=None, negative_prompt_2=None, prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
        negative_pooled_prompt_embeds=None, lora_scale=None)
    size = (height, width)
    ids = model._get_add_time_ids(size, (0, 0), size, dtype=pos.dtype)
    if do_classifier_free_guidance:
        pos, pooled, ids = torch.cat([neg, pos]), torch.cat([neg_pooled, pooled]), torch.cat([ids, ids])
    return pos.to(device), {"text_embeds": pooled.to(device), "time_ids": ids.to(device).repeat(batch_size, 1)}


def fused_scheduler(model) -> FusedDDIM:
    """One FusedDDIM per (pipeline, scheduler): the alpha table is read to the host once."""
    fused = getattr(model, "_ief_fused_ddim", None)
    if fused is None or fused.scheduler is not model.scheduler:
        fused = model._ief_fused_ddim = FusedDDIM(model.scheduler)
    return fused


def graph_runner(owner, model, controller, key_fn=None) -> Optional[GraphedUNet]:
    """The CUDA-graph runner of a pipeline-class instance built with graphs=True (None otherwise). It lives as long as the instance and
    is rebuilt when the UNet or the installed controller / editor changes: captured kernels read that controller's device tables, and
    its graph_key() names the phase a step replays (graphs.GraphedUNet). Reuse pays when one instance + one controller (reset()
    between images) serve many edits; a one-off call roughly breaks even (each phase is captured on its second occurrence)."""
    if not getattr(owner, "graphs", False):
        return None
    runner = getattr(owner, "_runner", None)
    if runner is None or runner.unet is not model.unet or runner.controller is not controller:
        if runner is not None:
            runner.close()
        runner = owner._runner = GraphedUNet(model.unet, controller, key_fn, launch_counter=_cabi.launch_count)
    runner.key_fn = key_fn
    return runner


def unet_eps(model, x: torch.Tensor, t, context: torch.Tensor, unet_kwargs: Optional[dict] = None,
             runner: Optional[GraphedUNet] = None) -> torch.Tensor:
    """One noise prediction; replayed from the runner's graph when there is one and the only extra UNet argument in play is SDXL's
    added_cond_kwargs (tensor inputs the runner copies like x and the context)."""
    kw = unet_kwargs or {}
    if runner is not None and all(v is None for k, v in kw.items() if k != "added_cond_kwargs"):
        return runner(x, t, context, kw.get("added_cond_kwargs"))
    return model.unet(x, t, encoder_hidden_states=context, **kw)["sample"]


def guided_step(model, latents: torch.Tensor, context: torch.Tensor, t, guidance_scale: float, unet_kwargs: Optional[dict] = None,
                always_guide: bool = True, runner: Optional[GraphedUNet] = None) -> torch.Tensor:
    """unet(cat[latents]*2) -> uncond + g (cond - uncond) -> DDIM step, the last two in one kernel launch.
    `always_guide=False` follows the drivers that skip guidance when guidance_scale <= 1 (single forward, plain step)."""
    fused = fused_scheduler(model)
    t = int(t)
    if always_guide or guidance_scale > 1.0:
        return fused.step(unet_eps(model, torch.cat([latents] * 2), t, context, unet_kwargs, runner), t, latents, guidance_scale)
    return fused.step(unet_eps(model, latents, t, context, unet_kwargs, runner), t, latents, None)


@torch.no_grad()
def decode_latents(vae, latents: torch.Tensor, return_type: str = "np"):
    """VAE decode -> [0, 1] -> uint8 NHWC numpy (`latent2image` of every reference driver); 'pt' keeps the [0, 1] tensor."""
    image = vae.decode(latents.detach() / vae.config.scaling_factor)["sample"]
    image = (image / 2 + 0.5).clamp(0, 1)
    if return_type == "pt":
        return image
    return (image.float().cpu().permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)     # .float(): numpy has no bf16


def start_latent(model, latent: Optional[torch.Tensor], height: int, width: int, generator, batch_size: int):
    """(x_T [1, C, h/8, w/8], x_T expanded over the prompt batch); a missing x_T is drawn on the host generator like randn_tensor."""
    c = model.unet.config.in_channels
    if latent is None:
        latent = torch.randn((1, c, height // 8, width // 8), generator=generator, dtype=model.unet.dtype).to(model.unet.device)
    latent = latent * model.scheduler.init_noise_sigma
    return latent, latent.expand(batch_size, c, height // 8, width // 8)
