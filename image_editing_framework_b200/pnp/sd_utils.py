"""Pipeline-level Plug-and-Play samplers with the reference's class names and call signatures (pnp/model/sd_utils.py):

    PnP          :11-128    SD-1.5; q/k injection for the first int(steps * pnp_attn_t) timesteps, resnet-feature injection for the
                            first int(steps * pnp_f_t)
    PnP_XL       :130-259   SDXL hooks + added_cond_kwargs
    PnP_NTI      :261-358   per-step null-text embeddings written into the negative rows of the context
    PnP_XL_NTI   :360-448   both

The reference carries the diffusers `__call__` signature through all four; the arguments it never reads (output_type, return_dict,
callback, callback_steps, guidance_rescale) are accepted and ignored here too. The injections run in the hooks of pnp/register.py
(row-source tables of one fused attention launch, one batch-row copy per resnet), the step update is ief_cfg_ddim_step.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from .. import pipelines
from . import register as hooks


class PnP:
    _register_attn = staticmethod(hooks.register_attention_control_efficient)
    _register_conv = staticmethod(hooks.register_conv_control_efficient)
    _register_time = staticmethod(hooks.register_time)
    _unregister_attn = staticmethod(hooks.unregister_attention_control_efficient)
    _unregister_conv = staticmethod(hooks.unregister_conv_control_efficient)

    def __init__(self, pipeline, num_inference_steps, graphs: bool = False) -> None:
        """graphs=True (an extension, off by default): UNet forwards are replayed from CUDA graphs keyed by which injections the
        timestep switches on; keep the instance to amortise the captures over edits."""
        self.model = pipeline
        self.model.scheduler.set_timesteps(num_inference_steps)
        self.graphs = graphs

    def init_pnp(self, conv_injection_t, qk_injection_t):
        ts = self.model.scheduler.timesteps
        self.qk_injection_timesteps = ts[:qk_injection_t] if qk_injection_t >= 0 else []
        self.conv_injection_timesteps = ts[:conv_injection_t] if conv_injection_t >= 0 else []
        self._register_attn(self.model, self.qk_injection_timesteps)
        self._register_conv(self.model, self.conv_injection_timesteps)

    # ---- what the variants differ in -----------------------------------------------------------------------------------
    def _conditioning(self, prompt, device, guided, negative_prompt, num_images_per_prompt, prompt_embeds, negative_prompt_embeds,
                      lora_scale, height, width, batch_size):
        pos, neg = self.model.encode_prompt(prompt, device, num_images_per_prompt, guided, negative_prompt, prompt_embeds=prompt_embeds,
                                            negative_prompt_embeds=negative_prompt_embeds, lora_scale=lora_scale)
        return (torch.cat([neg, pos]) if guided else pos), {}

    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str]] = None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt: Optional[Union[str, List[str]]] = None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil", return_dict: bool = True, callback: Optional[Callable] = None,
                 callback_steps: int = 1, cross_attention_kwargs: Optional[Dict[str, Any]] = None, guidance_rescale: float = 0.0,
                 pnp_attn_t=0.5, pnp_f_t=0.8, uncond_embeddings_list=None):
        model = self.model
        if eta != 0.0:
            raise NotImplementedError("the fused DDIM update is deterministic (eta = 0), which is what the reference's scripts use")
        device = model._execution_device
        model.scheduler.set_timesteps(num_inference_steps, device=device)
        side = model.unet.config.sample_size * model.vae_scale_factor
        height, width = height or side, width or side
        batch_size = 1 if isinstance(prompt, str) else len(prompt) if isinstance(prompt, list) else prompt_embeds.shape[0]
        guided = guidance_scale > 1.0
        lora_scale = cross_attention_kwargs.get("scale", None) if cross_attention_kwargs is not None else None
        context, extra = self._conditioning(prompt, device, guided, negative_prompt, num_images_per_prompt, prompt_embeds,
                                            negative_prompt_embeds, lora_scale, height, width, batch_size)
        channels = model.unet.config.in_channels
        latents = model.prepare_latents(num_images_per_prompt, channels, height, width, context.dtype, device, generator, latents)
        latents = latents.expand(batch_size, channels, height // 8, width // 8)
        self.init_pnp(conv_injection_t=int(num_inference_steps * pnp_f_t), qk_injection_t=int(num_inference_steps * pnp_attn_t))
        qk_on = frozenset(int(t) for t in self.qk_injection_timesteps)
        conv_on = frozenset(int(t) for t in self.conv_injection_timesteps)
        runner = pipelines.graph_runner(self, model, None, lambda t: (t in qk_on or t == 1000, t in conv_on or t == 1000))
        try:
            for i, t in enumerate(model.scheduler.timesteps.tolist()):
                self._register_time(model, t)
                if uncond_embeddings_list is not None:          # null-text embeddings replace the negative rows in place
                    half = context.shape[0] // 2
                    context[:half] = uncond_embeddings_list[i].expand(*context[half:].shape)
                latents = pipelines.guided_step(model, latents, context, t, guidance_scale,
                                                dict(extra, cross_attention_kwargs=cross_attention_kwargs), always_guide=False,
                                                runner=runner)
            return self.latent2image(latents)
        finally:
            self._unregister_attn(model)
            self._unregister_conv(model)

    @torch.no_grad()
    def latent2image(self, latents, return_type="np"):
        return pipelines.decode_latents(self.model.vae, latents, return_type)


class PnP_NTI(PnP):
    """`uncond_embeddings_list[i]` (null-text inversion's output) is step i's unconditional embedding."""


class PnP_XL(PnP):
    _register_attn = staticmethod(hooks.register_attention_control_efficient_xl)
    _register_conv = staticmethod(hooks.register_conv_control_efficient_xl)
    _register_time = staticmethod(hooks.register_time_xl)
    _unregister_attn = staticmethod(hooks.unregister_attention_control_efficient_xl)
    _unregister_conv = staticmethod(hooks.unregister_conv_control_efficient_xl)

    def _conditioning(self, prompt, device, guided, negative_prompt, num_images_per_prompt, prompt_embeds, negative_prompt_embeds,
                      lora_scale, height, width, batch_size):
        context, added = self.encode_prompt_xl(prompt, device, guided, height, width, batch_size)
        return context, {"added_cond_kwargs": added}

    def encode_prompt_xl(self, prompt, device, do_classifier_free_guidance, height, width, batch_size):
        return pipelines.sdxl_conditioning(self.model, prompt, device, do_classifier_free_guidance, height, width, batch_size)


class PnP_XL_NTI(PnP_XL):
    """SDXL hooks + null-text embeddings."""
