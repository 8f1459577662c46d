"""Pipeline-level Prompt-to-Prompt drivers with the reference's class and method names (p2p/model/sd_utils.py):

    P2P          :9-87      SD-1.5, 512^2          text2image_ldm_stable / diffusion_step / init_latent / latent2image
    P2P_NTI      :89-139    + per-step null-text embeddings (`uncond_embeddings_list`)
    P2P_XL       :141-221   SDXL, 1024^2, added_cond_kwargs
    P2P_XL_NTI   :223-308   both

The reference spells the loop out four times; here the variants differ only in how the context is produced (`_conditioning`)
and whether it is re-assembled per step (`_context_for_step`). The attention inside every UNet forward runs through
register_attention_control's fused closures, and CFG + DDIM update is one ief_cfg_ddim_step launch.

`low_resource=True` is refused: in the reference that branch indexes the concatenated [2n, 77, C] context with [0] / [1] and
hands a 2-D tensor to the attention closure, which cannot unpack it — there is no working behaviour to reproduce.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import pipelines
from .register import register_attention_control


class P2P:
    height = width = 512

    def __init__(self, model, num_inference_steps, graphs: bool = False) -> None:
        """graphs=True (an extension, off by default): UNet forwards are replayed from CUDA graphs keyed by the controller's phase;
        keep the instance and the controller (controller.reset() between images) to amortise the captures."""
        model.scheduler.set_timesteps(num_inference_steps)
        self.graphs = graphs

    # ---- pieces the variants override ---------------------------------------------------------------------------------
    def _conditioning(self, model, prompt: List[str], guidance_scale: float, uncond_embeddings_list):
        """-> (context, extra UNet kwargs); context = [uncond * n, cond * n], or the n conditional rows when null-text embeddings
        will supply the unconditional half step by step."""
        uncond, cond = pipelines.text_context(model, prompt, with_uncond=uncond_embeddings_list is None)
        return (cond if uncond is None else torch.cat([uncond, cond])), {}

    def _context_for_step(self, context, i: int, uncond_embeddings_list):
        if uncond_embeddings_list is None:
            return context
        return torch.cat([uncond_embeddings_list[i].expand(*context.shape), context])   # context = the conditional rows only

    # ---- reference API -------------------------------------------------------------------------------------------------
    def init_latent(self, latent, model, height, width, generator, batch_size):
        return pipelines.start_latent(model, latent, height, width, generator, batch_size)

    @torch.no_grad()
    def text2image_ldm_stable(self, model, prompt: List[str], controller, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                              generator: Optional[torch.Generator] = None, latent: Optional[torch.Tensor] = None,
                              low_resource: bool = False, uncond_embeddings_list=None):
        if controller is not None:
            register_attention_control(model, controller)
        context, extra = self._conditioning(model, prompt, guidance_scale, uncond_embeddings_list)
        latent, latents = self.init_latent(latent, model, self.height, self.width, generator, len(prompt))
        model.scheduler.set_timesteps(num_inference_steps)
        for i, t in enumerate(model.scheduler.timesteps.tolist()):
            latents = self._step(model, controller, latents, self._context_for_step(context, i, uncond_embeddings_list), t,
                                 guidance_scale, extra, low_resource)
        return self.latent2image(model.vae, latents), latent

    def _step(self, model, controller, latents, context, t, guidance_scale, extra, low_resource):
        runner = pipelines.graph_runner(self, model, controller)
        if low_resource:
            raise NotImplementedError("low_resource=True: the reference's two-pass branch passes context[0] / context[1] (2-D rows of the "
                                      "concatenated context) to the UNet and fails inside its attention closure; nothing to reproduce")
        latents = pipelines.guided_step(model, latents, context, t, guidance_scale, extra, runner=runner)
        return controller.step_callback(latents) if controller is not None else latents

    def diffusion_step(self, model, controller, latents, context, t, guidance_scale, low_resource=False):
        return self._step(model, controller, latents, context, t, guidance_scale, {}, low_resource)

    @torch.no_grad()
    def latent2image(self, vae, latents):
        return pipelines.decode_latents(vae, latents)


class P2P_NTI(P2P):
    """Null-text inversion replay: step i's unconditional rows are `uncond_embeddings_list[i]` (inversion/nti.py's output)."""


class P2P_XL(P2P):
    height = width = 1024

    def _conditioning(self, model, prompt, guidance_scale, uncond_embeddings_list):
        context, added = self.encode_prompt_xl(model, prompt, model._execution_device, guidance_scale > 1.0, self.height, self.width,
                                               len(prompt))
        return context, {"added_cond_kwargs": added}

    def _context_for_step(self, context, i, uncond_embeddings_list):
        return context

    def diffusion_step(self, model, controller, latents, context, t, guidance_scale, added_cond_kwargs, low_resource=False):
        return self._step(model, controller, latents, context, t, guidance_scale, {"added_cond_kwargs": added_cond_kwargs}, low_resource)

    def encode_prompt_xl(self, model, prompt, device, do_classifier_free_guidance, height, width, batch_size):
        return pipelines.sdxl_conditioning(model, prompt, device, do_classifier_free_guidance, height, width, batch_size)


class P2P_XL_NTI(P2P_XL):
    """SDXL + null-text: the negative half of the prompt embeddings is overwritten in place each step (reference :252)."""

    def _context_for_step(self, context, i, uncond_embeddings_list):
        half = context.shape[0] // 2
        context[:half] = uncond_embeddings_list[i].expand(*context[half:].shape)
        return context
