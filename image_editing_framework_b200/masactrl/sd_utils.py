"""Pipeline-level MasaCtrl samplers with the reference's class names and call signatures (masactrl/model/sd_utils.py):

    MasaCtrl          :7-124     SD-1.5 DDIM sampler; `unconditioning` = per-step null-text embeddings, `ref_intermediate_latents`
                                 = the inversion trajectory whose entries replace the source half of the batch step by step
    MasaCtrl_XL       :127-225   SDXL (encode_prompt + added_cond_kwargs)
    MasaCtrl_NTI      :227-303   SD-1.5 with `uncond_embeddings_list`
    MasaCtrl_XL_NTI   :305-381   SDXL with `uncond_embeddings_list`

One loop serves all four; a variant only says how its context is built and how a step's context is assembled. The mutual
self-attention itself happens in the closures installed by regiter_attention_editor_diffusers (one fused launch per layer), and
the guidance + DDIM update of a step is one ief_cfg_ddim_step launch.
"""
from __future__ import annotations

import torch

from .. import pipelines


class MasaCtrl:
    def __init__(self, pipeline, num_inference_steps, graphs: bool = False) -> None:
        """graphs=True (an extension, off by default): UNet forwards are replayed from CUDA graphs keyed by the registered editor's
        phase; keep the instance and the editor (editor.reset() between images) to amortise the captures."""
        self.model = pipeline
        self.model.scheduler.set_timesteps(num_inference_steps)
        self.graphs = graphs

    @torch.no_grad()
    def latent2image(self, latents, return_type="np"):
        return pipelines.decode_latents(self.model.vae, latents, return_type)

    # ---- what the variants differ in -----------------------------------------------------------------------------------
    def _conditioning(self, prompt, batch_size, height, width, guidance_scale, neg_prompt, direction):
        """-> (context, extra UNet kwargs). SD-1.5: [uncond; cond] when guidance is on, the conditional rows alone otherwise."""
        uncond, cond = pipelines.text_context(self.model, prompt, negative=neg_prompt or "", with_uncond=guidance_scale > 1.0, truncation=False)
        if direction:
            # nudges the last prompt along the principal axis of (second-to-last - last), masactrl/model/sd_utils.py:56-61
            delta = cond[-2] - cond[-1]
            _, _, axis = torch.pca_lowrank(delta.transpose(-1, -2), q=1, center=True)
            cond[-1] = cond[-1] + direction * axis
        return (cond if uncond is None else torch.cat([uncond, cond])), {}

    def _context_for_step(self, context, i, null_text):
        if not isinstance(null_text, list):
            return context
        cond = context.chunk(2)[1]
        return torch.cat([null_text[i].expand(*cond.shape), cond])

    # ---- the sampler ------------------------------------------------------------------------------------------------------
    def _sample(self, prompt, batch_size, height, width, num_inference_steps, guidance_scale, latents, ref_intermediate_latents,
                null_text, neg_prompt=None, direction=None):
        model = self.model
        if isinstance(prompt, list):
            batch_size = len(prompt)
        elif batch_size > 1:
            prompt = [prompt] * batch_size
        context, extra = self._conditioning(prompt, batch_size, height, width, guidance_scale, neg_prompt, direction)
        shape = (batch_size, model.unet.config.in_channels, height // 8, width // 8)
        if latents is None:
            latents = torch.randn(shape, device=model.unet.device, dtype=model.unet.dtype)
        elif tuple(latents.shape) != shape:
            raise AssertionError(f"The shape of input latent tensor {tuple(latents.shape)} should equal to predefined one {shape}.")
        init_latent = latents.clone()
        model.scheduler.set_timesteps(num_inference_steps)
        runner = pipelines.graph_runner(self, model, getattr(model.unet, "_ief_installed", None))
        for i, t in enumerate(model.scheduler.timesteps.tolist()):
            if ref_intermediate_latents is not None:       # the source branch is re-seated on its inversion trajectory every step
                latents = torch.cat([ref_intermediate_latents[-1 - i], latents.chunk(2)[1]])
            latents = pipelines.guided_step(model, latents, self._context_for_step(context, i, null_text), t, guidance_scale, extra,
                                            always_guide=False, runner=runner)
        return self.latent2image(latents, return_type="np"), init_latent

    @torch.no_grad()
    def __call__(self, prompt, batch_size=1, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, latents=None,
                 unconditioning=None, neg_prompt=None, ref_intermediate_latents=None, return_intermediates=False, **kwds):
        return self._sample(prompt, batch_size, height, width, num_inference_steps, guidance_scale, latents, ref_intermediate_latents,
                            unconditioning, neg_prompt, kwds.get("dir"))


class MasaCtrl_NTI(MasaCtrl):
    def _conditioning(self, prompt, batch_size, height, width, guidance_scale, neg_prompt, direction):
        return super()._conditioning(prompt, batch_size, height, width, 0.0, neg_prompt, direction)   # the conditional rows only

    def _context_for_step(self, context, i, null_text):
        return torch.cat([null_text[i].expand(*context.shape), context])

    @torch.no_grad()
    def __call__(self, prompt, batch_size=1, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, latents=None,
                 unconditioning=None, neg_prompt=None, ref_intermediate_latents=None, return_intermediates=False,
                 uncond_embeddings_list=None, **kwds):
        return self._sample(prompt, batch_size, height, width, num_inference_steps, guidance_scale, latents, ref_intermediate_latents,
                            uncond_embeddings_list, neg_prompt, kwds.get("dir"))


class MasaCtrl_XL(MasaCtrl):
    def _conditioning(self, prompt, batch_size, height, width, guidance_scale, neg_prompt, direction):
        context, added = self.encode_prompt_xl(prompt, self.model.unet.device, guidance_scale > 1.0, height, width, batch_size)
        return context, {"added_cond_kwargs": added}

    def _context_for_step(self, context, i, null_text):
        return context

    def encode_prompt_xl(self, prompt, device, do_classifier_free_guidance, height, width, batch_size):
        return pipelines.sdxl_conditioning(self.model, prompt, device, do_classifier_free_guidance, height, width, batch_size)

    @torch.no_grad()
    def __call__(self, prompt, batch_size=1, height=1024, width=1024, num_inference_steps=50, guidance_scale=7.5, latents=None,
                 ref_intermediate_latents=None, return_intermediates=False):
        return self._sample(prompt, batch_size, height, width, num_inference_steps, guidance_scale, latents, ref_intermediate_latents, None)


class MasaCtrl_XL_NTI(MasaCtrl_XL):
    def _context_for_step(self, context, i, null_text):
        half = context.shape[0] // 2
        context[:half] = null_text[i].expand(*context[half:].shape)     # the negative rows, in place (reference :363)
        return context

    @torch.no_grad()
    def __call__(self, prompt, batch_size=1, height=1024, width=1024, num_inference_steps=50, guidance_scale=7.5, latents=None,
                 ref_intermediate_latents=None, return_intermediates=False, uncond_embeddings_list=None):
        return self._sample(prompt, batch_size, height, width, num_inference_steps, guidance_scale, latents, ref_intermediate_latents,
                            uncond_embeddings_list)
