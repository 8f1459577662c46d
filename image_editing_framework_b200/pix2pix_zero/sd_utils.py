"""Pipeline-level Pix2Pix-zero editors with the reference's class names and call signatures (pix2pix-zero/model/sd_utils.py):

    P2P_Zero          :6-210     SD-1.5: loop 1 samples with prompt[0] and records the cross-attention maps, loop 2 samples with
                                 prompt[1] (+ edit_dir) while pulling its maps towards the recorded ones (one SGD step on the latents per
                                 timestep); returns (reconstruction, edit) uint8 images, or the reconstruction alone with only_sample
    P2P_Zero_XL       :212-424   SDXL (encode_prompt_xl per prompt, added_cond_kwargs on every forward)
    P2P_Zero_NTI      :426-617   row 0 of both contexts is overwritten with uncond_embeddings_list[i] at step i
    P2P_Zero_XL_NTI   :619-783   both

The two loops themselves are editing.pix2pix_zero_edit (maps cached on the device, fused kernels on the no-grad forwards, the
cross-attention backward kernel on the guidance pass); these classes add the text conditioning, x_T and the VAE decode.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import torch

from .. import editing, pipelines
from .attention_control import prep_unet


class P2P_Zero:
    def __init__(self, pipeline, num_inference_steps, graphs: bool = False):
        """graphs=True (an extension, off by default): the no-grad forwards of both loops are replayed from a CUDA graph kept on this
        instance, so keep the instance to amortise the capture over edits."""
        self.model = pipeline
        self.model.scheduler.set_timesteps(num_inference_steps)
        self.graphs = graphs

    # ---- what the variants differ in -----------------------------------------------------------------------------------
    def _conditioning(self, text, device, guided, num_images_per_prompt, negative_prompt, prompt_embeds, negative_prompt_embeds,
                      lora_scale, height, width):
        """-> ([uncond, cond] context of one prompt, extra UNet kwargs)."""
        pos, neg = self.model.encode_prompt(text, device, num_images_per_prompt, guided, negative_prompt, prompt_embeds=prompt_embeds,
                                            negative_prompt_embeds=negative_prompt_embeds, lora_scale=lora_scale)
        return torch.cat([neg, pos]), {}

    def __call__(self, prompt: Union[str, List[str]] = None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt: Optional[Union[str, List[str]]] = None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 cross_attention_kwargs: Optional[Dict[str, Any]] = None, guidance_amount=0.1, edit_dir=None, only_sample=False,
                 uncond_embeddings_list=None):
        model = self.model
        if eta != 0.0:
            raise NotImplementedError("the fused DDIM update is deterministic (eta = 0), which is what the reference's scripts use")
        if guidance_scale <= 1.0:
            raise NotImplementedError("pix2pix-zero without classifier-free guidance: the fused loops always run the [uncond, cond] pair")
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs are not forwarded (the reference's scripts pass None)")
        model.unet, self.original_processors = prep_unet(model.unet)
        side = model.unet.config.sample_size * model.vae_scale_factor
        height, width = height or side, width or side
        device = model._execution_device
        with torch.no_grad():
            src, kw_src = self._conditioning(prompt[0], device, True, num_images_per_prompt, negative_prompt, prompt_embeds,
                                             negative_prompt_embeds, None, height, width)
            edit, kw_edit = src, kw_src
            if not only_sample:
                edit, kw_edit = self._conditioning(prompt[1], device, True, num_images_per_prompt, None, None, None, None, height, width)
                edit = edit.clone()
                if edit_dir is not None:
                    edit += edit_dir
            model.scheduler.set_timesteps(num_inference_steps, device=device)
            latents = model.prepare_latents(num_images_per_prompt, model.unet.config.in_channels, height, width, src.dtype, device,
                                            generator, latents)
        rec, edited = editing.pix2pix_zero_edit(model, src, edit, latents, num_inference_steps, guidance_scale, guidance_amount,
                                                only_sample=only_sample,
                                                graphs=(pipelines.graph_runner(self, model, None) if not (kw_src or kw_edit) else None) or False,
                                                uncond_embeddings_list=uncond_embeddings_list,
                                                unet_kwargs_src=kw_src, unet_kwargs_edit=kw_edit)
        image_rec = self.latent2image(rec)
        return image_rec if only_sample else (image_rec, self.latent2image(edited))

    def latent2image(self, latents, return_type="np"):
        return pipelines.decode_latents(self.model.vae, latents, return_type)


class P2P_Zero_NTI(P2P_Zero):
    """`uncond_embeddings_list[i]` (null-text inversion's output) is step i's unconditional embedding in both loops."""


class P2P_Zero_XL(P2P_Zero):
    def _conditioning(self, text, device, guided, num_images_per_prompt, negative_prompt, prompt_embeds, negative_prompt_embeds,
                      lora_scale, height, width):
        context, added = self.encode_prompt_xl(text, device, guided, height, width, 1)
        return context, {"added_cond_kwargs": added}

    def encode_prompt_xl(self, prompt, device, do_classifier_free_guidance, height, width, batch_size):
        return pipelines.sdxl_conditioning(self.model, prompt, device, do_classifier_free_guidance, height, width, batch_size)


class P2P_Zero_XL_NTI(P2P_Zero_XL):
    """SDXL conditioning + null-text embeddings."""
