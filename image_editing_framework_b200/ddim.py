"""Fused CFG + DDIM update and DDIM inversion step (reference: p2p/model/sd_utils.py:75-76 and the same two lines in
every driver; */inversion/ddim.py:9-18).

The reference issues ~10 tiny elementwise launches per step and indexes `alphas_cumprod` with tensors; here a step is
ONE kernel (ief_cfg_ddim_step) and the alpha lookups are host floats cached per scheduler (no device sync).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ops


class FusedDDIM:
    """Wraps a diffusers-style DDIM scheduler (alphas_cumprod, final_alpha_cumprod, config.num_train_timesteps,
    num_inference_steps). eta = 0, epsilon prediction, no sample clipping — the configuration the reference's scripts
    build (p2p/edit_real.py:58-69)."""

    def __init__(self, scheduler):
        # scheduler.step honours these switches; the fused kernel implements exactly one setting of each, so anything else is refused
        # rather than silently computed differently (SD-2.1-768 checkpoints ship v_prediction, a default-constructed diffusers
        # DDIMScheduler has clip_sample=True)
        cfg = scheduler.config
        get = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
        unsupported = [f"{k}={get(k)!r}" for k, ok in (("prediction_type", ("epsilon", None)), ("clip_sample", (False, None)),
                                                       ("thresholding", (False, None))) if get(k) not in ok]
        if unsupported:
            raise ValueError("the fused CFG + DDIM step implements epsilon prediction without sample clipping / thresholding (the scheduler "
                             "the reference's scripts build, p2p/edit_real.py:58-69); this scheduler has " + ", ".join(unsupported))
        self.scheduler = scheduler
        self._alphas: List[float] = [float(a) for a in scheduler.alphas_cumprod.tolist()]
        self._final = float(scheduler.final_alpha_cumprod)
        self._train_steps = int(scheduler.config.num_train_timesteps)

    def _stride(self) -> int:
        return self._train_steps // int(self.scheduler.num_inference_steps)

    def _alpha(self, t: int) -> float:
        return self._alphas[t] if t >= 0 else self._final

    def step(self, noise_pred: torch.Tensor, t, latents: torch.Tensor, guidance_scale: Optional[float] = None) -> torch.Tensor:
        """`scheduler.step(eps_u + g (eps_c - eps_u), t, latents).prev_sample` in one launch.

        noise_pred: the UNet output for torch.cat([latents] * 2) when guidance_scale is given (uncond half first),
        else the already-combined prediction.
        """
        t = int(t)
        a_t, a_prev = self._alpha(t), self._alpha(t - self._stride())
        x = latents.contiguous()
        if guidance_scale is None:
            return ops.cfg_ddim_step(noise_pred.contiguous(), None, x, 0.0, a_t, a_prev)
        eu, ec = noise_pred.contiguous().chunk(2)
        return ops.cfg_ddim_step(eu, ec, x, guidance_scale, a_t, a_prev)

    def reverse_step(self, model_output: torch.Tensor, t, sample: torch.Tensor) -> torch.Tensor:
        """ddim_reverse (inversion/ddim.py:9-18): from timestep min(T-1, t - stride) up to t."""
        nxt = int(t)
        cur = min(self._train_steps - 1, nxt - self._stride())
        return ops.cfg_ddim_step(model_output.contiguous(), None, sample.contiguous(), 0.0, self._alpha(cur), self._alphas[nxt])


class ddim_inversion:
    """Same method names as the reference class (*/inversion/ddim.py:7-58); only the arithmetic moved to the kernel.

    `graphs` (an extension, off by default; set it on the class or an instance): the inversion's UNet forwards are replayed from a
    CUDA graph kept on the UNet (one runner per installed controller; a controller whose graph_key() is None runs eagerly)."""

    graphs = False

    def ddim_reverse(self, model, model_output, timestep, sample):
        fused = getattr(model, "_ief_fused_ddim", None)
        if fused is None or fused.scheduler is not model.scheduler:
            fused = model._ief_fused_ddim = FusedDDIM(model.scheduler)
        return fused.reverse_step(model_output, timestep.item() if torch.is_tensor(timestep) else timestep, sample)

    def get_context(self, model, prompt):
        batch = len(prompt)
        tok = model.tokenizer(prompt, padding="max_length", max_length=model.tokenizer.model_max_length, truncation=True, return_tensors="pt")
        cond = model.text_encoder(tok.input_ids.to(model.device))[0]
        un = model.tokenizer([""] * batch, padding="max_length", max_length=tok.input_ids.shape[-1], return_tensors="pt")
        uncond = model.text_encoder(un.input_ids.to(model.device))[0]
        return torch.cat([uncond, cond])

    def _runner(self, model, latent, unet_kwargs):
        """The CUDA-graph runner of an inversion (graphs=True only): kept on the UNet, one per installed controller. With no
        controller registered the UNet's own attention is replayed; with one (e.g. a do-nothing masactrl.AttentionBase, which
        routes the inversion's attention through the fused kernels) its graph_key() names the phase like in the edit loops."""
        extras = [v for k, v in unet_kwargs.items() if k != "encoder_hidden_states"]
        if not (self.graphs and latent.is_cuda and all(v is None for v in extras)):
            return None
        from . import _cabi
        from .graphs import GraphedUNet
        installed = getattr(model.unet, "_ief_installed", None)
        runners = model.unet.__dict__.setdefault("_ief_inversion_runners", {})
        runner = runners.get(id(installed))
        if runner is None or runner.unet is not model.unet or runner.controller is not installed:
            if len(runners) >= 4:       # bounded: each runner pins a private graph memory pool
                runners.pop(next(iter(runners))).close()
            runner = runners[id(installed)] = GraphedUNet(model.unet, installed, launch_counter=_cabi.launch_count)
        return runner

    def _invert(self, model, latent, unet_kwargs):
        """The loop both classes share: the scheduler's timesteps walked backwards, one UNet forward + one fused reverse step each."""
        trajectory = [latent]
        latent = latent.clone().detach()
        runner = self._runner(model, latent, unet_kwargs)
        for t in reversed(model.scheduler.timesteps.tolist()):   # one host copy instead of a device sync per step
            noise_pred = runner(latent, t, unet_kwargs["encoder_hidden_states"]) if runner is not None else \
                model.unet(latent, t, **unet_kwargs).sample
            latent = self.ddim_reverse(model, noise_pred, t, latent)
            trajectory.append(latent)
        return trajectory

    @torch.no_grad()
    def ddim_inversion_loop(self, model, latent, prompt, cross_attention_kwargs=None):
        context = self.get_context(model, prompt)
        _, cond = context.chunk(2)
        return self._invert(model, latent, dict(encoder_hidden_states=cond, cross_attention_kwargs=cross_attention_kwargs)), context

    @torch.no_grad()
    def image2latent(self, model, image, device, dtype):
        """uint8 HWC image -> scaled VAE latent mean (*/inversion/ddim.py:35-41)."""
        pixels = torch.as_tensor(np.asarray(image)).to(dtype).div(127.5).sub(1.0)
        latents = model.vae.encode(pixels.permute(2, 0, 1)[None].to(device))["latent_dist"].mean
        return latents * model.vae.config.scaling_factor


class ddim_inversion_xl(ddim_inversion):
    """SDXL flavour (*/inversion/ddim.py:60-109): the context is encode_prompt's 4-tuple and every forward carries the pooled text
    embedding + size/crop ids as added_cond_kwargs. Only the conditional half is inverted, as in the SD-1.5 class."""

    def get_context(self, model, prompt):
        return tuple(model.encode_prompt(prompt=prompt, prompt_2=None, device=model.unet.device, num_images_per_prompt=1,
                                         do_classifier_free_guidance=True, negative_prompt=None, negative_prompt_2=None,
                                         prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                                         negative_pooled_prompt_embeds=None, lora_scale=None))

    @torch.no_grad()
    def ddim_inversion_loop(self, model, latent, prompt, cross_attention_kwargs=None, height=1024, width=1024):
        context = self.get_context(model, prompt)
        prompt_embeds, _, pooled, _ = context
        device = model._execution_device
        size = (height, width)
        time_ids = model._get_add_time_ids(size, (0, 0), size, dtype=prompt_embeds.dtype).to(device)
        extra = {"text_embeds": pooled.to(device), "time_ids": time_ids}
        return self._invert(model, latent, dict(encoder_hidden_states=prompt_embeds.to(device), cross_attention_kwargs=cross_attention_kwargs,
                                                added_cond_kwargs=extra)), context
