"""Fused CFG + DDIM update and DDIM inversion step (reference: p2p/model/sd_utils.py:75-76 and the same two lines in
every driver; */inversion/ddim.py:9-18).

The reference issues ~10 tiny elementwise launches per step and indexes `alphas_cumprod` with tensors; here a step is
ONE kernel (ief_cfg_ddim_step) and the alpha lookups are host floats cached per scheduler (no device sync).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops


class FusedDDIM:
    """Wraps a diffusers-style DDIM scheduler (alphas_cumprod, final_alpha_cumprod, config.num_train_timesteps,
    num_inference_steps). eta = 0, epsilon prediction, no sample clipping — the configuration the reference's scripts
    build (p2p/edit_real.py:58-69)."""

    def __init__(self, scheduler):
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
    """Same method names as the reference class (*/inversion/ddim.py:7-58); only the arithmetic moved to the kernel."""

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

    @torch.no_grad()
    def ddim_inversion_loop(self, model, latent, prompt, cross_attention_kwargs=None):
        context = self.get_context(model, prompt)
        _, cond = context.chunk(2)
        all_latent = [latent]
        latent = latent.clone().detach()
        steps = model.scheduler.timesteps.tolist()  # one host copy instead of a sync per step
        for t in reversed(steps):
            noise_pred = model.unet(latent, t, encoder_hidden_states=cond).sample
            latent = self.ddim_reverse(model, noise_pred, t, latent)
            all_latent.append(latent)
        return all_latent, context
