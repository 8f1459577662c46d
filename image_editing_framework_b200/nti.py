"""Null-text inversion (reference: */inversion/nti.py, identical in the four method directories): per DDIM step, optimise the
unconditional text embedding with Adam so that the classifier-free-guided step from the current latent lands on the DDIM
inversion trajectory. The result, one [1, 77, C] embedding per step, is what the *_NTI pipeline classes replay.

    NTI       :7-46     SD-1.5; the embedding carries over from step to step, lr = 1e-2 (1 - i/100)
    NTI_XL    :48-96    SDXL; restarts from the negative prompt embedding each step, lr = lr (1 - i/500), added_cond_kwargs

Both are one search (`_search`) with different forwards. The UNet runs unhooked here (no controller is registered during
inversion), the inner loop needs d loss / d embedding through the UNet and therefore stays differentiable torch arithmetic; the
guided step that advances the latent after each search is the fused ief_cfg_ddim_step launch.
"""
from __future__ import annotations

from typing import Callable, List

import torch
import torch.nn.functional as F

from .ddim import FusedDDIM, ddim_inversion, ddim_inversion_xl


def _search(model, trajectory: List[torch.Tensor], guidance_scale: float, num_inner_steps: int, epsilon: float,
            start: torch.Tensor, carry_over: bool, lr_at: Callable[[int], float],
            eps_cond: Callable, eps_uncond: Callable, eps_pair: Callable) -> List[torch.Tensor]:
    """trajectory: the inversion's latents, x_0 first. eps_cond(x, t) / eps_uncond(x, t, emb) / eps_pair(x2, t, emb) are the three UNet
    forwards of the method (conditional, unconditional with the trainable embedding, both halves at once)."""
    fused = FusedDDIM(model.scheduler)
    steps = int(model.scheduler.num_inference_steps)
    timesteps = model.scheduler.timesteps.tolist()
    stride = fused._stride()
    found = []
    x = trajectory[-1]
    emb = start
    for i in range(steps):
        emb = (emb if carry_over else start).clone().detach().requires_grad_(True)
        adam = torch.optim.Adam([emb], lr=lr_at(i))
        target = trajectory[len(trajectory) - i - 2]
        t = timesteps[i]
        # 0-d fp32 host tensors, as scheduler.step reads them: the search is sensitive to the last bit (Adam normalises tiny gradients)
        ac = model.scheduler.alphas_cumprod
        a_t, a_prev = ac[t], (ac[t - stride] if t - stride >= 0 else model.scheduler.final_alpha_cumprod)
        with torch.no_grad():
            e_c = eps_cond(x, t)
        for _ in range(num_inner_steps):
            e_u = eps_uncond(x, t, emb)
            eps = e_u + guidance_scale * (e_c - e_u)
            x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5                  # scheduler.step, eta = 0, kept differentiable
            loss = F.mse_loss(a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps, target)
            adam.zero_grad()
            loss.backward()
            adam.step()
            if loss.item() < epsilon + i * 2e-5:
                break
        found.append(emb[:1].detach())
        with torch.no_grad():
            x = fused.step(eps_pair(torch.cat([x] * 2), t, emb.detach()), t, x, guidance_scale)
    return found


class NTI(ddim_inversion):
    def null_optimization(self, model, latents, context, num_inner_steps, epsilon, guidance_scale):
        uncond, cond = context.chunk(2)
        unet = model.unet
        return _search(model, latents, guidance_scale, num_inner_steps, epsilon, uncond, True, lambda i: 1e-2 * (1.0 - i / 100.0),
                       lambda x, t: unet(x, t, cond).sample,
                       lambda x, t, emb: unet(x, t, emb).sample,
                       lambda x2, t, emb: unet(x2, t, encoder_hidden_states=torch.cat([emb, cond]))["sample"])


class NTI_XL(ddim_inversion_xl):
    def null_optimization(self, model, latents, context, num_inner_steps, epsilon, guidance_scale, height=1024, width=1024, lr=0.5):
        cond, negative, pooled, negative_pooled = context
        size = (height, width)
        ids = model._get_add_time_ids(size, (0, 0), size, dtype=cond.dtype).to(model.unet.device)
        kw_c = {"text_embeds": pooled, "time_ids": ids}
        kw_u = {"text_embeds": negative_pooled.detach(), "time_ids": ids.detach()}
        kw_pair = {"text_embeds": torch.cat([negative_pooled, pooled]), "time_ids": torch.cat([ids, ids])}
        unet = model.unet
        return _search(model, latents, guidance_scale, num_inner_steps, epsilon, negative, False, lambda i: lr * (1.0 - i / 500.0),
                       lambda x, t: unet(x, t, cond, added_cond_kwargs=kw_c).sample,
                       lambda x, t, emb: unet(x, t, emb, added_cond_kwargs=kw_u).sample,
                       lambda x2, t, emb: unet(x2, t, encoder_hidden_states=torch.cat([emb, cond]), added_cond_kwargs=kw_pair)["sample"])
