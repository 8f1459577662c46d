"""Compact L4 drivers (the 50-step denoise loops of the reference's `*/model/sd_utils.py`) wired to the fused step.

These mirror the reference call sequences so end-to-end parity tests and bench.py can run a whole edit:
  p2p_edit       p2p/model/sd_utils.py:24-79        (register -> loop: unet(cat[latents]*2) -> CFG -> scheduler.step -> step_callback)
  masactrl_edit  masactrl/model/sd_utils.py:25-124
  pnp_edit       pnp/model/sd_utils.py:23-115       (register_time per step, q/k + feature injection schedules)
  pix2pix_zero_edit  pix2pix-zero/model/sd_utils.py:86-182  (map-collection loop, then per step: guidance gradient on the latents
                                                     from the cross-attention maps -> SGD step -> recomputed noise -> DDIM step)
Orchestration only — every attention call goes through the registered closures, every step update through FusedDDIM.
"""
from __future__ import annotations

from typing import List, Optional, Union

import torch

from . import _cabi
from .ddim import FusedDDIM
from .graphs import GraphedUNet
from .p2p.register import register_attention_control
from .pnp import register as pnp_register


def encode_prompts(model, prompts: List[str]) -> torch.Tensor:
    """[uncond * n, cond * n] context, as every reference driver builds it."""
    tok = model.tokenizer(prompts, padding="max_length", max_length=model.tokenizer.model_max_length, truncation=True, return_tensors="pt")
    cond = model.text_encoder(tok.input_ids.to(model.device))[0]
    un = model.tokenizer([""] * len(prompts), padding="max_length", max_length=tok.input_ids.shape[-1], return_tensors="pt")
    uncond = model.text_encoder(un.input_ids.to(model.device))[0]
    return torch.cat([uncond, cond])


@torch.no_grad()
def denoise(model, latents: torch.Tensor, context: torch.Tensor, num_inference_steps: int, guidance_scale: float,
            step_callback=None, per_step=None, graphs: Union[bool, GraphedUNet] = False, controller=None, graph_key_fn=None,
            stats: Optional[dict] = None) -> torch.Tensor:
    """The 50-step loop. With `graphs` each control phase of the UNet forward is replayed from a CUDA graph
    (graphs.GraphedUNet): `controller` is the installed controller / editor (its graph_key() names the phase),
    `graph_key_fn(t)` adds timestep-dependent host state (the PnP schedules). `graphs=True` builds a runner for this call
    only (captures included); pass a GraphedUNet built once for the same unet + controller to amortise them over edits."""
    model.scheduler.set_timesteps(num_inference_steps)
    fused = FusedDDIM(model.scheduler)
    own = graphs is True
    runner = GraphedUNet(model.unet, controller, graph_key_fn, launch_counter=_cabi.launch_count) if own else (graphs or None)
    if runner is not None and not own:
        if runner.controller is not controller:
            raise ValueError("the GraphedUNet passed as `graphs` was built for a different controller")
        if runner.key_fn is None:
            runner.key_fn = graph_key_fn
    try:
        for t in model.scheduler.timesteps.tolist():
            if per_step is not None:
                per_step(t)
            x = torch.cat([latents] * 2)
            noise_pred = runner(x, t, context) if runner is not None else model.unet(x, t, encoder_hidden_states=context)["sample"]
            latents = fused.step(noise_pred, t, latents, guidance_scale)
            if step_callback is not None:
                latents = step_callback(latents)
    finally:
        if runner is not None:
            if stats is not None:
                stats.update(replays=runner.replays, captures=runner.captures, eager_calls=runner.eager_calls, replayed_launches=runner.replayed_launches)
            if own:
                runner.close()
    return latents


@torch.no_grad()
def p2p_edit(model, prompts: List[str], controller, latent: torch.Tensor, num_inference_steps: int = 50, guidance_scale: float = 7.5,
             context: Optional[torch.Tensor] = None, graphs: Union[bool, GraphedUNet] = False, stats: Optional[dict] = None) -> torch.Tensor:
    if controller is not None:
        register_attention_control(model, controller)
    context = encode_prompts(model, prompts) if context is None else context
    latents = (latent * model.scheduler.init_noise_sigma).expand(len(prompts), *latent.shape[1:]).contiguous()
    return denoise(model, latents, context, num_inference_steps, guidance_scale,
                   step_callback=controller.step_callback if controller is not None else None, graphs=graphs, controller=controller, stats=stats)


@torch.no_grad()
def masactrl_edit(model, prompts: List[str], latents: torch.Tensor, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                  context: Optional[torch.Tensor] = None, graphs: Union[bool, GraphedUNet] = False, editor=None,
                  stats: Optional[dict] = None) -> torch.Tensor:
    """The editor must already be registered (masactrl/edit_real.py:137-138); pass it as `editor` with graphs=True."""
    context = encode_prompts(model, prompts) if context is None else context
    if graphs and editor is None:
        raise ValueError("masactrl_edit(graphs=True) needs the registered editor (its step list decides which graph a step replays)")
    return denoise(model, latents, context, num_inference_steps, guidance_scale, graphs=graphs, controller=editor, stats=stats)


@torch.no_grad()
def pnp_edit(model, prompts: List[str], latents: torch.Tensor, num_inference_steps: int = 50, guidance_scale: float = 7.5,
             pnp_attn_t: float = 0.5, pnp_f_t: float = 0.8, context: Optional[torch.Tensor] = None, xl: bool = False,
             graphs: Union[bool, GraphedUNet] = False, stats: Optional[dict] = None) -> torch.Tensor:
    model.scheduler.set_timesteps(num_inference_steps)
    ts = model.scheduler.timesteps
    qk_t, f_t = int(num_inference_steps * pnp_attn_t), int(num_inference_steps * pnp_f_t)
    qk_sched = ts[:qk_t] if qk_t >= 0 else []
    conv_sched = ts[:f_t] if f_t >= 0 else []
    reg = (pnp_register.register_attention_control_efficient_xl, pnp_register.register_conv_control_efficient_xl,
           pnp_register.register_time_xl, pnp_register.unregister_attention_control_efficient_xl,
           pnp_register.unregister_conv_control_efficient_xl) if xl else (
        pnp_register.register_attention_control_efficient, pnp_register.register_conv_control_efficient,
        pnp_register.register_time, pnp_register.unregister_attention_control_efficient,
        pnp_register.unregister_conv_control_efficient)
    reg[0](model, qk_sched)
    reg[1](model, conv_sched)
    context = encode_prompts(model, prompts) if context is None else context
    try:
        qk_set, conv_set = frozenset(int(t) for t in qk_sched), frozenset(int(t) for t in conv_sched)
        return denoise(model, latents, context, num_inference_steps, guidance_scale, per_step=lambda t: reg[2](model, t), graphs=graphs,
                       graph_key_fn=lambda t: (t in qk_set or t == 1000, t in conv_set or t == 1000), stats=stats)
    finally:
        reg[3](model)
        reg[4](model)


def _cross_modules(unet):
    return [(n, m) for n, m in unet.named_modules() if type(m).__name__ == "Attention" and "attn2" in n]


def pix2pix_zero_edit(model, embeds_src: torch.Tensor, embeds_edit: torch.Tensor, latents: torch.Tensor, num_inference_steps: int = 50,
                      guidance_scale: float = 7.5, guidance_amount: float = 0.1, only_sample: bool = False,
                      graphs: Union[bool, GraphedUNet] = False, map_dtype: torch.dtype = torch.float32, uncond_embeddings_list=None, unet_kwargs_src: Optional[dict] = None,
                      unet_kwargs_edit: Optional[dict] = None):
    """Pix2Pix-zero's two denoising loops on latents (pix2pix-zero/model/sd_utils.py:86-182). The UNet must carry MyAttnProcessor
    (prep_unet). embeds_*: [uncond, cond] context pairs ([2, 77, C]); latents: [1, 4, h, w]. Returns (reconstruction latents,
    edited latents).

    Differences from the reference, none of them arithmetic: the reference cross-attention maps of loop 1 stay on the device
    (the reference moves 16 maps per step to the host with a blocking `.cpu()` and back in loop 2, :105-110,169); loop 1 and the
    recomputed-noise pass of loop 2 run the fused kernels (optionally replayed from a CUDA graph); only the guidance pass, which
    needs d loss / d latents, runs differentiable torch arithmetic.

    `uncond_embeddings_list[i]` (null-text inversion) overwrites row 0 of both contexts at step i (P2P_Zero_NTI, :518,:582);
    `unet_kwargs_*` are extra UNet keyword arguments of the two loops (SDXL's added_cond_kwargs), which rule out graph replay."""
    kw_src, kw_edit = unet_kwargs_src or {}, unet_kwargs_edit or {}
    if graphs and (kw_src or kw_edit):
        raise ValueError("pix2pix_zero_edit: graph replay covers unet(x, t, context) only; drop graphs=True when passing UNet kwargs")
    model.scheduler.set_timesteps(num_inference_steps)
    fused = FusedDDIM(model.scheduler)
    ts = model.scheduler.timesteps.tolist()
    cross = _cross_modules(model.unet)
    latents_init = latents.clone()
    # graphs=True: a runner for this call only; a GraphedUNet built once for this unet (controller None) amortises the captures over edits
    own = graphs is True
    runner = GraphedUNet(model.unet, None, None, launch_counter=_cabi.launch_count) if own else (graphs or None)
    ref_maps = {}
    with torch.no_grad():  # loop 1: reference maps (:92-122)
        for i, t in enumerate(ts):
            if uncond_embeddings_list is not None:
                embeds_src[0] = uncond_embeddings_list[i]
            x = torch.cat([latents] * 2)
            replays = runner.replays if runner is not None else 0
            eps = runner(x, t, embeds_src) if runner is not None else model.unet(x, t, encoder_hidden_states=embeds_src, **kw_src)["sample"]
            maps = [m.attn_probs for _, m in cross]
            if runner is not None and runner.replays > replays:
                # a replay writes the buffers the capture allocated, whatever the modules' attn_probs point to by now (the guidance
                # pass of an earlier edit re-pointed them): remember them when the capture happens, read them afterwards
                if getattr(runner, "_map_buffers", None) is None:
                    runner._map_buffers = maps
                maps = runner._map_buffers
            # those buffers are rewritten every step: the cache must own its copy either way
            ref_maps[t] = [p.detach().to(map_dtype, copy=True) for p in maps]
            latents = fused.step(eps, t, latents, guidance_scale)
    latents_rec = latents
    if only_sample:
        if own:
            runner.close()
        return latents_rec, None
    latents = latents_init
    for i, t in enumerate(ts):  # loop 2: cross-attention guidance (:152-182)
        if uncond_embeddings_list is not None:
            embeds_edit[0] = uncond_embeddings_list[i]
        x_in = torch.cat([latents] * 2).detach().clone().requires_grad_(True)
        with torch.enable_grad():
            model.unet(x_in, t, encoder_hidden_states=embeds_edit.detach(), **kw_edit)
            loss = 0.0
            for (_, m), ref in zip(cross, ref_maps[t]):
                loss = loss + ((m.attn_probs.float() - ref.float()) ** 2).sum((1, 2)).mean(0)
            grad, = torch.autograd.grad(loss, x_in)
        with torch.no_grad():
            x_new = (x_in - guidance_amount * grad).detach()   # torch.optim.SGD([x_in], lr=guidance_amount).step()
            eps = runner(x_new, t, embeds_edit) if runner is not None else \
                model.unet(x_new, t, encoder_hidden_states=embeds_edit, **kw_edit)["sample"]
            latents = x_new.chunk(2)[0]
            latents = fused.step(eps, t, latents, guidance_scale)
    if own:
        runner.close()
    return latents_rec, latents
