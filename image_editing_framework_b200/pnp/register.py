"""Plug-and-Play hooks with the names and module tables of pnp/model/register.py.

  register_time[_xl]                          :5-19, :348-365   stamp `t` on the conv module and the attn1 modules
  register_attention_control_efficient[_xl]   :27-88, :188-250  self-attention q/k injection on the decoder layers
  register_conv_control_efficient[_xl]        :100-182, :260-342 feature injection after conv2 of one ResnetBlock2D
  unregister_*                                :91-98, :184-186, :252-258, :344-346

The reference copies q[1]=q[2], k[1]=k[2], q[3]=q[2], k[3]=k[2] and then materialises the probabilities (:48-75).
Here the copies are per-row source indices of ONE ief_attn_fwd launch. `t in injection_schedule` on a CUDA tensor
(a host<->device sync per layer per step, :41-42) is replaced by a python set built at registration.
"""
from __future__ import annotations

import os

import torch

from .. import ops
from ..hooks import project_qkv, out_linear, reject_mask

_SD_ATTN = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}  # decoder blocks 4-11 (reference :82)
_XL_ATTN = {1: [0, 1, 2]}                             # reference :243


def _schedule_set(injection_schedule):
    if injection_schedule is None:
        return None
    if torch.is_tensor(injection_schedule):
        return frozenset(int(x) for x in injection_schedule.tolist())
    return frozenset(int(x) for x in injection_schedule)


def _injecting(module) -> bool:
    sched = getattr(module, "_injection_set", None)
    if sched is None:
        return False
    t = getattr(module, "t", None)
    return t is not None and (int(t) in sched or int(t) == 1000)


def _qk_sources(batch: int):
    """q/k source rows for the 4-group batch [uncond_src, uncond_tgt, cond_src, cond_tgt] (reference :46-52)."""
    s = batch // 4
    src = list(range(batch))
    for r in range(s, 2 * s):
        src[r] = 2 * s + (r - s)
    for r in range(3 * s, 4 * s):
        src[r] = 2 * s + (r - 3 * s)
    return src


def _sa_forward(module):
    def forward(x, encoder_hidden_states=None, attention_mask=None):
        reject_mask(attention_mask)
        is_cross = encoder_hidden_states is not None
        q, k, v = project_qkv(module, x, encoder_hidden_states if is_cross else None)
        if not is_cross and _injecting(module):
            src = _qk_sources(q.shape[0])
            out = ops.attention(q, k, v, module.heads, module.scale, q_src=src, k_src=src)
        elif is_cross and k.shape[1] <= 80:
            out = ops.cross_attention_edit(q, k, v, module.heads, module.scale)
        else:
            out = ops.attention(q, k, v, module.heads, module.scale)
        return out_linear(module)(out.to(x.dtype))

    return forward


def _attn_modules(model, table, all_blocks: bool):
    for res, blocks in table.items():
        for block in blocks:
            tbs = model.unet.up_blocks[res].attentions[block].transformer_blocks
            for tb in (tbs if all_blocks else tbs[:1]):
                yield tb.attn1


def _register_attn(model, injection_schedule, table, all_blocks):
    sched = _schedule_set(injection_schedule)
    for module in _attn_modules(model, table, all_blocks):
        module.ori_forward = module.forward
        module.forward = _sa_forward(module)
        setattr(module, 'injection_schedule', injection_schedule)
        module._injection_set = sched


def _unregister_attn(model, table, all_blocks):
    for module in _attn_modules(model, table, all_blocks):
        module.forward = module.ori_forward


def register_attention_control_efficient(model, injection_schedule):
    _register_attn(model, injection_schedule, _SD_ATTN, all_blocks=False)


def unregister_attention_control_efficient(model):
    _unregister_attn(model, _SD_ATTN, all_blocks=False)


def register_attention_control_efficient_xl(model, injection_schedule):
    _register_attn(model, injection_schedule, _XL_ATTN, all_blocks=True)


def unregister_attention_control_efficient_xl(model):
    _unregister_attn(model, _XL_ATTN, all_blocks=True)


# ---- feature (conv) injection: batch-row copies after conv2 — "next" row of the scope table, plain torch plumbing ----
def _conv_forward(module):
    original = module.forward

    def inject(_conv, _inp, out):
        if _injecting(module):
            s = out.shape[0] // 4
            out[s:2 * s] = out[2 * s:3 * s]
            out[3 * s:4 * s] = out[2 * s:3 * s]
        return out

    def forward(input_tensor, temb, scale=1.):
        handle = module.conv2.register_forward_hook(inject)
        try:
            return original(input_tensor, temb)
        finally:
            handle.remove()

    return forward


def _register_conv(module, injection_schedule):
    module.ori_forward = module.forward
    module.forward = _conv_forward(module)
    setattr(module, 'injection_schedule', injection_schedule)
    module._injection_set = _schedule_set(injection_schedule)


def register_conv_control_efficient(model, injection_schedule):
    _register_conv(model.unet.up_blocks[1].resnets[1], injection_schedule)


def unregister_conv_control_efficient(model):
    m = model.unet.up_blocks[1].resnets[1]
    m.forward = m.ori_forward


def register_conv_control_efficient_xl(model, injection_schedule):
    _register_conv(model.unet.up_blocks[1].resnets[0], injection_schedule)


def unregister_conv_control_efficient_xl(model):
    m = model.unet.up_blocks[1].resnets[0]
    m.forward = m.ori_forward


def register_time(model, t):
    setattr(model.unet.up_blocks[1].resnets[1], 't', t)
    for res, blocks in {1: [0, 1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}.items():
        for block in blocks:
            setattr(model.unet.up_blocks[res].attentions[block].transformer_blocks[0].attn1, 't', t)
    for res, blocks in {0: [0, 1], 1: [0, 1], 2: [0, 1]}.items():
        for block in blocks:
            setattr(model.unet.down_blocks[res].attentions[block].transformer_blocks[0].attn1, 't', t)
    setattr(model.unet.mid_block.attentions[0].transformer_blocks[0].attn1, 't', t)


def register_time_xl(model, t):
    setattr(model.unet.up_blocks[1].resnets[0], 't', t)
    for res, blocks in {0: [0, 1, 2], 1: [0, 1, 2]}.items():
        for block in blocks:
            for tb in model.unet.up_blocks[res].attentions[block].transformer_blocks:
                setattr(tb.attn1, 't', t)
    for res, blocks in {1: [0, 1], 2: [0, 1]}.items():
        for block in blocks:
            for tb in model.unet.down_blocks[res].attentions[block].transformer_blocks:
                setattr(tb.attn1, 't', t)
    for tb in model.unet.mid_block.attentions[0].transformer_blocks:
        setattr(tb.attn1, 't', t)


def load_source_latents_t(t, latents_path):
    """The inversion latent saved for timestep t as `noisy_latents_{t}.pt` (pnp/model/register.py:21-25)."""
    path = os.path.join(latents_path, f"noisy_latents_{t}.pt")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no inversion latent for timestep {t}: {path}")
    return torch.load(path)
