"""register_attention_control / unregister_attention_control with the reference's names, discovery rule and closure
signature (p2p/model/register.py:3-117). The closure keeps every torch step that is NOT on the hot path
(spatial/group norm, q/k/v and output projections, residual, rescale) and hands [B, N, H*d] projections to
`controller.attend`, which runs one fused kernel instead of get_attention_scores -> controller -> bmm (:47-50).
"""
from __future__ import annotations

import torch

from .. import ops
from ..hooks import project_qkv, out_linear, reject_mask


class DummyController:
    """Stands in when `controller is None` (reference :66-73): plain attention, no state."""

    def __init__(self):
        self.num_att_layers = 0

    def attend(self, q, k, v, heads, scale, is_cross, place_in_unet):
        if is_cross and k.shape[1] <= 80:
            return ops.cross_attention_edit(q, k, v, heads, scale)
        return ops.attention(q, k, v, heads, scale)


def _make_forward(module, controller, place_in_unet):
    proj = out_linear(module)  # dropout (to_out[1]) is skipped, as in the reference :5-9,54

    def forward(hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        is_cross = encoder_hidden_states is not None
        residual = hidden_states
        if module.spatial_norm is not None:
            hidden_states = module.spatial_norm(hidden_states, temb)
        spatial = hidden_states.ndim == 4
        if spatial:
            b, c, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        reject_mask(attention_mask)
        if module.group_norm is not None:
            hidden_states = module.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        context = None
        if is_cross:
            context = module.norm_encoder_hidden_states(encoder_hidden_states) if module.norm_cross else encoder_hidden_states
        q, k, v = project_qkv(module, hidden_states, context)
        out = controller.attend(q, k, v, module.heads, module.scale, is_cross, place_in_unet)
        out = proj(out.to(hidden_states.dtype))
        if spatial:
            out = out.transpose(-1, -2).reshape(b, c, hh, ww)
        if module.residual_connection:
            out = out + residual
        return out / module.rescale_output_factor

    return forward


def _walk(net, place_in_unet, controller) -> int:
    if net.__class__.__name__ == 'Attention':
        if "_original_forward" not in vars(net):      # registering again (every call of the pipeline classes does) keeps the true original
            net._original_forward = net.forward
        net.forward = _make_forward(net, controller, place_in_unet)
        return 1
    return sum(_walk(child, place_in_unet, controller) for child in net.children()) if hasattr(net, 'children') else 0


def _places(model):
    for name, child in model.unet.named_children():
        for place in ("down", "up", "mid"):  # same precedence as the reference's if/elif chain :89-95
            if place in name:
                yield place, child
                break


def register_attention_control(model, controller):
    if controller is None:
        controller = DummyController()
    controller.num_att_layers = sum(_walk(child, place, controller) for place, child in _places(model))
    model.unet._ief_installed = controller


def unregister_attention_control(model, controller):
    def restore(net):
        if '_original_forward' in vars(net):
            net.forward = vars(net).pop('_original_forward')
        elif hasattr(net, 'children'):
            for child in net.children():
                restore(child)

    for _, child in _places(model):
        restore(child)
    controller.num_att_layers = 0
    model.unet._ief_installed = None
