"""regiter_attention_editor_diffusers [sic] / unregister_attention_control — names, discovery rule and closure
signature of masactrl/model/register.py:6-89. The reference closure materialises `sim` and `attn` for EVERY layer
(:35,44) even when the editor recomputes its own; here only q, k, v are projected and the editor runs one kernel.
"""
from __future__ import annotations

from ..hooks import project_qkv, out_linear, reject_mask
from .attention_base import AttentionBase


def _make_forward(module, editor, place_in_unet):
    def forward(x, encoder_hidden_states=None, attention_mask=None, context=None, mask=None):
        if encoder_hidden_states is not None:
            context = encoder_hidden_states
        if attention_mask is not None:
            mask = attention_mask
        reject_mask(mask)
        is_cross = context is not None
        q, k, v = project_qkv(module, x, context)
        out = editor.attend(q, k, v, is_cross, place_in_unet, module.heads, module.scale)
        return out_linear(module)(out.to(x.dtype))

    return forward


def regiter_attention_editor_diffusers(model, editor: AttentionBase):
    def visit(net, place_in_unet) -> int:
        # the reference tests the class name of the parent while iterating its children (:54-62): a module with no
        # children is never patched, and an 'Attention' module is patched once
        count = 0
        for _, subnet in net.named_children():
            if net.__class__.__name__ == 'Attention':
                if "_original_forward" not in vars(net):      # a second registration keeps the true original
                    net._original_forward = net.forward
                net.forward = _make_forward(net, editor, place_in_unet)
                return count + 1
            count += visit(subnet, place_in_unet)
        return count

    total = 0
    for name, net in model.unet.named_children():
        for place in ("down", "mid", "up"):  # reference precedence :66-71
            if place in name:
                total += visit(net, place)
                break
    editor.num_att_layers = total
    model.unet._ief_installed = editor      # lets the pipeline classes find the editor whose phases key their CUDA graphs


def unregister_attention_control(model, editor):
    def restore(net):
        for _, subnet in net.named_children():
            if '_original_forward' in vars(net):
                net.forward = vars(net).pop('_original_forward')
            else:
                restore(subnet)

    for name, net in model.unet.named_children():
        if "down" in name or "mid" in name or "up" in name:
            restore(net)
    editor.num_att_layers = 0
    model.unet._ief_installed = None
