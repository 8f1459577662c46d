"""Pix2Pix-zero attention processor — MyAttnProcessor / prep_unet / restore_original_processors of
pix2pix-zero/model/attention_control.py:4-96.

Under `torch.no_grad()` (two of the three UNet passes per step, pix2pix-zero/model/sd_utils.py:92-122,177-178)
the processor runs the fused kernels: cross-attention layers emit their [B*heads, N, 77] fp32 probabilities straight
from the kernel into `attn.attn_probs` (:46 — the only maps the method ever reads, sd_utils.py:108-110,169-171);
self-attention layers run the flash kernel and stash nothing (the reference keeps every 4096^2 map alive, unread).
When autograd is recording (the guidance pass, sd_utils.py:163-174) the cross-attention probabilities must stay
differentiable: on CUDA they come from `_CrossAttention`, an autograd Function over the forward kernel and
ief_cross_attn_bwd; self-attention layers (maps never read) use library SDPA; on CPU, with an attention mask or with
IEF_P2Z_KERNEL_BACKWARD=0 the pass runs the reference's torch arithmetic.
"""
from __future__ import annotations

import torch

from .. import ops
import os

from ..hooks import compute_dtype, project_qkv, reject_mask


def _kernel_backward_enabled() -> bool:
    return os.environ.get("IEF_P2Z_KERNEL_BACKWARD", "1") != "0"


class _CrossAttention(torch.autograd.Function):
    """Differentiable 77-key cross-attention on the kernels: forward = ief_cross_attn_edit_fwd with the probability output (the
    maps the guidance loss is built on), backward = ief_cross_attn_bwd (dQ, dS) plus two small batched GEMMs for dK and dV."""

    @staticmethod
    def forward(ctx, q, k, v, heads: int, scale: float):
        B, N, M = q.shape[0], q.shape[1], k.shape[1]
        probs = torch.empty((B * heads, N, M), dtype=torch.float32, device=q.device)
        out = ops.cross_attention_edit(q, k, v, heads, scale, probs_out=probs)
        ctx.save_for_backward(q, k, v, probs)
        ctx.heads, ctx.scale = heads, scale
        return out, probs

    @staticmethod
    def backward(ctx, dout, dprobs):
        q, k, v, probs = ctx.saved_tensors
        heads, scale = ctx.heads, ctx.scale
        need_kv = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dprobs = None if dprobs is None else dprobs.contiguous().float()
        dq, ds = ops.cross_attention_backward(q, k, v, dout.contiguous().to(q.dtype), heads, scale, dprobs=dprobs, want_ds=need_kv)
        dk = dv = None
        if need_kv:
            B, N, C = q.shape
            M, d = k.shape[1], C // heads
            qh = q.reshape(B, N, heads, d).permute(0, 2, 1, 3).reshape(B * heads, N, d).float()
            gh = dout.reshape(B, N, heads, d).permute(0, 2, 1, 3).reshape(B * heads, N, d).float()
            dk = torch.bmm(ds.transpose(1, 2), qh).reshape(B, heads, M, d).permute(0, 2, 1, 3).reshape(B, M, C).to(k.dtype)   # dK = dS^T Q
            dv = torch.bmm(probs.transpose(1, 2), gh).reshape(B, heads, M, d).permute(0, 2, 1, 3).reshape(B, M, C).to(v.dtype)  # dV = P^T dO
        return dq, dk, dv, None, None


class MyAttnProcessor:
    def __init__(self, stash_self_probs: bool = False):
        self.stash_self_probs = stash_self_probs

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        spatial = hidden_states.ndim == 4
        if spatial:
            b, c, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        context = None
        if encoder_hidden_states is not None:
            context = attn.norm_encoder_hidden_states(encoder_hidden_states) if attn.norm_cross else encoder_hidden_states

        needs_grad = torch.is_grad_enabled() and (hidden_states.requires_grad or (context is not None and context.requires_grad)
                                                  or attn.to_q.weight.requires_grad)
        if needs_grad:
            out = self._autograd_pass(attn, hidden_states, context, attention_mask)
        else:
            reject_mask(attention_mask)
            q, k, v = project_qkv(attn, hidden_states, context)
            B, N, M = q.shape[0], q.shape[1], k.shape[1]
            if context is not None and M <= 80:
                probs = torch.empty((B * attn.heads, N, M), dtype=torch.float32, device=q.device)
                out = ops.cross_attention_edit(q, k, v, attn.heads, attn.scale, probs_out=probs)
                attn.attn_probs = probs
            elif self.stash_self_probs:
                probs = torch.empty((B * attn.heads, N, M), dtype=torch.float32, device=q.device)
                out = ops.attention(q, k, v, attn.heads, attn.scale, probs_out=probs)
                attn.attn_probs = probs
            else:
                out = ops.attention(q, k, v, attn.heads, attn.scale)
                attn.attn_probs = None
            out = out.to(hidden_states.dtype)
        out = attn.to_out[1](attn.to_out[0](out))  # linear proj + dropout (:51-53)
        if spatial:
            out = out.transpose(-1, -2).reshape(b, c, hh, ww)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor

    @staticmethod
    def _autograd_pass(attn, hidden_states, context, attention_mask):
        # out of scope for the forward-only kernels (needs d loss / d probs): differentiable torch arithmetic
        src = hidden_states if context is None else context
        if context is None and hidden_states.is_cuda and attention_mask is None:
            # self-attention maps are never read (sd_utils.py:108-110,169-171): a memory-efficient differentiable library
            # attention instead of retaining N x N fp32 probabilities of every layer for the backward pass
            h = attn.heads
            q, k, v = attn.to_q(hidden_states), attn.to_k(src), attn.to_v(src)
            q4, k4, v4 = (t.view(t.shape[0], t.shape[1], h, t.shape[2] // h).transpose(1, 2) for t in (q, k, v))
            out = torch.nn.functional.scaled_dot_product_attention(q4, k4, v4, scale=attn.scale)
            attn.attn_probs = None
            return out.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)
        if context is not None and hidden_states.is_cuda and attention_mask is None and context.shape[1] <= 80 and _kernel_backward_enabled():
            # the maps the loss reads, with their gradient, on the fused kernels (forward + ief_cross_attn_bwd)
            q, k, v = attn.to_q(hidden_states), attn.to_k(src), attn.to_v(src)
            dt = compute_dtype(q)
            out, probs = _CrossAttention.apply(q.to(dt), k.to(dt), v.to(dt), attn.heads, attn.scale)
            attn.attn_probs = probs
            return out.to(q.dtype)
        q = attn.head_to_batch_dim(attn.to_q(hidden_states))
        k = attn.head_to_batch_dim(attn.to_k(src))
        v = attn.head_to_batch_dim(attn.to_v(src))
        probs = attn.get_attention_scores(q, k, attention_mask)
        attn.attn_probs = probs
        return attn.batch_to_head_dim(torch.bmm(probs, v))


def prep_unet(unet):
    """attn2 parameters trainable, everything else frozen; every `Attention` gets a MyAttnProcessor (:76-90)."""
    original_processors = {}
    for name, params in unet.named_parameters():
        params.requires_grad = 'attn2' in name
    for name, module in unet.named_modules():
        if type(module).__name__ == "Attention":
            original_processors[name] = module.get_processor()
            module.set_processor(MyAttnProcessor())
    return unet, original_processors


def restore_original_processors(unet, original_processors):
    for name, module in unet.named_modules():
        if type(module).__name__ == "Attention" and name in original_processors:
            module.set_processor(original_processors[name])
