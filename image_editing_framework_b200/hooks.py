"""Plumbing shared by the four register modules: q/k/v projection around the fused kernels.

The reference closures (p2p/model/register.py:11-64, masactrl/model/register.py:11-50, pnp/model/register.py:35-78,
pix2pix-zero/model/attention_control.py:5-64) all do: project -> head_to_batch_dim -> materialise softmax(QK^T) ->
edit -> bmm -> batch_to_head_dim -> to_out. Here the three middle stages are one kernel reading the
[B, N, H*d] projections in place, so only the projection and the output linear remain in torch.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

_COMPUTE = {"bf16": torch.bfloat16, "fp16": torch.float16}


def compute_dtype(t: torch.Tensor) -> torch.dtype:
    """Kernel arithmetic type: the tensor's own 16-bit type, else IEF_COMPUTE_DTYPE (default bf16)."""
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.dtype
    return _COMPUTE[os.environ.get("IEF_COMPUTE_DTYPE", "bf16")]


def project_qkv(module, hidden_states: torch.Tensor, context: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """to_q / to_k / to_v in the module's dtype, then cast (if needed) to the kernel dtype. Shapes stay [B, N, H*d]."""
    src = hidden_states if context is None else context
    q, k, v = module.to_q(hidden_states), module.to_k(src), module.to_v(src)
    dt = compute_dtype(q)
    if q.dtype != dt:
        q, k, v = q.to(dt), k.to(dt), v.to(dt)
    return q, k, v


def out_linear(module):
    to_out = module.to_out
    return to_out[0] if isinstance(to_out, torch.nn.ModuleList) else to_out


def reject_mask(attention_mask) -> None:
    if attention_mask is not None:
        raise NotImplementedError(
            "the fused controlled-attention kernels take no attention_mask (no reference script passes one); "
            "refusing to silently ignore it")
