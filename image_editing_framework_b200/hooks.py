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


_warned_downcast = False


def compute_dtype(t: torch.Tensor) -> torch.dtype:
    """Kernel arithmetic type: the tensor's own 16-bit type, else IEF_COMPUTE_DTYPE (default bf16). An fp32 pipeline (what the
    reference's scripts build) therefore has its attention — and the stored probability maps that feed the LocalBlend / MaskAuto
    thresholds and the pix2pix-zero loss — computed from 16-bit q, k, v with fp32 accumulation: within 2e-2 of the fp32 result
    (the tolerance BASELINE.json states), not bit-equal to it. Said once, loudly, instead of silently."""
    global _warned_downcast
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.dtype
    choice = os.environ.get("IEF_COMPUTE_DTYPE", "bf16")
    if not _warned_downcast:
        _warned_downcast = True
        import warnings
        warnings.warn(f"image_editing_framework_b200: {t.dtype} attention inputs are computed in {choice} on the tensor cores (fp32 accumulation, "
                      "outputs within 2e-2 of fp32). IEF_COMPUTE_DTYPE=fp16 trades range for 3 more mantissa bits.", stacklevel=3)
    return _COMPUTE[choice]


def _fused_weight(module, names: Tuple[str, ...]):
    """Row-concatenation of the named projection weights (and biases), cached on the module and rebuilt when a parameter is
    replaced or modified in place. None when the projections cannot be fused (LoRA-wrapped layers, mixed bias / dtype)."""
    layers = [getattr(module, n) for n in names]
    if not all(type(l) is torch.nn.Linear for l in layers):
        return None
    ws = [l.weight for l in layers]
    bs = [l.bias for l in layers]
    if any(w.requires_grad and torch.is_grad_enabled() for w in ws) or len({w.dtype for w in ws}) != 1 or len({b is None for b in bs}) != 1:
        return None
    key = tuple((w.data_ptr(), w._version) for w in ws) + tuple((b.data_ptr(), b._version) for b in bs if b is not None)
    cache = module.__dict__.setdefault("_ief_fused", {})
    hit = cache.get(names)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            w = torch.cat([x.detach() for x in ws], 0).contiguous()
            b = None if bs[0] is None else torch.cat([x.detach() for x in bs], 0).contiguous()
        hit = cache[names] = (key, w, b)
    return hit[1], hit[2]


def fused_projections_enabled() -> bool:
    return os.environ.get("IEF_FUSED_QKV", "1") != "0"


def project_qkv(module, hidden_states: torch.Tensor, context: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """to_q / to_k / to_v in the module's dtype, then cast (if needed) to the kernel dtype. Shapes stay [B, N, H*d].

    SURVEY.md section 8(f) rank 1: on CUDA under no_grad the three (self-attention) or two (cross-attention: K and V of the
    context) projections run as ONE GEMM against the row-concatenated weights; q, k, v come back as strided views of its
    output — the kernels read [B, N, H, d] through explicit strides, so nothing is copied or permuted afterwards."""
    src = hidden_states if context is None else context
    q = k = v = None
    if hidden_states.is_cuda and not torch.is_grad_enabled() and fused_projections_enabled():
        c = module.to_q.out_features if type(module.to_q) is torch.nn.Linear else 0
        if context is None:
            fw = _fused_weight(module, ("to_q", "to_k", "to_v"))
            if fw is not None and c % 8 == 0:
                qkv = torch.nn.functional.linear(hidden_states, fw[0], fw[1])
                q, k, v = qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]
        else:
            fw = _fused_weight(module, ("to_k", "to_v"))
            if fw is not None and c % 8 == 0:
                q = module.to_q(hidden_states)
                kv = torch.nn.functional.linear(src, fw[0], fw[1])
                k, v = kv[..., :c], kv[..., c:]
    if q is None:
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
