"""MasaCtrl editors — classes and counters of masactrl/model/attention_base.py:5-66 on the fused kernels.

`AttentionBase.attend(q, k, v, is_cross, place, num_heads, scale)` is what the registered closure calls with the
[B, N, H*d] projections. `AttentionBase.__call__` keeps the reference signature
`(q, k, v, sim, attn, is_cross, place_in_unet, num_heads, **kwargs)` for callers that still pass the head-major
'(b h) n d' tensors: `sim` / `attn` are ignored (may be None) — nothing materialises them any more.
"""
from __future__ import annotations

import abc
from typing import List

import torch

from .. import ops


def _heads_last(t: torch.Tensor, num_heads: int) -> torch.Tensor:
    """'(b h) n d' -> strided [b, n, h, d] view (no copy)."""
    bh, n, d = t.shape
    return t.view(bh // num_heads, num_heads, n, d).permute(0, 2, 1, 3)


class AttentionBase(abc.ABC):
    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0

    def after_step(self):
        pass

    def _tick(self):
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.after_step()

    def attend(self, q, k, v, is_cross, place_in_unet, num_heads, scale) -> torch.Tensor:
        out = self.fused_forward(q, k, v, is_cross, place_in_unet, num_heads, scale)
        self._tick()
        return out

    def __call__(self, q, k, v, sim, attn, is_cross, place_in_unet, num_heads, **kwargs):
        scale = kwargs.get("scale")
        if scale is None:
            scale = q.shape[-1] ** -0.5
        return self.attend(_heads_last(q, num_heads), _heads_last(k, num_heads), _heads_last(v, num_heads), is_cross,
                           place_in_unet, num_heads, scale)

    def fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale) -> torch.Tensor:
        """Plain attention, output already in 'b n (h d)' (reference forward :24-27)."""
        if is_cross and k.shape[1] <= 80:
            return ops.cross_attention_edit(q, k, v, num_heads, scale)
        return ops.attention(q, k, v, num_heads, scale)

    def forward(self, q, k, v, sim, attn, is_cross, place_in_unet, num_heads, **kwargs):
        return self.fused_forward(_heads_last(q, num_heads), _heads_last(k, num_heads), _heads_last(v, num_heads), is_cross,
                                  place_in_unet, num_heads, kwargs.get("scale") or q.shape[-1] ** -0.5)

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    # ---- CUDA-graph replay protocol (graphs.GraphedUNet) ---------------------------------------------------------
    _graph_mode = False

    def graph_key(self):
        return () if self.cur_att_layer == 0 else None

    def graph_prepare(self) -> None:
        return None

    def graph_advance(self) -> None:
        self.cur_att_layer = 0
        self.cur_step += 1
        self.after_step()


class AttentionStore(AttentionBase):
    """masactrl/model/attention_base.py:33-66: keeps every map with N <= 64^2 for steps in (min_step, max_step).
    Maps come from the kernels' probability output; the per-step `+=` loop is one ief_store_accumulate launch."""

    def __init__(self, res=[32], min_step=0, max_step=1000):
        super().__init__()
        self.res = res
        self.min_step = min_step
        self.max_step = max_step
        self.valid_steps = 0
        self.self_attns: List[torch.Tensor] = []
        self.cross_attns: List[torch.Tensor] = []
        self.self_attns_step: List[torch.Tensor] = []
        self.cross_attns_step: List[torch.Tensor] = []

    def graph_key(self):
        return None  # keeps per-step python lists of freshly allocated maps: not replayable

    def after_step(self):
        if self.cur_step > self.min_step and self.cur_step < self.max_step:
            self.valid_steps += 1
            if len(self.self_attns) == 0:
                self.self_attns = list(self.self_attns_step)
                self.cross_attns = list(self.cross_attns_step)
            else:
                n = min(len(self.self_attns), len(self.cross_attns))  # the reference indexes both lists by len(self_attns)
                ops.store_accumulate(self.self_attns[:n] + self.cross_attns[:n], self.self_attns_step[:n] + self.cross_attns_step[:n])
        self.self_attns_step = []
        self.cross_attns_step = []

    def fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale):
        B, N, M = q.shape[0], q.shape[1], k.shape[1]
        probs = None
        if N <= 64 ** 2:  # reference :61 "avoid OOM"
            probs = torch.empty((B * num_heads, N, M), dtype=torch.float32, device=q.device)
            (self.cross_attns_step if is_cross else self.self_attns_step).append(probs)
        if is_cross and M <= 80:
            return ops.cross_attention_edit(q, k, v, num_heads, scale, probs_out=probs)
        return ops.attention(q, k, v, num_heads, scale, probs_out=probs)
