"""MutualSelfAttentionControl (+ Union) with the constructor of masactrl/model/attention_control.py:10-35.

The reference stacks the queries of all rows of a CFG half along the sequence axis and attends to the keys/values of
the half's first (source) row (attn_batch :37-50, forward :52-68). That is plain attention with per-row source
indices: row r of a half reads K,V of the half's row 0 — one ief_attn_fwd launch, no [h, 2N, N] tensor.
"""
from __future__ import annotations

from .. import ops
from .attention_base import AttentionBase


class MutualSelfAttentionControl(AttentionBase):
    MODEL_TYPE = {"SD": 16, "SDXL": 70}

    def __init__(self, start_step=4, start_layer=10, layer_idx=None, step_idx=None, total_steps=50, model_type="SD"):
        super().__init__()
        self.total_steps = total_steps
        self.total_layers = self.MODEL_TYPE.get(model_type, 16)
        self.start_step = start_step
        self.start_layer = start_layer
        self.layer_idx = layer_idx if layer_idx is not None else list(range(start_layer, self.total_layers))
        self.step_idx = step_idx if step_idx is not None else list(range(start_step, total_steps))
        self._layers, self._steps = frozenset(self.layer_idx), frozenset(self.step_idx)
        print("MasaCtrl at denoising steps: ", self.step_idx)
        print("MasaCtrl at U-Net layers: ", self.layer_idx)

    def _controlled(self, is_cross) -> bool:
        # reference gate :56 — note the self-attention layer index is cur_att_layer // 2
        return not (is_cross or self.cur_step not in self._steps or self.cur_att_layer // 2 not in self._layers)

    def fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale):
        if not self._controlled(is_cross):
            return super().fused_forward(q, k, v, is_cross, place_in_unet, num_heads, scale)
        B = q.shape[0]
        half = B // 2
        src = [0 if b < half else half for b in range(B)]  # ku[:num_heads] / kc[:num_heads]: first row of each CFG half
        return ops.attention(q, k, v, num_heads, scale, k_src=src, v_src=src)


class MutualSelfAttentionControlUnion(MutualSelfAttentionControl):
    """Target rows attend to the concatenation [K_src; K_tgt] (reference :87-107); source rows stay plain."""

    def fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale):
        if not self._controlled(is_cross):
            return AttentionBase.fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale)
        B = q.shape[0]
        if B != 4:
            raise ValueError(f"MutualSelfAttentionControlUnion expects the 4-row batch (u_s, u_t, c_s, c_t), got {B} rows")
        out = ops.attention(q, k, v, num_heads, scale, rows=[0, 2])
        # rows 1 and 3: keys/values = [source row ; own row]
        return ops.attention(q, k, v, num_heads, scale, k_src=[0, 0, 2, 2], v_src=[0, 0, 2, 2], k_src2=[0, 1, 2, 3], v_src2=[0, 1, 2, 3],
                             rows=[1, 3], out=out)
