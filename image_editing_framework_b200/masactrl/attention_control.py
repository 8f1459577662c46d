"""MutualSelfAttentionControl (+ Union, Mask, MaskAuto) with the constructors of masactrl/model/attention_control.py.

The reference stacks the queries of all rows of a CFG half along the sequence axis and attends to the keys/values of
the half's first (source) row (attn_batch :37-50, forward :52-68). That is plain attention with per-row source
indices: row r of a half reads K,V of the half's row 0 — one ief_attn_fwd launch, no [h, 2N, N] tensor.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np
import torch
import torch.nn.functional as F

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

    def graph_key(self):
        base = super().graph_key()
        return None if base is None else base + (self.cur_step in self._steps,)

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


def _save_mask_png(mask2d: torch.Tensor, path: str) -> None:
    """8-bit grey PNG of a [H, W] mask in [0, 1] (the reference calls torchvision's save_image, :126-127, 298, 308)."""
    img = (mask2d.detach().float().clamp(0, 1) * 255 + 0.5).to(torch.uint8).cpu().numpy()
    h, w = img.shape
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw))
                + chunk(b"IEND", b""))


def _key_biases(mask_flat: torch.Tensor) -> torch.Tensor:
    """[2, N] fp32 additive key biases (foreground, background) of a flattened source mask, exactly as the reference builds
    them (:143-145, 244-245): masked keys get finfo.min, the others get the mask value itself added to their score."""
    m = mask_flat.to(torch.float32)
    fmin = torch.finfo(torch.float32).min
    return torch.stack([m.masked_fill(m == 0, fmin), m.masked_fill(m == 1, fmin)]).contiguous()


def _binarize(m: torch.Tensor, thres: float) -> torch.Tensor:
    """`m[m >= thres] = 1; m[m < thres] = 0` (:241-243, 312-314) without the host sync of boolean-index assignment."""
    m = torch.where(m >= thres, torch.ones_like(m), m)
    return torch.where(m < thres, torch.zeros_like(m), m)


class _MaskedMutualMixin:
    """Rows (u_s, u_t, c_s, c_t): source rows plain; target rows attend to the source row's K/V twice — foreground keys
    only and background keys only (two softmaxes) — and are blended per query position with the target mask
    (MutualSelfAttentionControlMask.forward :152-181, MaskAuto.forward :271-326).

    Three ief_attn_fwd launches (source rows; fg pass; bg pass) + one ief_mask_blend;
    the reference materialises [2h, N, N] scores and probabilities per target row instead."""

    def _masked_forward(self, q, k, v, num_heads, scale, key_bias: torch.Tensor, spatial_w: torch.Tensor) -> torch.Tensor:
        if q.shape[0] != 4:
            raise ValueError(f"masked MasaCtrl expects the 4-row batch (u_s, u_t, c_s, c_t), got {q.shape[0]} rows")
        src = [0, 0, 2, 2]
        out = ops.attention(q, k, v, num_heads, scale, rows=[0, 2])
        ops.attention(q, k, v, num_heads, scale, k_src=src, v_src=src, key_bias=key_bias, bias_sel=[-1, 0, -1, 0], rows=[1, 3], out=out)
        bg = torch.empty_like(out)
        ops.attention(q, k, v, num_heads, scale, k_src=src, v_src=src, key_bias=key_bias, bias_sel=[-1, 1, -1, 1], rows=[1, 3], out=bg)
        return ops.mask_blend(out, bg, spatial_w.to(torch.float32).reshape(-1).contiguous(), rows=[1, 3])


class MutualSelfAttentionControlMask(_MaskedMutualMixin, MutualSelfAttentionControl):
    """Mask-guided MasaCtrl with user-supplied source / target masks of shape (h, w) (reference :110-181)."""

    def __init__(self, start_step=4, start_layer=10, layer_idx=None, step_idx=None, total_steps=50, mask_s=None, mask_t=None,
                 mask_save_dir=None, model_type="SD"):
        super().__init__(start_step, start_layer, layer_idx, step_idx, total_steps, model_type)
        self.mask_s = mask_s
        self.mask_t = mask_t
        print("Using mask-guided MasaCtrl")
        if mask_save_dir is not None:
            os.makedirs(mask_save_dir, exist_ok=True)
            _save_mask_png(self.mask_s, os.path.join(mask_save_dir, "mask_s.png"))
            _save_mask_png(self.mask_t, os.path.join(mask_save_dir, "mask_t.png"))
        self._tables = {}  # (H, W) -> (key biases [2, N], target weights [N]) on the compute device

    def _mask_tables(self, res: int, device):
        key = (res, str(device))
        if key not in self._tables:
            ms = F.interpolate(self.mask_s.unsqueeze(0).unsqueeze(0).float(), (res, res)).flatten().to(device)
            mt = F.interpolate(self.mask_t.unsqueeze(0).unsqueeze(0).float(), (res, res)).flatten().to(device)
            self._tables[key] = (_key_biases(ms), mt.contiguous())
        return self._tables[key]

    def fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale):
        if not self._controlled(is_cross) or self.mask_s is None:
            return super().fused_forward(q, k, v, is_cross, place_in_unet, num_heads, scale)
        if self.mask_t is None:
            raise ValueError("MutualSelfAttentionControlMask needs mask_t when mask_s is given (the reference returns a 6-row batch here)")
        bias, w = self._mask_tables(int(np.sqrt(q.shape[1])), q.device)
        return self._masked_forward(q, k, v, num_heads, scale, bias, w)


class MutualSelfAttentionControlMaskAuto(_MaskedMutualMixin, MutualSelfAttentionControl):
    """MasaCtrl with the masks derived from the step's 16x16 cross-attention maps (reference :184-326): the source-key mask from
    the source prompt's `ref_token_idx` columns, the target spatial mask from the target prompt's `cur_token_idx` columns."""

    def __init__(self, start_step=4, start_layer=10, layer_idx=None, step_idx=None, total_steps=50, thres=0.1, ref_token_idx=[1],
                 cur_token_idx=[1], mask_save_dir=None, model_type="SD"):
        super().__init__(start_step, start_layer, layer_idx, step_idx, total_steps, model_type)
        print("Using MutualSelfAttentionControlMaskAuto")
        self.thres = thres
        self.ref_token_idx = ref_token_idx
        self.cur_token_idx = cur_token_idx
        self.self_attns = []
        self.cross_attns = []
        self.cross_attns_mask = None
        self.self_attns_mask = None
        self.mask_save_dir = mask_save_dir
        if self.mask_save_dir is not None:
            os.makedirs(self.mask_save_dir, exist_ok=True)

    def after_step(self):
        self.self_attns = []
        self.cross_attns = []

    def aggregate_cross_attn_map(self, idx):
        attn_map = torch.stack(self.cross_attns, dim=1).mean(1)  # (B, N, dim)
        res = int(np.sqrt(attn_map.shape[-2]))
        attn_map = attn_map.reshape(-1, res, res, attn_map.shape[-1])
        if isinstance(idx, list):
            # = attn_map[..., idx].sum(-1); python-int column views instead of a list index, which would build an index
            # tensor on the host and copy it over (not allowed inside CUDA-graph capture)
            image = attn_map[..., idx[0]]
            for i in idx[1:]:
                image = image + attn_map[..., i]
        else:
            image = attn_map[..., idx]
        image_min = image.min(dim=1, keepdim=True)[0].min(dim=2, keepdim=True)[0]
        image_max = image.max(dim=1, keepdim=True)[0].max(dim=2, keepdim=True)[0]
        return (image - image_min) / (image_max - image_min)

    def fused_forward(self, q, k, v, is_cross, place_in_unet, num_heads, scale):
        if is_cross and q.shape[1] == 16 * 16:
            # head-averaged 16x16 cross maps of every row (:262-265), from the cross kernel's probability output
            B, N, M = q.shape[0], q.shape[1], k.shape[1]
            probs = torch.empty((B, num_heads, N, M), dtype=torch.float32, device=q.device)
            if M <= 80:
                out = ops.cross_attention_edit(q, k, v, num_heads, scale, probs_out=probs)
            else:
                out = ops.attention(q, k, v, num_heads, scale, probs_out=probs)
            self.cross_attns.append(probs.mean(1))
            return out
        if not self._controlled(is_cross) or len(self.cross_attns) == 0:
            self.self_attns_mask = None
            return super().fused_forward(q, k, v, is_cross, place_in_unet, num_heads, scale)
        res = int(np.sqrt(q.shape[1]))
        mask_source = self.aggregate_cross_attn_map(idx=self.ref_token_idx)[-2]  # cond source row
        self.self_attns_mask = F.interpolate(mask_source.unsqueeze(0).unsqueeze(0), (res, res)).flatten()
        if self.mask_save_dir is not None:
            _save_mask_png(self.self_attns_mask.reshape(res, res), os.path.join(self.mask_save_dir, f"mask_s_{self.cur_step}_{self.cur_att_layer}.png"))
        mask_target = self.aggregate_cross_attn_map(idx=self.cur_token_idx)[-1]  # cond target row
        spatial_mask = F.interpolate(mask_target.unsqueeze(0).unsqueeze(0), (res, res)).reshape(-1)
        if self.mask_save_dir is not None:
            _save_mask_png(spatial_mask.reshape(res, res), os.path.join(self.mask_save_dir, f"mask_t_{self.cur_step}_{self.cur_att_layer}.png"))
        bias = _key_biases(_binarize(self.self_attns_mask, self.thres))
        out = self._masked_forward(q, k, v, num_heads, scale, bias, _binarize(spatial_mask, self.thres))
        self.self_attns_mask = None
        return out
