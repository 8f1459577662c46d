"""The three Prompt-to-Prompt edit controllers, constructor-compatible with p2p/model/attention_control.py:8-46.

In the reference each class overrides `replace_cross_attention(attn_base, att_replace)`, which transforms a materialised
[heads, pixels, 77] probability tensor:

    AttentionReplace   einsum('hpw,bwn->bhpn', base, mapper)                        (:15-16)
    AttentionRefine    base[:, :, mapper] * alphas + replace * (1 - alphas)        (:28-31)
    AttentionReweight  (previous controller's edit, if any) * equalizer             (:42-46)

Here no such tensor exists. A controller only owns the small device tables of its edit and describes them to the fused
cross-attention kernel through `cross_edit()`; ief_cross_attn_edit_fwd applies the edit to the probabilities while they are
still in registers (csrc/cross_attn.cu). Public attributes the reference exposes (`mapper`, `alphas`, `equalizer`,
`prev_controller`) keep their names, shapes and dtypes.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from . import seq_aligner
from .attention_base import AttentionControlEdit
from .ptp_utils import LocalBlend

_CUDA0 = torch.device("cuda:0")


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


def _expect(table: torch.Tensor, shape, what: str) -> torch.Tensor:
    """The kernel indexes these tables with (edit slot, token): a mis-shaped one must fail here, not read out of bounds there."""
    if tuple(table.shape) != tuple(shape):
        raise ValueError(f"{what}: expected shape {tuple(shape)}, got {tuple(table.shape)} — one row per target prompt, "
                         "tokenizer.model_max_length tokens")
    return table


class AttentionReplace(AttentionControlEdit):
    """Word swap: the target prompt's probabilities become the source prompt's, routed through the [77, 77] token mapper."""

    def __init__(
        self,
        prompts,
        tokenizer,
        num_steps: int,
        cross_replace_steps: float,
        self_replace_steps: float,
        local_blend: Optional[LocalBlend] = None,
        device=_CUDA0,
        LOW_RESOURCE=False,
        dtype=torch.float32,
    ):
        AttentionControlEdit.__init__(self, prompts, tokenizer, num_steps, cross_replace_steps, self_replace_steps, local_blend, device, LOW_RESOURCE)
        aligned = seq_aligner.get_replacement_mapper(prompts, tokenizer)     # [targets, 77, 77], rows sum to 1
        tokens = aligned.shape[-1]
        self.mapper = _expect(aligned, (len(prompts) - 1, tokens, tokens), "replacement mapper").to(device).to(dtype)

    def cross_edit(self) -> ops.CrossEdit:
        # CrossEdit derives the sparse (<= 8 source tokens per target token) form the kernel prefers from the dense mapper
        return ops.CrossEdit(ops.IEF_EDIT_REPLACE, self.mapper.shape[0], mapper=_f32(self.mapper))

    def replace_cross_attention(self, attn_base, att_replace):
        """Materialised form (reference :15-16); the registered closures use the fused equivalent described by cross_edit()."""
        return torch.einsum("hpw,bwn->bhpn", attn_base, self.mapper.to(attn_base.dtype))

    replace_cross_attention._ief_builtin = True

    def _retarget_tables(self, prompts, tokenizer) -> None:
        aligned = seq_aligner.get_replacement_mapper(prompts, tokenizer)
        self.mapper.copy_(_expect(aligned, tuple(self.mapper.shape), "replacement mapper"))
        if self._edit is None:
            return
        fresh = ops.CrossEdit(ops.IEF_EDIT_REPLACE, self.mapper.shape[0], mapper=_f32(self.mapper))
        if (fresh.mapper_nz_idx is None) != (self._edit.mapper_nz_idx is None):
            self._edit, self._tables_epoch = fresh, self._tables_epoch + 1     # sparse <-> dense form: other kernel flavour, new graphs
            return
        self._edit.mapper.copy_(fresh.mapper)
        if fresh.mapper_nz_idx is not None:
            self._edit.mapper_nz_idx.copy_(fresh.mapper_nz_idx)
            self._edit.mapper_nz_w.copy_(fresh.mapper_nz_w)


class AttentionRefine(AttentionControlEdit):
    """Prompt refinement: target tokens aligned to a source token take its probability (gather), new tokens keep their own."""

    def __init__(
        self,
        prompts,
        tokenizer,
        num_steps: int,
        cross_replace_steps: float,
        self_replace_steps: float,
        local_blend: Optional[LocalBlend] = None,
        device=_CUDA0,
        LOW_RESOURCE=False,
    ):
        AttentionControlEdit.__init__(self, prompts, tokenizer, num_steps, cross_replace_steps, self_replace_steps, local_blend, device, LOW_RESOURCE)
        gather_idx, keep = seq_aligner.get_refinement_mapper(prompts, tokenizer)   # int64 [targets, 77] (may hold -1), 0/1 [targets, 77]
        _expect(keep, gather_idx.shape, "refinement alphas")
        if int(gather_idx.min()) < -1 or int(gather_idx.max()) >= gather_idx.shape[-1]:
            raise ValueError("refinement mapper holds token positions outside [-1, max_length)")
        keep = keep.to(device)
        self.mapper = _expect(gather_idx, (len(prompts) - 1, gather_idx.shape[-1]), "refinement mapper").to(device)
        self.alphas = keep.reshape(keep.shape[0], 1, 1, keep.shape[1])

    def cross_edit(self) -> ops.CrossEdit:
        targets = self.mapper.shape[0]
        return ops.CrossEdit(ops.IEF_EDIT_REFINE, targets, mapper_idx=self.mapper.to(torch.int32).contiguous(),
                             refine_alpha=_f32(self.alphas.reshape(targets, -1)))

    def replace_cross_attention(self, attn_base, att_replace):
        """Materialised form (reference :28-31): gather along the token axis (a -1 entry reads the last token, as torch indexing does)."""
        gathered = attn_base[:, :, self.mapper].permute(2, 0, 1, 3)
        return gathered * self.alphas + att_replace * (1. - self.alphas)

    replace_cross_attention._ief_builtin = True

    def _retarget_tables(self, prompts, tokenizer) -> None:
        gather_idx, keep = seq_aligner.get_refinement_mapper(prompts, tokenizer)
        if int(gather_idx.min()) < -1 or int(gather_idx.max()) >= gather_idx.shape[-1]:
            raise ValueError("refinement mapper holds token positions outside [-1, max_length)")
        self.mapper.copy_(_expect(gather_idx, tuple(self.mapper.shape), "refinement mapper"))
        self.alphas.copy_(keep.reshape(self.alphas.shape))
        if self._edit is not None:
            self._edit.mapper_idx.copy_(self.mapper.to(torch.int32))
            self._edit.refine_alpha.copy_(self.alphas.reshape(self.mapper.shape[0], -1).to(torch.float32))


class AttentionReweight(AttentionControlEdit):
    """Per-token re-weighting with an equalizer, optionally on top of another controller's edit."""

    def __init__(
        self,
        prompts,
        tokenizer,
        num_steps: int,
        cross_replace_steps: float,
        self_replace_steps: float,
        equalizer,
        local_blend: Optional[LocalBlend] = None,
        controller: Optional[AttentionControlEdit] = None,
        device=_CUDA0,
        LOW_RESOURCE=False,
        dtype=torch.float32,
    ):
        AttentionControlEdit.__init__(self, prompts, tokenizer, num_steps, cross_replace_steps, self_replace_steps, local_blend, device, LOW_RESOURCE)
        self.prev_controller = controller
        self.equalizer = equalizer.to(device).to(dtype)

    def replace_cross_attention(self, attn_base, att_replace):
        """Materialised form (reference :42-46): the previous controller's edit, if any, scaled per token."""
        if self.prev_controller is not None:
            attn_base = self.prev_controller.replace_cross_attention(attn_base, att_replace)
        return attn_base[None, :, :, :] * self.equalizer[:, None, None, :]

    replace_cross_attention._ief_builtin = True

    def cross_edit(self) -> ops.CrossEdit:
        targets = self.batch_size - 1
        if self.equalizer.shape[0] != targets:
            raise ValueError(f"equalizer has {self.equalizer.shape[0]} rows but there are {targets} target prompts")
        inner = self.prev_controller.cross_edit() if self.prev_controller is not None else ops.CrossEdit(ops.IEF_EDIT_NONE, targets)
        return ops.CrossEdit(inner.mode, targets, mapper=inner.mapper, mapper_idx=inner.mapper_idx, refine_alpha=inner.refine_alpha,
                             equalizer=_f32(self.equalizer), mapper_nz_idx=inner.mapper_nz_idx, mapper_nz_w=inner.mapper_nz_w)
