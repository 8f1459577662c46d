"""AttentionReplace / AttentionRefine / AttentionReweight with the constructor signatures of
p2p/model/attention_control.py:8-46. Each class only prepares device tables; the arithmetic of
`replace_cross_attention` (einsum with the 77x77 mapper :16, gather + alpha blend :29-30, equaliser :45)
runs inside ief_cross_attn_edit_fwd.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from . import seq_aligner
from .attention_base import AttentionControlEdit
from .ptp_utils import LocalBlend


class AttentionReplace(AttentionControlEdit):

    def __init__(self, prompts, tokenizer, num_steps: int, cross_replace_steps: float, self_replace_steps: float,
                 local_blend: Optional[LocalBlend] = None, device=torch.device("cuda:0"), LOW_RESOURCE=False, dtype=torch.float32):
        super().__init__(prompts, tokenizer, num_steps, cross_replace_steps, self_replace_steps, local_blend, device, LOW_RESOURCE)
        self.mapper = seq_aligner.get_replacement_mapper(prompts, tokenizer).to(device).to(dtype)

    def cross_edit(self) -> ops.CrossEdit:
        return ops.CrossEdit(ops.IEF_EDIT_REPLACE, self.mapper.shape[0], mapper=self.mapper.to(torch.float32).contiguous())


class AttentionRefine(AttentionControlEdit):

    def __init__(self, prompts, tokenizer, num_steps: int, cross_replace_steps: float, self_replace_steps: float,
                 local_blend: Optional[LocalBlend] = None, device=torch.device("cuda:0"), LOW_RESOURCE=False):
        super().__init__(prompts, tokenizer, num_steps, cross_replace_steps, self_replace_steps, local_blend, device, LOW_RESOURCE)
        mapper, alphas = seq_aligner.get_refinement_mapper(prompts, tokenizer)
        self.mapper, alphas = mapper.to(device), alphas.to(device)
        self.alphas = alphas.reshape(alphas.shape[0], 1, 1, alphas.shape[1])

    def cross_edit(self) -> ops.CrossEdit:
        n = self.mapper.shape[0]
        return ops.CrossEdit(ops.IEF_EDIT_REFINE, n, mapper_idx=self.mapper.to(torch.int32).contiguous(),
                             refine_alpha=self.alphas.reshape(n, -1).to(torch.float32).contiguous())


class AttentionReweight(AttentionControlEdit):

    def __init__(self, prompts, tokenizer, num_steps: int, cross_replace_steps: float, self_replace_steps: float, equalizer,
                 local_blend: Optional[LocalBlend] = None, controller: Optional[AttentionControlEdit] = None,
                 device=torch.device("cuda:0"), LOW_RESOURCE=False, dtype=torch.float32):
        super().__init__(prompts, tokenizer, num_steps, cross_replace_steps, self_replace_steps, local_blend, device, LOW_RESOURCE)
        self.equalizer = equalizer.to(device).to(dtype)
        self.prev_controller = controller

    def cross_edit(self) -> ops.CrossEdit:
        n = self.batch_size - 1
        if self.equalizer.shape[0] != n:
            raise ValueError(f"equalizer has {self.equalizer.shape[0]} rows but there are {n} target prompts")
        if self.prev_controller is not None:
            e = self.prev_controller.cross_edit()
        else:
            e = ops.CrossEdit(ops.IEF_EDIT_NONE, n)
        return ops.CrossEdit(e.mode, n, mapper=e.mapper, mapper_idx=e.mapper_idx, refine_alpha=e.refine_alpha,
                             equalizer=self.equalizer.to(torch.float32).contiguous(), mapper_nz_idx=e.mapper_nz_idx, mapper_nz_w=e.mapper_nz_w)
