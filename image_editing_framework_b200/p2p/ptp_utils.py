"""Step x word alpha schedule and LocalBlend (host side of p2p/model/ptp_utils.py).

  get_time_words_attention_alpha  reference :66-84   -> [num_steps+1, n_prompts-1, 1, 1, 77] of 0/1
  LocalBlend                      reference :6-32    -> mask from the five stored 16x16 cross maps, blend of latents;
                                  the arithmetic is one fused launch pair in libief_b200 (ief_local_blend).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch

from .. import ops
from .seq_aligner import get_word_inds  # re-exported: the reference defines the same helper in both modules

__all__ = ["LocalBlend", "get_word_inds", "update_alpha_time_word", "get_time_words_attention_alpha"]


class LocalBlend:
    """Same constructor and call signature as the reference class; `__call__` runs on the GPU store.

    The reference hard-codes 16x16 maps (:22-23): on SD-1.5/2.1 @512 that is down_cross[2:4] + up_cross[:3].
    `attention_store` is the *summed* store AttentionControlEdit.step_callback passes (attention_base.py:129).
    """

    def __init__(self, tokenizer, prompts: List[str], words, threshold: float = .3, device=torch.device("cuda:0"), MAX_NUM_WORDS: int = 77):
        alpha = torch.zeros(len(prompts), 1, 1, 1, 1, MAX_NUM_WORDS)
        for i, (prompt, words_) in enumerate(zip(prompts, words)):
            for word in ([words_] if type(words_) is str else words_):
                alpha[i, :, :, :, :, get_word_inds(prompt, word, tokenizer)] = 1
        self.alpha_layers = alpha.to(device)
        self.threshold = threshold
        self.MAX_NUM_WORDS = MAX_NUM_WORDS
        self.res = 16
        self.masks = None   # set to a list to have every call append its per-prompt masks [n_prompts, H, W] (tests / inspection)

    def __call__(self, x_t: torch.Tensor, attention_store: Dict[str, List[torch.Tensor]]) -> torch.Tensor:
        maps = attention_store["down_cross"][2:4] + attention_store["up_cross"][:3]
        n_prompts = self.alpha_layers.shape[0]
        for m in maps:
            if m.shape[1] != self.res * self.res:
                raise ValueError(f"LocalBlend expects {self.res}x{self.res} cross maps, got {m.shape[1]} tokens "
                                 "(the reference reshape would silently fold the difference into the head axis)")
        work = x_t.to(torch.float32, copy=True).contiguous()  # the reference returns a new tensor, input untouched
        alpha = self.alpha_layers.reshape(n_prompts, self.MAX_NUM_WORDS)
        if self.masks is not None:
            _, mask = ops.local_blend(work, maps, n_prompts, alpha, self.threshold, res=self.res, return_mask=True)
            self.masks.append(mask)
        else:
            ops.local_blend(work, maps, n_prompts, alpha, self.threshold, res=self.res)
        return work.to(x_t.dtype)


def update_alpha_time_word(alpha: torch.Tensor, bounds: Union[float, Tuple[float, float]], prompt_ind: int,
                           word_inds: Optional[torch.Tensor] = None) -> torch.Tensor:
    lo, hi = (0, bounds) if type(bounds) is float else bounds
    start, end = int(lo * alpha.shape[0]), int(hi * alpha.shape[0])
    if word_inds is None:
        word_inds = torch.arange(alpha.shape[2])
    alpha[:start, prompt_ind, word_inds] = 0
    alpha[start:end, prompt_ind, word_inds] = 1
    alpha[end:, prompt_ind, word_inds] = 0
    return alpha


def get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, tokenizer, max_num_words: int = 77) -> torch.Tensor:
    if type(cross_replace_steps) is not dict:
        cross_replace_steps = {"default_": cross_replace_steps}
    cross_replace_steps.setdefault("default_", (0., 1.))
    n_tgt = len(prompts) - 1
    table = torch.zeros(num_steps + 1, n_tgt, max_num_words)
    for i in range(n_tgt):
        table = update_alpha_time_word(table, cross_replace_steps["default_"], i)
    for word, bounds in cross_replace_steps.items():
        if word == "default_":
            continue
        for i in range(n_tgt):
            inds = get_word_inds(prompts[i + 1], word, tokenizer)
            if len(inds) > 0:
                table = update_alpha_time_word(table, bounds, i, inds)
    return table.reshape(num_steps + 1, n_tgt, 1, 1, max_num_words)
