"""Token alignment for Prompt-to-Prompt edits (host side, integer work — outputs must be bit-exact).

Independent restatement of p2p/model/seq_aligner.py of the reference:
  get_refinement_mapper  :121-128  (Needleman-Wunsch global alignment, gap=0 match=1 mismatch=-1, :110;
                                    ties resolved left, then up, then diagonal, :70-75)
  get_replacement_mapper :189-195  (word-swap mapper with the 1/len(target) ratio branch, :169-172)
  get_word_inds          :131-149
  get_equalizer          :197-207
Only `.encode(str) -> [bos, ..., eos]` and `.decode([id]) -> str` of the tokenizer are used.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

_LEFT, _UP, _DIAG, _STOP = 1, 2, 3, 4  # traceback codes (same numbering as the reference, :53-58)


def _nw_traceback(x: Sequence[int], y: Sequence[int], gap: int = 0, match: int = 1, mismatch: int = -1) -> np.ndarray:
    """Traceback table of a global alignment of token sequences x (rows) and y (columns)."""
    nx, ny = len(x), len(y)
    score = np.zeros((nx + 1, ny + 1), dtype=np.int32)
    score[0, 1:] = gap * np.arange(1, ny + 1)
    score[1:, 0] = gap * np.arange(1, nx + 1)
    tb = np.zeros((nx + 1, ny + 1), dtype=np.int32)
    tb[0, 1:], tb[1:, 0], tb[0, 0] = _LEFT, _UP, _STOP
    for i in range(1, nx + 1):
        xi = x[i - 1]
        for j in range(1, ny + 1):
            cand_left = score[i, j - 1] + gap
            cand_up = score[i - 1, j] + gap
            cand_diag = score[i - 1, j - 1] + (match if xi == y[j - 1] else mismatch)
            best = max(cand_left, cand_up, cand_diag)
            score[i, j] = best
            # tie-break order is part of the contract: left wins over up wins over diagonal
            tb[i, j] = _LEFT if best == cand_left else (_UP if best == cand_up else _DIAG)
    return tb


def _y_to_x(x: Sequence[int], y: Sequence[int], tb: np.ndarray) -> torch.Tensor:
    """For every position of y (in order): the aligned position of x, or -1 for an insertion."""
    pairs: List[Tuple[int, int]] = []
    i, j = len(x), len(y)
    while i > 0 or j > 0:
        code = tb[i, j]
        if code == _DIAG:
            i, j = i - 1, j - 1
            pairs.append((j, i))
        elif code == _LEFT:
            j -= 1
            pairs.append((j, -1))
        elif code == _UP:
            i -= 1
        else:  # _STOP
            break
    pairs.reverse()
    return torch.tensor(pairs, dtype=torch.int64)


def get_mapper(x: str, y: str, tokenizer, max_len: int = 77):
    x_ids, y_ids = tokenizer.encode(x), tokenizer.encode(y)
    base = _y_to_x(x_ids, y_ids, _nw_traceback(x_ids, y_ids))
    n = base.shape[0]
    alphas = torch.ones(max_len)
    alphas[:n] = base[:, 1].ne(-1).float()
    mapper = torch.zeros(max_len, dtype=torch.int64)
    mapper[:n] = base[:, 1]
    mapper[n:] = len(y_ids) + torch.arange(max_len - len(y_ids))
    return mapper, alphas


def get_refinement_mapper(prompts: Sequence[str], tokenizer, max_len: int = 77):
    out = [get_mapper(prompts[0], p, tokenizer, max_len) for p in prompts[1:]]
    return torch.stack([m for m, _ in out]), torch.stack([a for _, a in out])


def get_word_inds(text: str, word_place: Union[int, str], tokenizer) -> np.ndarray:
    """Token positions (1-based past BOS) of a word given by index or by string."""
    words = text.split(" ")
    if type(word_place) is str:
        places = [i for i, w in enumerate(words) if w == word_place]
    elif type(word_place) is int:
        places = [word_place]
    else:
        places = word_place
    found: List[int] = []
    if len(places) > 0:
        pieces = [tokenizer.decode([t]).strip("#") for t in tokenizer.encode(text)][1:-1]
        consumed, w = 0, 0
        for pos, piece in enumerate(pieces):
            consumed += len(piece)
            if w in places:
                found.append(pos + 1)
            if consumed >= len(words[w]):
                w, consumed = w + 1, 0
    return np.array(found)


def get_replacement_mapper_(x: str, y: str, tokenizer, max_len: int = 77) -> torch.Tensor:
    wx, wy = x.split(" "), y.split(" ")
    if len(wx) != len(wy):
        raise ValueError(f"attention replacement edit can only be applied on prompts with the same length"
                         f" but prompt A has {len(wx)} words and prompt B has {len(wy)} words.")
    swapped = [i for i in range(len(wy)) if wy[i] != wx[i]]
    src = [get_word_inds(x, i, tokenizer) for i in swapped]
    tgt = [get_word_inds(y, i, tokenizer) for i in swapped]
    m = np.zeros((max_len, max_len))
    i = j = nxt = 0
    while i < max_len and j < max_len:
        if nxt < len(src) and src[nxt][0] == i:
            s, t = src[nxt], tgt[nxt]
            if len(s) == len(t):
                m[s, t] = 1
            else:
                for col in t:
                    m[s, col] = 1 / len(t)
            nxt += 1
            i, j = i + len(s), j + len(t)
        else:
            if nxt < len(src):
                m[i, j] = 1
            else:
                m[j, j] = 1
            i, j = i + 1, j + 1
    return torch.from_numpy(m).float()


def get_replacement_mapper(prompts: Sequence[str], tokenizer, max_len: int = 77) -> torch.Tensor:
    return torch.stack([get_replacement_mapper_(prompts[0], p, tokenizer, max_len) for p in prompts[1:]])


def get_equalizer(tokenizer, text: str, word_select, values) -> torch.Tensor:
    if type(word_select) is int or type(word_select) is str:
        word_select = (word_select,)
    eq = torch.ones(len(values), 77)
    vals = torch.tensor(values, dtype=torch.float32)
    for word in word_select:
        for i in get_word_inds(text, word, tokenizer):
            eq[:, i] = vals
    return eq
