"""Prompt-to-Prompt controllers — same classes, constructors and public state as p2p/model/attention_base.py,
but each layer call is ONE fused kernel instead of "materialise softmax(QK^T), edit it, bmm".

Reference                                    here
  AttentionControl.__call__ :16-28            AttentionControl.attend(): same gating on cur_att_layer / LOW_RESOURCE,
                                              same counter arithmetic (_tick), cond-half-only editing
  AttentionStore :57-91                       maps written by the attention kernels' epilogues straight into
                                              `attention_store` (overwrite on the first step, += afterwards): there is
                                              no step_store and no python `+=` loop
  AttentionControlEdit.forward :113-125       per-row source indices (self) / ief_cross_attn_edit_fwd tables (cross)
  replace_self_attention :132-136             rows 1.. of the cond half read Q,K of the base row when N <= 16^2

The counters are advanced exactly like the reference so `cur_step` / `cur_att_layer` stay drop-in observable.
"""
from __future__ import annotations

import abc
from typing import Dict, List, Optional, Tuple, Union

import torch

from .. import ops
from . import ptp_utils
from .ptp_utils import LocalBlend

_STORE_MAX_TOKENS = 32 ** 2   # reference :66 "avoid memory overhead"
_SELF_REPLACE_MAX_TOKENS = 16 ** 2  # reference :133


class AttentionControl(abc.ABC):

    def __init__(self, LOW_RESOURCE):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0
        self.LOW_RESOURCE = LOW_RESOURCE

    # ---- counters (reference :16-28) ---------------------------------------------------------------------------
    @property
    def num_uncond_att_layers(self):
        return self.num_att_layers if self.LOW_RESOURCE else 0

    def _tick(self) -> None:
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers + self.num_uncond_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.between_steps()

    def _edited_rows(self, batch: int) -> Tuple[int, int]:
        """[first, last) UNet batch rows the controller acts on: everything in LOW_RESOURCE's cond pass,
        else the cond half `attn[h // 2:]` (reference :20-22)."""
        return (0, batch) if self.LOW_RESOURCE else (batch // 2, batch)

    # ---- fused entry point used by register_attention_control ---------------------------------------------------
    def attend(self, q, k, v, heads: int, scale: float, is_cross: bool, place_in_unet: str) -> torch.Tensor:
        """Equivalent of `bmm(self(softmax(scale q k^T), is_cross, place_in_unet), v)` on [B, N, H*d] projections."""
        if self._needs_probabilities():
            return self._attend_materialised(q, k, v, heads, scale, is_cross, place_in_unet)
        if self.cur_att_layer >= self.num_uncond_att_layers:
            out = self.fused_forward(q, k, v, heads, scale, is_cross, place_in_unet)
        else:
            out = _plain(q, k, v, heads, scale, is_cross)
        self._tick()
        return out

    def fused_forward(self, q, k, v, heads, scale, is_cross: bool, place_in_unet: str) -> torch.Tensor:
        raise NotImplementedError(f"{type(self).__name__} defines neither fused_forward() nor the reference's forward(attn, is_cross, place_in_unet)")

    # ---- the reference's materialised-probability interface (attention_base.py:16-32) -----------------------------
    # Controllers written against the reference implement `forward(attn, is_cross, place_in_unet)` on a [rows*heads, N, M] probability
    # tensor, or override replace_cross_attention / replace_self_attention of an edit controller. Such code cannot run inside a
    # fused kernel, so for those controllers (and only for them) the closure takes the compatibility route: the kernel EMITS the
    # probabilities (its map output), `controller(attn, is_cross, place)` runs exactly as in the reference, and P'V is one batched
    # GEMM. Cheap for the 77-key cross-attention it is meant for; for self-attention it costs what the reference costs (N x N fp32).
    def _needs_probabilities(self) -> bool:
        cls = type(self)
        cached = cls.__dict__.get("_ief_route")
        if cached is None:
            names = ("forward", "replace_cross_attention", "replace_self_attention")
            cached = any(hasattr(cls, n) and not getattr(getattr(cls, n), "_ief_builtin", False) for n in names)
            cls._ief_route = cached
        return cached

    _MATERIALISE_LIMIT_BYTES = 8 << 30

    def _attend_materialised(self, q, k, v, heads, scale, is_cross, place_in_unet):
        B, N, M = q.shape[0], q.shape[1], k.shape[1]
        if B * heads * N * M * 4 > self._MATERIALISE_LIMIT_BYTES:
            raise RuntimeError(f"{type(self).__name__} edits materialised probabilities; this layer's map would be {B * heads} x {N} x {M} fp32 "
                               f"({B * heads * N * M * 4 / 2 ** 30:.1f} GiB). Implement fused_forward() for layers of this size.")
        probs = torch.empty((B * heads, N, M), dtype=torch.float32, device=q.device)
        if is_cross and M <= 80:
            ops.cross_attention_edit(q, k, v, heads, scale, probs_out=probs)
        else:
            ops.attention(q, k, v, heads, scale, probs_out=probs)
        edited = self(probs, is_cross, place_in_unet)          # reference semantics, counters included
        v4 = v if v.dim() == 4 else v.view(B, M, heads, -1)
        out = torch.einsum("bhnm,bmhd->bnhd", edited.view(B, heads, N, M).to(v4.dtype), v4)
        return out.reshape(B, N, -1)

    def __call__(self, attn, is_cross: bool, place_in_unet: str):
        """controller(attn, is_cross, place_in_unet) on materialised probabilities, as the reference defines it (:16-28)."""
        if self.cur_att_layer >= self.num_uncond_att_layers:
            if self.LOW_RESOURCE:
                attn = self.forward(attn, is_cross, place_in_unet)
            else:
                half = attn.shape[0] // 2
                attn[half:] = self.forward(attn[half:], is_cross, place_in_unet)   # only the conditional rows are edited
        self._tick()
        return attn

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        raise NotImplementedError

    forward._ief_builtin = True

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return None

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    # ---- CUDA-graph replay protocol (graphs.GraphedUNet) ---------------------------------------------------------
    _graph_mode = False

    def graph_key(self):
        """Host state that shapes the launches of the next UNet forward; None = not replayable."""
        if self.LOW_RESOURCE or self.cur_att_layer != 0:
            return None  # two UNet passes per step / a forward in flight: run eagerly
        return ()

    def graph_prepare(self) -> None:
        return None

    def graph_advance(self) -> None:
        """Effect of one complete forward on the counters (the num_att_layers `_tick`s of :23-27)."""
        self.cur_att_layer = 0
        self.cur_step += 1
        self.between_steps()


def _plain(q, k, v, heads, scale, is_cross):
    if is_cross and k.shape[1] <= 80:
        return ops.cross_attention_edit(q, k, v, heads, scale)
    return ops.attention(q, k, v, heads, scale)


class EmptyControl(AttentionControl):

    def __init__(self, LOW_RESOURCE):
        super().__init__(LOW_RESOURCE)

    def fused_forward(self, q, k, v, heads, scale, is_cross, place_in_unet):
        return _plain(q, k, v, heads, scale, is_cross)

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        return attn

    forward._ief_builtin = True


class AttentionStore(AttentionControl):
    """`attention_store[key][i]` holds the SUM over finished steps of the cond-half maps of the i-th stored layer of
    that key, shape [rows*heads, N, M] fp32 — the tensors the reference ends up with (:67 alias + :81 `+=`)."""

    def __init__(self, LOW_RESOURCE):
        super().__init__(LOW_RESOURCE)
        self.step_store = self.get_empty_store()  # kept for API parity; the fused epilogue writes the store directly
        self.attention_store: Dict[str, List[torch.Tensor]] = {}
        self._slot = self._zero_slots()
        self._store_enabled = True
        # Opt-in (not reference behaviour): keep only the maps a LocalBlend reads (cross-attention at its resolution). The other
        # store entries become None placeholders so list positions stay the reference's; get_average_attention skips them.
        self.lean_store = False
        self._lean_tokens: Optional[int] = None
        self._spare: Dict[str, List[torch.Tensor]] = {}  # buffers of the store before the last reset(), reused in place

    @staticmethod
    def get_empty_store():
        return {"down_cross": [], "mid_cross": [], "up_cross": [], "down_self": [], "mid_self": [], "up_self": []}

    @staticmethod
    def _zero_slots():
        return {"down_cross": 0, "mid_cross": 0, "up_cross": 0, "down_self": 0, "mid_self": 0, "up_self": 0}

    def _store_target(self, key: str, rows: int, heads: int, n: int, m: int, device) -> Tuple[torch.Tensor, bool]:
        """Buffer for this layer's maps and whether the kernel must accumulate into it (every step but the first)."""
        if len(self.attention_store) == 0:
            self.attention_store = self.get_empty_store()
        bufs = self.attention_store[key]
        i = self._slot[key]
        self._slot[key] = i + 1
        if i == len(bufs):
            spare = self._spare.get(key, [])
            shape = (rows * heads, n, m)
            if i < len(spare) and spare[i] is not None and tuple(spare[i].shape) == shape and spare[i].device == torch.device(device):
                bufs.append(spare[i])  # same storage as before reset(): captured graphs keep pointing at live memory
            else:
                bufs.append(torch.empty(shape, dtype=torch.float32, device=device))
            return bufs[i], False
        return bufs[i], True

    def _stores(self, n_tokens: int) -> bool:
        return self._store_enabled and n_tokens <= _STORE_MAX_TOKENS

    def _wanted(self, n_tokens: int, is_cross: bool) -> bool:
        return not self.lean_store or self._lean_tokens is None or (is_cross and n_tokens == self._lean_tokens)

    def _skip_slot(self, key: str) -> None:
        """Lean store: this layer's maps are not kept; a None placeholder preserves the list positions LocalBlend slices by."""
        if len(self.attention_store) == 0:
            self.attention_store = self.get_empty_store()
        bufs = self.attention_store[key]
        i = self._slot[key]
        self._slot[key] = i + 1
        if i == len(bufs):
            bufs.append(None)

    def fused_forward(self, q, k, v, heads, scale, is_cross, place_in_unet):
        return self._attend_and_store(q, k, v, heads, scale, is_cross, place_in_unet)

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        """Materialised form (reference :64-68): maps of layers with at most 32^2 queries are summed into `attention_store`."""
        if self._stores(attn.shape[1]):
            key = f"{place_in_unet}_{'cross' if is_cross else 'self'}"
            if self._wanted(attn.shape[1], is_cross):
                buf, accum = self._store_target(key, attn.shape[0], 1, attn.shape[1], attn.shape[2], attn.device)
                buf.add_(attn) if accum else buf.copy_(attn)
            else:
                self._skip_slot(key)
        return attn

    forward._ief_builtin = True

    def _attend_and_store(self, q, k, v, heads, scale, is_cross, place_in_unet, *, self_src=None, cross_kw=None):
        B, N, M = q.shape[0], q.shape[1], k.shape[1]
        lo, hi = self._edited_rows(B)
        probs = accum = slots = None
        if self._stores(N):
            key = f"{place_in_unet}_{'cross' if is_cross else 'self'}"
            if self._wanted(N, is_cross):
                probs, accum = self._store_target(key, hi - lo, heads, N, M, q.device)
                slots = [(b - lo) if lo <= b < hi else -1 for b in range(B)]
            else:
                self._skip_slot(key)
        if is_cross:
            if M > 80:
                raise NotImplementedError(f"cross-attention with {M} keys: the fused edit kernel holds at most 80")
            return ops.cross_attention_edit(q, k, v, heads, scale, probs_out=probs, probs_accum=bool(accum), store_slot=slots,
                                            **(cross_kw or {}))
        return ops.attention(q, k, v, heads, scale, probs_out=probs, probs_accum=bool(accum), probs_slot=slots, **(self_src or {}))

    def between_steps(self):
        self._slot = self._zero_slots()

    def get_average_attention(self):
        return {key: [None if item is None else item / self.cur_step for item in self.attention_store[key]] for key in self.attention_store}

    def graph_key(self):
        base = super().graph_key()
        if base is None:
            return None
        # the first stored step allocates + overwrites the store, later ones accumulate into the same buffers
        return base + (("store", len(self.attention_store) != 0) if self._store_enabled else ("nostore",),)

    def graph_advance(self) -> None:
        if self._store_enabled and not self.attention_store and self._spare:
            # replay of a first stored step after reset(): the kernels overwrote the previous buffers in place; publish
            # them again the way _store_target does during an eager forward
            self.attention_store = {key: list(bufs) for key, bufs in self._spare.items()}
        super().graph_advance()

    def reset(self):
        super().reset()
        self.step_store = self.get_empty_store()
        # Under CUDA-graph replay the captured kernels keep writing into the previous store's buffers, so those are handed out again
        # (in place) for the next image. In eager mode the reference's behaviour holds: reset() starts a fresh store and whatever
        # the caller still holds of the old one stays untouched.
        self._spare = self.attention_store if (self.attention_store and self._graph_mode) else {}
        self.attention_store = {}
        self._slot = self._zero_slots()


class AttentionControlEdit(AttentionStore, abc.ABC):
    """As shipped, the reference derives this class from AttentionControl only, so `local_blend` crashes on the missing
    `attention_store`. Here it derives from AttentionStore the way upstream Prompt-to-Prompt does;
    maps are only stored when a LocalBlend needs them, so without one the behaviour (and memory) is the reference's."""

    def __init__(self, prompts, tokenizer, num_steps: int,
                 cross_replace_steps: Union[float, Tuple[float, float], Dict[str, Tuple[float, float]]],
                 self_replace_steps: Union[float, Tuple[float, float]],
                 local_blend: Optional[LocalBlend], device=torch.device("cuda:0"), LOW_RESOURCE=False):
        super().__init__(LOW_RESOURCE)
        self.batch_size = len(prompts)
        self._num_steps, self._cross_replace_steps = num_steps, cross_replace_steps
        self.cross_replace_alpha = ptp_utils.get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, tokenizer).to(device)
        if type(self_replace_steps) is float:
            self_replace_steps = 0, self_replace_steps
        self.num_self_replace = int(num_steps * self_replace_steps[0]), int(num_steps * self_replace_steps[1])
        self.local_blend = local_blend
        self._store_enabled = local_blend is not None
        self._lean_tokens = local_blend.res ** 2 if local_blend is not None else None
        self._device = torch.device(device)
        # [num_steps+1, n_targets, 77] fp32 contiguous: row `cur_step` is handed to the kernel as-is
        self._alpha_table = self.cross_replace_alpha.reshape(num_steps + 1, self.batch_size - 1, -1).to(torch.float32).contiguous()
        self._alpha_cur = self._alpha_table[0].clone()  # fixed buffer read by captured kernels (graph mode only)
        self._edit: Optional[ops.CrossEdit] = None

    @abc.abstractmethod
    def cross_edit(self) -> ops.CrossEdit:
        """Device tables describing replace_cross_attention for the fused kernel."""
        raise NotImplementedError

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        """Materialised form of the edit (reference :113-125) — what runs when a subclass overrides replace_cross_attention /
        replace_self_attention, or when the controller is called directly with a probability tensor."""
        AttentionStore.forward(self, attn, is_cross, place_in_unet)
        if is_cross or self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1]:
            per_prompt = attn.reshape(self.batch_size, attn.shape[0] // self.batch_size, *attn.shape[1:])
            base, others = per_prompt[0], per_prompt[1:]
            if is_cross:
                w = self.cross_replace_alpha[self.cur_step].to(attn.dtype)
                per_prompt[1:] = self.replace_cross_attention(base, others) * w + (1 - w) * others
            else:
                per_prompt[1:] = self.replace_self_attention(base, others)
            attn = per_prompt.reshape(attn.shape)
        return attn

    forward._ief_builtin = True

    def replace_self_attention(self, attn_base, att_replace):
        """Reference :132-136: the base prompt's self-attention maps for layers with at most 16^2 queries."""
        if att_replace.shape[2] <= _SELF_REPLACE_MAX_TOKENS:
            return attn_base.unsqueeze(0).expand(att_replace.shape[0], *attn_base.shape)
        return att_replace

    replace_self_attention._ief_builtin = True

    def replace_cross_attention(self, attn_base, att_replace):
        raise NotImplementedError

    replace_cross_attention._ief_builtin = True

    def _row_tables(self, batch: int):
        lo, hi = self._edited_rows(batch)
        if hi - lo != self.batch_size:
            raise ValueError(f"controller was built for {self.batch_size} prompts but the edited half of the UNet batch has {hi - lo} rows")
        base = [-1] * batch
        slot = [0] * batch
        for i in range(1, self.batch_size):
            base[lo + i], slot[lo + i] = lo, i - 1
        return lo, base, slot

    def fused_forward(self, q, k, v, heads, scale, is_cross, place_in_unet):
        B, N = q.shape[0], q.shape[1]
        if is_cross:
            if self._edit is None:
                self._edit = self.cross_edit()
            _, base, slot = self._row_tables(B)
            alpha = self._alpha_cur if self._graph_mode else self._alpha_table[self.cur_step]
            kw = dict(edit=self._edit, step_alpha=alpha, base_row=base, edit_slot=slot)
            return self._attend_and_store(q, k, v, heads, scale, True, place_in_unet, cross_kw=kw)
        src = None
        if self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1] and N <= _SELF_REPLACE_MAX_TOKENS:
            lo, base, _ = self._row_tables(B)
            qk = [base[b] if base[b] >= 0 else b for b in range(B)]
            src = dict(q_src=qk, k_src=qk)  # probabilities of the base prompt, values of the row itself
        return self._attend_and_store(q, k, v, heads, scale, False, place_in_unet, self_src=src)

    def step_callback(self, x_t):
        if self.local_blend is not None:
            x_t = self.local_blend(x_t, self.attention_store)
        return x_t

    # ---- extension (not in the reference): one controller, many prompt pairs ------------------------------------------
    _tables_epoch = 0

    def retarget(self, prompts, tokenizer) -> None:
        """Re-point this controller at another prompt set of the same size: the alpha schedule and the edit tables are recomputed
        on the host exactly as the constructor does and written INTO the existing device buffers, so CUDA graphs captured with this
        controller stay valid (the sweeps of */test.py build a new controller per image; with graphs that would re-capture every
        time). Falls back to fresh buffers — and a new graph key — when a table changes shape or form. Counters are reset."""
        if len(prompts) != self.batch_size:
            raise ValueError(f"retarget: controller was built for {self.batch_size} prompts, got {len(prompts)}")
        if self.local_blend is not None:
            raise NotImplementedError("retarget: a LocalBlend is tied to its prompts; build a new controller")
        alpha = ptp_utils.get_time_words_attention_alpha(prompts, self._num_steps, self._cross_replace_steps, tokenizer).to(self._device)
        self.cross_replace_alpha.copy_(alpha)
        self._alpha_table.copy_(alpha.reshape(self._alpha_table.shape).to(torch.float32))
        self._retarget_tables(prompts, tokenizer)
        self.reset()

    def _retarget_tables(self, prompts, tokenizer) -> None:
        raise NotImplementedError(f"{type(self).__name__} does not support retarget()")

    def graph_key(self):
        base = super().graph_key()
        if base is None or self.cur_step >= self._alpha_table.shape[0]:
            return None
        return base + (self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1], self._tables_epoch)

    def graph_prepare(self) -> None:
        if self._graph_mode and self.cur_step < self._alpha_table.shape[0]:
            self._alpha_cur.copy_(self._alpha_table[self.cur_step], non_blocking=True)
