"""torch-tensor front end of the C ABI: pointer/stride extraction and stream plumbing only.

Every function launches CUDA kernels from libief_b200.so on torch's current stream; tensors must
live on a CUDA device (there is deliberately no CPU implementation here — the CPU restatement
is `oracle/`, which only tests and bench.py's cpu_baseline leg may use).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _cabi
from ._cabi import (IEF_BF16, IEF_F16, IEF_F32, IEF_IMPL_AUTO, IEF_IMPL_MMA, IEF_IMPL_TCGEN05,
                    IEF_EDIT_NONE, IEF_EDIT_REPLACE, IEF_EDIT_REFINE)

_DTYPES = {torch.bfloat16: IEF_BF16, torch.float16: IEF_F16, torch.float32: IEF_F32}


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "image_editing_framework_b200 ops need CUDA tensors: the controlled-attention hot path has no CPU fallback "
                f"(got a tensor on {t.device})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _as4(t: torch.Tensor, heads: int) -> torch.Tensor:
    """View as [B, N, H, d]: either a [B, N, H*d] projection (head axis = stride-d view of the channels) or an explicit
    4-D view with arbitrary batch/token/head strides (e.g. the reference's head-major '(b h) n d' layout, permuted)."""
    if t.dim() == 3:
        if t.shape[2] % heads:
            raise ValueError(f"channel count {t.shape[2]} is not divisible by heads={heads}")
        t = t.unflatten(2, (heads, t.shape[2] // heads))
    if t.dim() != 4 or t.shape[2] != heads or t.stride(3) != 1:
        raise ValueError(f"expected [B, N, H*d] or [B, N, H, d] with contiguous channels, got shape {tuple(t.shape)} strides {t.stride()}")
    return t


def _t4(t: torch.Tensor, heads: int) -> _cabi.Tensor4:
    t = _as4(t, heads)
    return _cabi.Tensor4(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float, *,
              q_src: Optional[Sequence[int]] = None, k_src: Optional[Sequence[int]] = None, v_src: Optional[Sequence[int]] = None,
              k_src2: Optional[Sequence[int]] = None, v_src2: Optional[Sequence[int]] = None,
              impl: int = IEF_IMPL_AUTO, probs_out: Optional[torch.Tensor] = None, probs_accum: bool = False,
              probs_slot: Optional[Sequence[int]] = None, rows: Optional[Sequence[int]] = None,
              key_bias: Optional[torch.Tensor] = None, bias_sel: Optional[Sequence[int]] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """O[b] = softmax(scale * Q[q_src[b]] K[k_src[b]]^T + key_bias[bias_sel[b]]) V[v_src[b]]  (ief_attn_fwd).

    q: [B, Nq, H*d], k/v: [B, Nk, H*d], bf16 or fp16. Returns [B, Nq, H*d] in the same dtype.
    key_bias: fp32 [n_bias, Nk] per-key additive bias (masked MasaCtrl); bias_sel[b] < 0 leaves row b unbiased.
    """
    _require_cuda(q, k, v, probs_out, out, key_bias)
    if q.dtype not in (torch.bfloat16, torch.float16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError(f"attention needs matching bf16/fp16 q,k,v; got {q.dtype}, {k.dtype}, {v.dtype}")
    q4, k4, v4 = _as4(q, heads), _as4(k, heads), _as4(v, heads)
    B, Nq, _, d = q4.shape
    Nk = k4.shape[1]
    if k4.shape[3] != d or v4.shape[3] != d or v4.shape[1] != Nk or k4.shape[0] != B or v4.shape[0] != B:
        raise ValueError(f"inconsistent shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)} heads={heads}")
    if out is None:
        out = torch.empty((B, Nq, heads * d), dtype=q.dtype, device=q.device)
    p = _cabi.AttnParams()
    p.q, p.k, p.v, p.o = _t4(q4, heads), _t4(k4, heads), _t4(v4, heads), _t4(out, heads)
    p.dtype = _DTYPES[q.dtype]
    p.B, p.H, p.Nq, p.Nk, p.d = B, heads, Nq, Nk, d
    p.scale = float(scale)
    p.impl = int(impl)
    keep = [_cabi.i32_array(x) for x in (q_src, k_src, v_src, k_src2, v_src2, probs_slot, bias_sel)]
    if (key_bias is None) != (bias_sel is None):
        raise ValueError("key_bias and bias_sel go together")
    if key_bias is not None:
        if key_bias.dtype != torch.float32 or not key_bias.is_contiguous() or key_bias.dim() != 2 or key_bias.shape[1] != Nk:
            raise TypeError(f"key_bias must be a contiguous fp32 [n_bias, Nk={Nk}] tensor")
        p.key_bias, p.n_bias = key_bias.data_ptr(), key_bias.shape[0]
        p.bias_sel = C.cast(keep[6], C.POINTER(C.c_int32))
    for x in keep:
        if x is not None and len(x) != B:
            raise ValueError("per-row index arrays must have one entry per batch row")
    row_mask = None
    if rows is not None:  # only these batch rows are computed; the others' output is left untouched
        row_mask = (C.c_uint8 * B)(*[1 if i in set(rows) else 0 for i in range(B)])
        p.row_mask = C.cast(row_mask, C.POINTER(C.c_uint8))
    p.q_src, p.k_src, p.v_src, p.k_src2, p.v_src2, p.probs_slot = [
        C.cast(x, C.POINTER(C.c_int32)) if x is not None else None for x in keep[:6]]
    if probs_out is not None:
        if probs_out.dtype != torch.float32 or not probs_out.is_contiguous():
            raise TypeError("probs_out must be a contiguous fp32 tensor")
        p.probs_out = probs_out.data_ptr()
        p.probs_accum = 1 if probs_accum else 0
    with torch.cuda.device(q.device):
        ws_bytes = _cabi.lib().ief_attn_workspace_bytes(C.byref(p)) if impl != IEF_IMPL_MMA else 0
        if ws_bytes > 0:  # scratch for the key-norm pre-pass of the bf16 tcgen05 kernel (caching allocator: no cudaMalloc per call)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
            p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        _cabi.check("ief_attn_fwd", _cabi.lib().ief_attn_fwd(C.byref(p), _stream()))
    return out


def mask_blend(fg: torch.Tensor, bg: torch.Tensor, w: torch.Tensor, rows: Optional[Sequence[int]] = None) -> torch.Tensor:
    """In place fg[b, n, :] = fg[b, n, :] * w[n] + bg[b, n, :] * (1 - w[n]) for the batch rows in `rows` (ief_mask_blend)."""
    _require_cuda(fg, bg, w)
    if fg.dtype not in _DTYPES or bg.dtype != fg.dtype or fg.shape != bg.shape or fg.dim() != 3:
        raise TypeError(f"mask_blend needs two [B, N, C] tensors of one dtype, got {tuple(fg.shape)} {fg.dtype} / {tuple(bg.shape)} {bg.dtype}")
    if not fg.is_contiguous() or not bg.is_contiguous():
        raise TypeError("mask_blend needs contiguous tensors")
    B, N, Cc = fg.shape
    if w.dtype != torch.float32 or not w.is_contiguous() or w.numel() != N:
        raise TypeError(f"w must be a contiguous fp32 vector of {N} weights")
    row_mask = None
    if rows is not None:
        row_mask = (C.c_uint8 * B)(*[1 if i in set(rows) else 0 for i in range(B)])
    with torch.cuda.device(fg.device):
        _cabi.check("ief_mask_blend", _cabi.lib().ief_mask_blend(fg.data_ptr(), bg.data_ptr(), w.data_ptr(), _DTYPES[fg.dtype], B, N, Cc,
                                                                 C.cast(row_mask, C.POINTER(C.c_uint8)) if row_mask is not None else None, _stream()))
    return fg


class CrossEdit:
    """Device-side tables of one P2P cross-attention edit (built once per controller)."""

    MAX_NZ = 8  # kNZ in csrc/cross_attn.cu

    def __init__(self, mode: int, n_slots: int, mapper=None, mapper_idx=None, refine_alpha=None, equalizer=None,
                 mapper_nz_idx=None, mapper_nz_w=None):
        self.mode, self.n_slots = mode, n_slots
        self.mapper, self.mapper_idx, self.refine_alpha, self.equalizer = mapper, mapper_idx, refine_alpha, equalizer
        self.mapper_nz_idx, self.mapper_nz_w = mapper_nz_idx, mapper_nz_w
        if mode == IEF_EDIT_REPLACE and mapper is not None and mapper_nz_idx is None:
            self.mapper_nz_idx, self.mapper_nz_w = self.sparsify(mapper)

    @classmethod
    def sparsify(cls, mapper: torch.Tensor):
        """Sparse form of a [slots, Nk, Nk] replacement mapper: per target token the (<= 8) contributing source tokens in
        ascending order. Returns (None, None) when a column has more non-zeros (the kernel then multiplies the dense form)."""
        m = mapper.detach().float().cpu()
        slots, nk, _ = m.shape
        idx = torch.full((slots, nk, cls.MAX_NZ), -1, dtype=torch.int32)
        w = torch.zeros((slots, nk, cls.MAX_NZ), dtype=torch.float32)
        for s_ in range(slots):
            for n in range(nk):
                nz = torch.nonzero(m[s_, :, n], as_tuple=False).flatten()
                if nz.numel() > cls.MAX_NZ:
                    return None, None
                idx[s_, n, :nz.numel()] = nz.to(torch.int32)
                w[s_, n, :nz.numel()] = m[s_, nz, n]
        return idx.to(mapper.device).contiguous(), w.to(mapper.device).contiguous()


def cross_attention_edit(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float, *,
                         edit: Optional[CrossEdit] = None, step_alpha: Optional[torch.Tensor] = None,
                         base_row: Optional[Sequence[int]] = None, edit_slot: Optional[Sequence[int]] = None,
                         probs_out: Optional[torch.Tensor] = None, probs_accum: bool = False,
                         store_slot: Optional[Sequence[int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """<=80-key cross-attention with the P2P replace/refine/reweight edit fused (ief_cross_attn_edit_fwd)."""
    _require_cuda(q, k, v, probs_out, out, step_alpha)
    if q.dtype not in (torch.bfloat16, torch.float16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError(f"cross_attention_edit needs matching bf16/fp16 q,k,v; got {q.dtype}, {k.dtype}, {v.dtype}")
    q4, k4, v4 = _as4(q, heads), _as4(k, heads), _as4(v, heads)
    B, Nq, _, d = q4.shape
    Nk = k4.shape[1]
    if out is None:
        out = torch.empty((B, Nq, heads * d), dtype=q.dtype, device=q.device)
    p = _cabi.CrossParams()
    p.q, p.k, p.v, p.o = _t4(q4, heads), _t4(k4, heads), _t4(v4, heads), _t4(out, heads)
    p.dtype = _DTYPES[q.dtype]
    p.B, p.H, p.Nq, p.Nk, p.d = B, heads, Nq, Nk, d
    p.scale = float(scale)
    keep = [_cabi.i32_array(x) for x in (base_row, edit_slot, store_slot)]
    p.base_row, p.edit_slot, p.store_slot = [C.cast(x, C.POINTER(C.c_int32)) if x is not None else None for x in keep]
    if edit is not None:
        p.mode, p.n_slots = edit.mode, edit.n_slots
        for name in ("mapper", "mapper_nz_idx", "mapper_nz_w", "mapper_idx", "refine_alpha", "equalizer"):
            t = getattr(edit, name)
            if t is not None:
                _require_cuda(t)
                want = torch.int32 if name in ("mapper_idx", "mapper_nz_idx") else torch.float32
                if t.dtype != want or not t.is_contiguous():
                    raise TypeError(f"CrossEdit.{name} must be contiguous {want}")
                setattr(p, name, t.data_ptr())
        if step_alpha is not None:
            if step_alpha.dtype != torch.float32 or not step_alpha.is_contiguous() or step_alpha.numel() != edit.n_slots * Nk:
                raise TypeError("step_alpha must be contiguous fp32 [n_slots, Nk]")
            p.step_alpha = step_alpha.data_ptr()
    else:
        p.mode, p.n_slots = IEF_EDIT_NONE, 0
    if probs_out is not None:
        if probs_out.dtype != torch.float32 or not probs_out.is_contiguous():
            raise TypeError("probs_out must be a contiguous fp32 tensor")
        p.probs_out = probs_out.data_ptr()
        p.probs_accum = 1 if probs_accum else 0
    with torch.cuda.device(q.device):
        _cabi.check("ief_cross_attn_edit_fwd", _cabi.lib().ief_cross_attn_edit_fwd(C.byref(p), _stream()))
    return out


def cross_attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dout: torch.Tensor, heads: int, scale: float, *,
                             dprobs: Optional[torch.Tensor] = None, want_ds: bool = True):
    """Backward of plain <= 80-key cross-attention (ief_cross_attn_bwd): returns (dQ [B, N, H*d] in q's dtype, dS fp32
    [B*heads, N, Nk] or None). dprobs: gradient arriving directly on the probabilities, fp32 [B*heads, N, Nk]."""
    _require_cuda(q, k, v, dout, dprobs)
    if q.dtype not in (torch.bfloat16, torch.float16) or k.dtype != q.dtype or v.dtype != q.dtype or dout.dtype != q.dtype:
        raise TypeError(f"cross_attention_backward needs matching bf16/fp16 q,k,v,dout; got {q.dtype}, {k.dtype}, {v.dtype}, {dout.dtype}")
    q4, k4, v4, g4 = _as4(q, heads), _as4(k, heads), _as4(v, heads), _as4(dout, heads)
    B, Nq, _, d = q4.shape
    Nk = k4.shape[1]
    dq = torch.empty((B, Nq, heads * d), dtype=q.dtype, device=q.device)
    p = _cabi.CrossBwdParams()
    p.q, p.k, p.v, p.dout, p.dq = _t4(q4, heads), _t4(k4, heads), _t4(v4, heads), _t4(g4, heads), _t4(dq, heads)
    p.dtype = _DTYPES[q.dtype]
    p.B, p.H, p.Nq, p.Nk, p.d = B, heads, Nq, Nk, d
    p.scale = float(scale)
    if dprobs is not None:
        if dprobs.dtype != torch.float32 or not dprobs.is_contiguous() or dprobs.numel() != B * heads * Nq * Nk:
            raise TypeError("dprobs must be a contiguous fp32 [B*heads, Nq, Nk] tensor")
        p.dprobs = dprobs.data_ptr()
    ds = torch.empty((B * heads, Nq, Nk), dtype=torch.float32, device=q.device) if want_ds else None
    if ds is not None:
        p.ds_out = ds.data_ptr()
    with torch.cuda.device(q.device):
        _cabi.check("ief_cross_attn_bwd", _cabi.lib().ief_cross_attn_bwd(C.byref(p), _stream()))
    return dq, ds


def store_accumulate(dst: Sequence[torch.Tensor], src: Sequence[torch.Tensor]) -> None:
    """dst[i] += src[i] for all i in one launch (ief_store_accumulate)."""
    if len(dst) != len(src):
        raise ValueError("dst and src lists differ in length")
    for i in range(0, len(dst), 64):
        d, s = dst[i:i + 64], src[i:i + 64]
        n = len(d)
        if n == 0:
            return
        _require_cuda(*d, *s)
        for a, b in zip(d, s):
            if a.dtype != torch.float32 or b.dtype != torch.float32 or not a.is_contiguous() or not b.is_contiguous() or a.numel() != b.numel():
                raise TypeError("store_accumulate needs matching contiguous fp32 tensors")
        dp = (C.c_void_p * n)(*[t.data_ptr() for t in d])
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in s])
        ne = (C.c_int64 * n)(*[t.numel() for t in d])
        with torch.cuda.device(d[0].device):
            _cabi.check("ief_store_accumulate", _cabi.lib().ief_store_accumulate(dp, sp, ne, n, _stream()))


def local_blend(x_t: torch.Tensor, maps: Sequence[torch.Tensor], n_prompts: int, word_alpha: torch.Tensor, threshold: float,
                res: int = 16, return_mask: bool = False):
    """In-place LocalBlend on x_t (fp32 [n_prompts, C, H, W]) from stored cross maps [n_prompts*heads, res*res, words]."""
    _require_cuda(x_t, word_alpha, *maps)
    if x_t.dtype != torch.float32 or not x_t.is_contiguous():
        raise TypeError("x_t must be contiguous fp32")
    n_words = word_alpha.shape[-1]
    wa = word_alpha.reshape(n_prompts, n_words).to(torch.float32).contiguous()
    heads = []
    for m in maps:
        if m.dtype != torch.float32 or not m.is_contiguous() or m.shape[-1] != n_words or m.shape[-2] != res * res:
            raise TypeError(f"map of shape {tuple(m.shape)} is not contiguous fp32 [n_prompts*heads, {res * res}, {n_words}]")
        heads.append(m.numel() // (n_prompts * res * res * n_words))
    n = len(maps)
    p = _cabi.LocalBlendParams()
    mp = (C.c_void_p * n)(*[m.data_ptr() for m in maps])
    hp = (C.c_int32 * n)(*heads)
    p.maps, p.map_heads = C.cast(mp, C.POINTER(C.c_void_p)), C.cast(hp, C.POINTER(C.c_int32))
    p.n_maps, p.n_prompts, p.res, p.n_words = n, n_prompts, res, n_words
    p.word_alpha = wa.data_ptr()
    p.threshold = float(threshold)
    p.x_t = x_t.data_ptr()
    p.C, p.Hx, p.Wx = x_t.shape[1], x_t.shape[2], x_t.shape[3]
    work = torch.empty(n_prompts * res * res, dtype=torch.float32, device=x_t.device)
    p.workspace = work.data_ptr()
    mask = torch.empty((n_prompts, x_t.shape[2], x_t.shape[3]), dtype=torch.float32, device=x_t.device) if return_mask else None
    if mask is not None:
        p.mask_out = mask.data_ptr()
    with torch.cuda.device(x_t.device):
        _cabi.check("ief_local_blend", _cabi.lib().ief_local_blend(C.byref(p), _stream()))
    return (x_t, mask) if return_mask else x_t


def cfg_ddim_step(eps_uncond: torch.Tensor, eps_cond: Optional[torch.Tensor], x: torch.Tensor, guidance: float,
                  alpha_t: float, alpha_prev: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fused launch: eps = eu + g(ec-eu); x' = sqrt(a_prev) (x - sqrt(1-a_t) eps)/sqrt(a_t) + sqrt(1-a_prev) eps."""
    _require_cuda(eps_uncond, eps_cond, x, out)
    if x.dtype not in _DTYPES or eps_uncond.dtype != x.dtype or (eps_cond is not None and eps_cond.dtype != x.dtype):
        raise TypeError("cfg_ddim_step needs eps and x of one dtype (fp32, bf16 or fp16)")
    if not (eps_uncond.is_contiguous() and x.is_contiguous() and (eps_cond is None or eps_cond.is_contiguous())):
        raise TypeError("cfg_ddim_step needs contiguous tensors")
    if eps_uncond.numel() != x.numel() or (eps_cond is not None and eps_cond.numel() != x.numel()):
        raise ValueError("eps and x differ in size")
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check("ief_cfg_ddim_step", _cabi.lib().ief_cfg_ddim_step(
            eps_uncond.data_ptr(), eps_cond.data_ptr() if eps_cond is not None else None, x.data_ptr(), out.data_ptr(),
            x.numel(), _DTYPES[x.dtype], float(guidance), float(alpha_t), float(alpha_prev), _stream()))
    return out


def umma_probe(a: torch.Tensor, b: torch.Tensor, b_mn_major: bool, a_from_tmem: bool) -> torch.Tensor:
    """Diagnostic single-CTA tcgen05 GEMM: D[128,N] = A[128,K] @ (B^T if K-major else B)."""
    _require_cuda(a, b)
    K = a.shape[1]
    N = b.shape[1] if b_mn_major else b.shape[0]
    d = torch.empty((128, N), dtype=torch.float32, device=a.device)
    p = _cabi.UmmaProbeParams(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, int(b_mn_major), int(a_from_tmem), _DTYPES[a.dtype])
    with torch.cuda.device(a.device):
        _cabi.check("ief_umma_probe", _cabi.lib().ief_umma_probe(C.byref(p), _stream()))
    return d
