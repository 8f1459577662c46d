"""Parity of every CUDA kernel (called through the C ABI) against the CPU oracle on seeded inputs.

Tolerance: north_star states 2e-2 max-abs for bf16/fp16 attention outputs relative to the fp32 reference; integer /
fp32-elementwise work (DDIM step, store accumulate) is checked bit-exact.
"""
import math
import os

import pytest
import torch

from image_editing_framework_b200 import ops, _cabi
from oracle import controlled_attention as orc

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _qkv(B, N, M, H, d, seed, dtype=torch.bfloat16, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, N, H * d, generator=g) * spread).to(dtype)
    k = (torch.randn(B, M, H * d, generator=g) * spread).to(dtype)
    v = torch.randn(B, M, H * d, generator=g).to(dtype)
    return q, k, v


# ---------------------------------------------------------------------------------------------------- UMMA probe
@pytest.mark.parametrize("N,K,b_mn,a_tmem", [
    (128, 64, False, False), (128, 48, False, False), (128, 128, False, False), (64, 16, False, False),
    (64, 128, True, False), (48, 128, True, False), (64, 128, True, True), (48, 128, True, True),
    (80, 128, True, True), (128, 128, True, True), (160, 128, True, True), (192, 64, True, True),
])
def test_umma_probe(cuda, N, K, b_mn, a_tmem):
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g).to(torch.bfloat16)
    b = (torch.randn(K, N, generator=g) if b_mn else torch.randn(N, K, generator=g)).to(torch.bfloat16)
    want = a.float() @ (b.float() if b_mn else b.float().t())
    got = ops.umma_probe(a.to(cuda), b.to(cuda), b_mn, a_tmem).cpu()
    torch.cuda.synchronize()
    assert torch.allclose(got, want, atol=1e-2, rtol=1e-3), f"max err {(got - want).abs().max().item()}"


# ---------------------------------------------------------------------------------------------------- self attention
SHAPES_MMA = [  # B, H, Nq, Nk, d
    (4, 2, 256, 256, 40), (2, 2, 64, 64, 160), (2, 3, 300, 300, 64), (4, 2, 100, 77, 80), (1, 1, 17, 5, 8), (2, 2, 128, 130, 48),
]
SHAPES_TC = [
    (2, 8, 1024, 1024, 80), (2, 8, 4096, 4096, 40), (2, 5, 576, 576, 64), (2, 2, 1024, 384, 64), (2, 2, 256, 256, 160),
    (4, 2, 700, 333, 40), (1, 1, 128, 128, 64), (2, 2, 130, 77, 48),
]


def _check_attn(cuda, shape, impl, dtype=torch.bfloat16, src=None, seed=0):
    B, H, N, M, d = shape
    q, k, v = _qkv(B, N, M, H, d, seed, dtype)
    scale = d ** -0.5
    kw = src or {}
    want = orc.indexed_attention(q, k, v, H, scale, **kw)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, impl=impl, **kw)
    torch.cuda.synchronize()
    err = (got.float().cpu() - want).abs().max().item()
    assert err < TOL, f"max abs err {err}"
    return err


@pytest.mark.parametrize("shape", SHAPES_MMA)
def test_attn_mma(cuda, shape):
    _check_attn(cuda, shape, ops.IEF_IMPL_MMA)
    assert _cabi.last_attn_impl() == "mma"


@pytest.mark.parametrize("shape", SHAPES_TC)
def test_attn_tcgen05(cuda, shape):
    _check_attn(cuda, shape, ops.IEF_IMPL_TCGEN05)
    assert _cabi.last_attn_impl() == "tcgen05"


@pytest.mark.parametrize("impl", [ops.IEF_IMPL_MMA, ops.IEF_IMPL_TCGEN05])
def test_attn_fp16(cuda, impl):
    _check_attn(cuda, (2, 4, 512, 512, 64), impl, dtype=torch.float16)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [
    (2, 4, 512, 384, 16), (2, 2, 640, 640, 32), (1, 2, 600, 300, 8), (2, 2, 768, 768, 48), (2, 2, 768, 768, 56), (1, 8, 4096, 4096, 40),
    (3, 2, 1024, 77, 40),
])
def test_attn_tcgen05_row_sum_mma_variants(cuda, shape, dtype):
    """Generation 3b keeps the row sums in the 16 accumulator columns after O when head_dim <= 48 (column offset 16 / 32 / 48 for
    head dims 8-16 / 17-32 / 33-48, ones tile in the kernel's dtype); head_dim 56 and 64 take the variant with register sums.
    Both launch flavours (256-row pair CTAs, split-KV CTAs) are hit: small grids go split, B=1 x 8 heads x 4096 goes pair."""
    _check_attn(cuda, shape, ops.IEF_IMPL_TCGEN05, dtype=dtype, seed=5)


@pytest.mark.parametrize("impl", [ops.IEF_IMPL_MMA, ops.IEF_IMPL_TCGEN05])
@pytest.mark.parametrize("name,src", [
    ("p2p_self_replace", dict(q_src=[0, 1, 2, 2], k_src=[0, 1, 2, 2])),
    ("masactrl", dict(k_src=[0, 0, 2, 2], v_src=[0, 0, 2, 2])),
    ("pnp", dict(q_src=[0, 2, 2, 2], k_src=[0, 2, 2, 2])),
    ("union", dict(k_src=[0, 0, 2, 2], v_src=[0, 0, 2, 2], k_src2=[0, 1, 2, 3], v_src2=[0, 1, 2, 3])),
])
def test_attn_row_sources(cuda, impl, name, src):
    _check_attn(cuda, (4, 2, 640, 640, 40), impl, src=src, seed=3)


def test_attn_masactrl_matches_reference_formulation(cuda):
    """The stacked-queries form of masactrl/model/attention_control.py:37-68 equals per-row source indices."""
    B, H, N, d = 4, 8, 1024, 80
    q, k, v = _qkv(B, N, N, H, d, 7)
    want = orc.masactrl_mutual(q, k, v, H, d ** -0.5)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, d ** -0.5, k_src=[0, 0, 2, 2], v_src=[0, 0, 2, 2])
    assert (got.float().cpu() - want).abs().max().item() < TOL


def test_attn_large_logits_lazy_rescale(cuda):
    """Peaked scores exercise the lazy O-rescale branch of the tcgen05 kernel (threshold 2^8)."""
    _check_attn(cuda, (1, 2, 512, 2048, 64), ops.IEF_IMPL_TCGEN05, seed=11, src=None)
    B, H, N, d = 1, 2, 512, 64
    q, k, v = _qkv(B, N, 2048, H, d, 5, spread=4.0)
    want = orc.indexed_attention(q, k, v, H, d ** -0.5)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, d ** -0.5, impl=ops.IEF_IMPL_TCGEN05)
    assert (got.float().cpu() - want).abs().max().item() < TOL
    # head_dim 40: the row sums live in the accumulator (row-sum MMA) and must be rescaled together with O; keys sorted so that
    # the row maxima keep growing from tile to tile
    for dt in (torch.bfloat16, torch.float16):
        q, k, v = _qkv(1, 640, 2048, 2, 40, 6, dtype=dt, spread=3.0)
        order = (k.float()[0] @ q.float()[0].mean(0)).argsort()
        k, v = k[:, order].contiguous(), v[:, order].contiguous()
        want = orc.indexed_attention(q, k, v, 2, 40 ** -0.5)
        got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), 2, 40 ** -0.5, impl=ops.IEF_IMPL_TCGEN05)
        assert (got.float().cpu() - want).abs().max().item() < TOL


@pytest.mark.parametrize("case", ["plain", "late_large_keys", "huge_norms", "union"])
def test_attn_tcgen05_max_skipping_guard(cuda, case):
    """bf16, 48 < head_dim <= 64, >= 24 key tiles: the kernel runs a key-norm pre-pass and skips the running-maximum pass on tiles
    whose Cauchy-Schwarz score bound is harmless. The softmax must stay the same: (a) ordinary data, (b) later tiles with far
    larger scores than the first tile (the stale maximum is kept, probabilities grow up to ~2^40), (c) norms so large that
    the bound fails and every tile takes the exact path, (d) two key/value blocks."""
    B, H, N, M, d = 2, 2, 640, 3200, 64
    q, k, v = _qkv(B, N, M, H, d, 31, spread=6.0 if case == "huge_norms" else 1.0)
    if case == "late_large_keys":
        k = k.clone()
        k[:, 1500:] = (k[:, 1500:].float() * 6).to(k.dtype)
    kw = dict(k_src=[0, 0], v_src=[0, 0], k_src2=[0, 1], v_src2=[0, 1]) if case == "union" else {}
    if case == "union":
        q, k, v = q[:, :, :], k[:, :1600].contiguous(), v[:, :1600].contiguous()
    scale = d ** -0.5
    want = orc.indexed_attention(q, k, v, H, scale, **kw)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, impl=ops.IEF_IMPL_TCGEN05, **kw)
    torch.cuda.synchronize()
    assert _cabi.last_attn_impl() == "tcgen05"
    assert torch.isfinite(got).all()
    err = (got.float().cpu() - want).abs().max().item()
    assert err < TOL, f"{case}: max abs err {err}"


@pytest.mark.parametrize("shape", [
    (2, 2, 640, 3200, 64),     # 128-row split-KV CTAs, row sums in registers
    (2, 2, 640, 1100, 40),     # split-KV, odd number of key tiles (stream 1 one step short), row-sum MMA
    (2, 8, 2560, 1024, 40),    # 256-row pair CTAs (+ hybrid remainder), row-sum MMA
    (2, 10, 2304, 2304, 64),   # pair CTAs, row sums in registers
    (2, 10, 4096, 1024, 64),   # persistent pair CTAs (320 items on 148 SMs: up to three per CTA, flagged ones repeated last), row sums in registers
])
@pytest.mark.parametrize("case", ["plain", "overflow", "underflow", "one_batch_row", "few_query_rows", "union"])
def test_attn_tcgen05_unshifted_softmax_second_pass(cuda, shape, case):
    """bf16: the kernel first runs its key loop WITHOUT a running maximum (P = exp2(scaled score), attn_tc3 MAXMODE 2) and
    repeats it with the exact online softmax inside the same CTA when a row sum leaves (2^-100, 2^100). Same softmax either way:
    (a) ordinary data (first pass only), (b) scores far above 2^127 -> inf row sums, (c) every score of some rows below -200
    -> zero row sums, (d) only one batch row overflows (other CTAs keep their first pass), (e) a handful of peaked query rows,
    (f) two key/value blocks with overflow."""
    B, H, N, M, d = shape
    q, k, v = _qkv(B, N, M, H, d, 57 + N)
    q, k = q.float(), k.float()
    if case in ("overflow", "union"):
        q, k = q * 7, k * 7
    elif case == "underflow":
        q[..., 0::d], k[..., 0::d] = 48.0, -48.0        # channel 0 of every head: -2304 * scale on every score
    elif case == "one_batch_row":
        q[1], k[1] = q[1] * 7, k[1] * 7
    elif case == "few_query_rows":
        q[:, 5::97] *= 60
    q, k = q.to(torch.bfloat16), k.to(torch.bfloat16)
    kw = {}
    if case == "union":
        kw = dict(k_src=[0] * B, v_src=[0] * B, k_src2=list(range(B)), v_src2=list(range(B)))
    scale = d ** -0.5
    want = orc.indexed_attention(q, k, v, H, scale, **kw)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, impl=ops.IEF_IMPL_TCGEN05, **kw)
    torch.cuda.synchronize()
    assert _cabi.last_attn_impl() == "tcgen05"
    assert torch.isfinite(got).all(), f"{case}: non-finite output"
    err = (got.float().cpu() - want).abs().max().item()
    assert err < TOL, f"{case} {shape}: max abs err {err}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("d", [40, 64])
def test_attn_tcgen05_persistent_ctas_with_masked_rows(cuda, d, dtype):
    """More 256-row work items than SMs and an even number of key tiles: one CTA per SM walks several items (attn_tc3, PERSIST).
    Batch rows left out with `rows=` are skipped by every role of a CTA alike, rows of `out` that are not computed stay untouched,
    and an overflowing batch row between clean ones sends only its own items through the exact repeat after the CTA's last item
    (bf16; fp16 runs the exact loop on every item)."""
    B, H, N, M = 6, 8, 2048, 1024
    q, k, v = _qkv(B, N, M, H, d, 77 + d)
    q, k = q.float(), k.float()
    q[4], k[4] = q[4] * 7, k[4] * 7          # row 4: unshifted row sums overflow
    q, k, v = q.to(dtype), k.to(dtype), v.to(dtype)
    src = [0, 0, 2, 2, 4, 4]                  # MasaCtrl-style: odd rows read the K, V of the even row before them
    scale = d ** -0.5
    want = orc.indexed_attention(q, k, v, H, scale, k_src=src, v_src=src)
    rows = [0, 1, 3, 4, 5]
    out = torch.full((B, N, H * d), 123.0, dtype=dtype, device=cuda)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, impl=ops.IEF_IMPL_TCGEN05, k_src=src, v_src=src, rows=rows, out=out)
    torch.cuda.synchronize()
    assert _cabi.last_attn_impl() == "tcgen05"
    got = got.float().cpu()
    assert torch.isfinite(got).all()
    assert (got[2] == 123.0).all(), "a row outside `rows` was written"
    err = (got[rows] - want[rows]).abs().max().item()
    assert err < TOL, f"max abs err {err}"


def test_attn_auto_dispatch(cuda):
    q, k, v = _qkv(2, 4096, 4096, 8, 40, 0)
    ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), 8, 40 ** -0.5)
    assert _cabi.last_attn_impl() == "tcgen05"
    q, k, v = _qkv(2, 64, 64, 8, 160, 0)
    ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), 8, 160 ** -0.5)
    assert _cabi.last_attn_impl() == "mma"


def test_attn_head_major_strided_view(cuda):
    """The reference's '(b h) n d' layout is consumed through strides, without a copy."""
    B, H, N, d = 2, 4, 512, 64
    q, k, v = _qkv(B, N, N, H, d, 1)
    want = orc.plain_attention(q, k, v, H, d ** -0.5)
    hm = [orc.head_to_batch(t, H).to(cuda).contiguous().view(B, H, N, d).permute(0, 2, 1, 3) for t in (q, k, v)]
    for impl in (ops.IEF_IMPL_MMA, ops.IEF_IMPL_TCGEN05):
        got = ops.attention(hm[0], hm[1], hm[2], H, d ** -0.5, impl=impl)
        assert (got.float().cpu() - want).abs().max().item() < TOL


@pytest.mark.parametrize("shape", [(4, 2, 256, 256, 40), (2, 2, 1024, 1024, 80), (2, 3, 100, 77, 64)])
def test_attn_probs_out(cuda, shape):
    B, H, N, M, d = shape
    q, k, v = _qkv(B, N, M, H, d, 2)
    scale = d ** -0.5
    want_p = orc.attention_probs(q, k, H, scale)
    want_o = orc.apply_probs(want_p, v, H)
    lo = B // 2
    probs = torch.zeros((B - lo) * H, N, M, device=cuda)
    slots = [b - lo if b >= lo else -1 for b in range(B)]
    out = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, probs_out=probs, probs_slot=slots)
    assert (out.float().cpu() - want_o).abs().max().item() < TOL
    assert (probs.cpu() - want_p[lo * H:]).abs().max().item() < 5e-3
    ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, probs_out=probs, probs_accum=True, probs_slot=slots)
    assert (probs.cpu() - 2 * want_p[lo * H:]).abs().max().item() < 1e-2


@pytest.mark.parametrize("shape", [(4, 2, 256, 256, 40), (4, 2, 1024, 1024, 80), (4, 2, 100, 100, 64), (4, 8, 4096, 4096, 40)])
def test_attn_key_bias_masked_masactrl(cuda, shape):
    """softmax(scale QK^T + key_bias[bias_sel[b]]): fore-/background key masks of the masked MasaCtrl variants, including a
    row set whose keys are ALL masked (the reference then degenerates to the uniform average — finfo.min absorbs the scores)."""
    B, H, N, M, d = shape
    q, k, v = _qkv(B, N, M, H, d, 21)
    g = torch.Generator().manual_seed(3)
    m = (torch.rand(M, generator=g) > 0.6).float()
    fmin = torch.finfo(torch.float32).min
    bias = torch.stack([m.masked_fill(m == 0, fmin), m.masked_fill(m == 1, fmin), torch.full((M,), fmin), 0.5 * torch.randn(M, generator=g)])
    sel = [-1, 0, 2, 1] if N <= 256 else [3, 0, -1, 1]
    src = [0, 0, 2, 2]
    scale = d ** -0.5
    full = torch.stack([bias[s_] if s_ >= 0 else torch.zeros(M) for s_ in sel])
    p = orc.attention_probs(q, k[src], H, scale, key_bias=full)
    want = orc.apply_probs(p, v[src], H)
    impls = [ops.IEF_IMPL_MMA, ops.IEF_IMPL_AUTO] + ([ops.IEF_IMPL_TCGEN05] if d <= 64 else [])
    for impl in impls:
        got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, k_src=src, v_src=src, key_bias=bias.to(cuda), bias_sel=sel, impl=impl)
        torch.cuda.synchronize()
        if impl == ops.IEF_IMPL_TCGEN05 or (impl == ops.IEF_IMPL_AUTO and d <= 64 and N >= 512):
            assert _cabi.last_attn_impl() == "tcgen05"   # head_dim <= 64: the biased variant of the third-generation kernel
        elif impl == ops.IEF_IMPL_MMA or d > 64:
            assert _cabi.last_attn_impl() == "mma"
        err = (got.float().cpu() - want).abs().max().item()
        assert err < TOL, f"impl {impl}: max abs err {err}"
        if 2 in sel:  # all keys masked -> uniform average of V
            b = sel.index(2)
            assert (got[b].float().cpu() - v[src[b]].float().mean(0, keepdim=True)).abs().max().item() < TOL
    if d > 64:
        with pytest.raises(_cabi.IefError):
            ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, key_bias=bias.to(cuda), bias_sel=sel, impl=ops.IEF_IMPL_TCGEN05)
    with pytest.raises(_cabi.IefError):
        ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, key_bias=bias.to(cuda), bias_sel=[0, 1, 2, 7])
    with pytest.raises(_cabi.IefError):
        ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, key_bias=bias.to(cuda), bias_sel=sel, k_src2=src, v_src2=src)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_mask_blend(cuda, dtype):
    g = torch.Generator().manual_seed(4)
    B, N, Cc = 4, 1024, 320
    fg, bg = torch.randn(B, N, Cc, generator=g).to(dtype), torch.randn(B, N, Cc, generator=g).to(dtype)
    w = torch.rand(N, generator=g)
    w[::3] = 1.0
    w[1::3] = 0.0
    want = fg.clone()
    wf = w.reshape(1, N, 1)
    want[[1, 3]] = (fg[[1, 3]].float() * wf + bg[[1, 3]].float() * (1 - wf)).to(dtype)  # reference :176-177 (fp32 there)
    got = ops.mask_blend(fg.to(cuda), bg.to(cuda), w.to(cuda), rows=[1, 3])
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), want), "mask blend must be bit-exact (three separately rounded fp32 ops, one final rounding)"
    with pytest.raises(TypeError):
        ops.mask_blend(fg.to(cuda), bg.to(cuda)[:, :, :8], w.to(cuda))


def test_attn_rejects_bad_arguments(cuda):
    q, k, v = _qkv(2, 64, 64, 2, 36, 0)  # head_dim 36 is not a multiple of 8
    with pytest.raises(_cabi.IefError):
        ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), 2, 0.1)
    q, k, v = _qkv(2, 64, 64, 2, 40, 0)
    with pytest.raises(_cabi.IefError):
        ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), 2, 0.1, k_src=[0, 5])
    with pytest.raises(RuntimeError):
        ops.attention(q, k, v, 2, 0.1)  # CPU tensors: no fallback


# ---------------------------------------------------------------------------------------------------- cross attention edit
def _edit_tables(n_tgt, M, seed, mode):
    g = torch.Generator().manual_seed(seed)
    alpha = (torch.rand(n_tgt, M, generator=g) > 0.3).float()
    if mode in ("replace", "replace_dense"):
        mapper = torch.eye(M).repeat(n_tgt, 1, 1)
        mapper[:, 3, 3] = 0
        mapper[:, 3, 4] = 0.5
        mapper[:, 3, 5] = 0.5
        rows = 4 if mode == "replace" else 20   # 20 dense rows: > 8 source tokens per target token -> the dense-mapper flavour
        mapper[:, 10:10 + rows] = torch.rand(n_tgt, rows, M, generator=g).softmax(-1)
        return dict(mapper=mapper), alpha
    if mode == "refine":
        idx = torch.arange(M).repeat(n_tgt, 1)
        idx[:, 5:20] = torch.arange(4, 19)
        idx[:, 7] = -1
        idx[:, 30] = -1
        ra = torch.ones(n_tgt, M)
        ra[:, 7] = 0
        ra[:, 30] = 0
        return dict(mapper=idx, refine_alphas=ra.reshape(n_tgt, 1, 1, M)), alpha
    return {}, alpha


@pytest.mark.parametrize("mode", ["replace", "replace_dense", "refine", "none"])
@pytest.mark.parametrize("equalize", [False, True])
@pytest.mark.parametrize("shape", [(4, 8, 1024, 77, 80), (6, 2, 300, 77, 40), (4, 8, 4096, 77, 40)])
def test_cross_attention_edit(cuda, mode, equalize, shape):
    B, H, N, M, d = shape
    n_prompts = B // 2
    n_tgt = n_prompts - 1
    q, k, v = _qkv(B, N, M, H, d, 4)
    scale = d ** -0.5
    tables, alpha = _edit_tables(n_tgt, M, 9, mode)
    eq = None
    if equalize:
        eq = torch.ones(n_tgt, M)
        eq[:, 2] = 4.0
        eq[:, 6] = -1.5
    alpha_table = alpha.reshape(1, n_tgt, 1, 1, M)
    probs = orc.attention_probs(q, k, H, scale)
    edited = orc.p2p_edit_probs(probs, H, n_prompts, True, 0, mode=mode.split("_")[0], alpha_table=alpha_table, equalizer=eq, **tables)
    want_o = orc.apply_probs(edited, v, H)
    dev = {}
    if mode.startswith("replace"):
        dev["mapper"] = tables["mapper"].to(cuda).contiguous()
    if mode == "refine":
        dev["mapper_idx"] = tables["mapper"].to(torch.int32).to(cuda).contiguous()
        dev["refine_alpha"] = tables["refine_alphas"].reshape(n_tgt, M).to(cuda).contiguous()
    if eq is not None:
        dev["equalizer"] = eq.to(cuda).contiguous()
    edit = ops.CrossEdit({"replace": ops.IEF_EDIT_REPLACE, "refine": ops.IEF_EDIT_REFINE, "none": ops.IEF_EDIT_NONE}[mode.split("_")[0]], n_tgt, **dev)
    if mode.startswith("replace"):
        assert (edit.mapper_nz_idx is None) == (mode == "replace_dense")   # sparse form unless a column has more than 8 non-zeros
    lo = B // 2
    base = [-1] * B
    slot = [0] * B
    for i in range(1, n_prompts):
        base[lo + i], slot[lo + i] = lo, i - 1
    store = torch.zeros(n_prompts * H, N, M, device=cuda)
    sslot = [b - lo if b >= lo else -1 for b in range(B)]
    got = ops.cross_attention_edit(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, edit=edit, step_alpha=alpha.to(cuda).contiguous(),
                                   base_row=base, edit_slot=slot, probs_out=store, store_slot=sslot)
    torch.cuda.synchronize()
    # which kernel served it: the tensor-pipe edit kernel from 128 queries on, unless the mapper only exists in its dense form
    on_tensor_pipe = N >= 128 and mode != "replace_dense" and os.environ.get("IEF_CROSS_TC", "1") != "0" and os.environ.get("IEF_CROSS_TC_EDIT", "1") != "0"
    assert _cabi.last_cross_impl() == ("tcgen05-edit" if on_tensor_pipe else "mma")
    assert (got.float().cpu() - want_o).abs().max().item() < TOL * (4 if equalize else 1)
    assert (store.cpu() - edited[lo * H:]).abs().max().item() < 5e-3 * (4 if equalize else 1)
    # the same call without the map output and accumulating into a pre-filled store (other template flavours of the kernel)
    got2 = ops.cross_attention_edit(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, edit=edit, step_alpha=alpha.to(cuda).contiguous(),
                                    base_row=base, edit_slot=slot)
    assert (got2.float() - got.float()).abs().max().item() < 1e-2
    ops.cross_attention_edit(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, edit=edit, step_alpha=alpha.to(cuda).contiguous(),
                             base_row=base, edit_slot=slot, probs_out=store, probs_accum=True, store_slot=sslot)
    torch.cuda.synchronize()
    assert (store.cpu() - 2 * edited[lo * H:]).abs().max().item() < 1e-2 * (4 if equalize else 1)


def test_cross_attention_plain_matches_self_kernel(cuda):
    q, k, v = _qkv(2, 1024, 77, 8, 80, 12)
    want = orc.plain_attention(q, k, v, 8, 80 ** -0.5)
    got = ops.cross_attention_edit(q.to(cuda), k.to(cuda), v.to(cuda), 8, 80 ** -0.5)
    assert (got.float().cpu() - want).abs().max().item() < TOL


# ---------------------------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("n", [4 * 4 * 64 * 64, 1000003, 7])
def test_cfg_ddim_step_bit_exact_fp32(cuda, n):
    g = torch.Generator().manual_seed(n)
    eu, ec, x = (torch.randn(n, generator=g) for _ in range(3))
    a_t, a_p = torch.tensor(0.4217), torch.tensor(0.5531)
    want = orc.ddim_step(orc.cfg_combine(eu, ec, 7.5), x, a_t, a_p)
    got = ops.cfg_ddim_step(eu.to(cuda), ec.to(cuda), x.to(cuda), 7.5, float(a_t), float(a_p)).cpu()
    assert torch.equal(got, want), f"max abs diff {(got - want).abs().max().item()}"
    want2 = orc.ddim_step(eu, x, a_p, a_t)
    got2 = ops.cfg_ddim_step(eu.to(cuda), None, x.to(cuda), 0.0, float(a_p), float(a_t)).cpu()
    assert torch.equal(got2, want2)


def test_cfg_ddim_step_bf16(cuda):
    g = torch.Generator().manual_seed(0)
    eu, ec, x = (torch.randn(4096, generator=g).to(torch.bfloat16) for _ in range(3))
    want = orc.ddim_step(orc.cfg_combine(eu.float(), ec.float(), 7.5), x.float(), torch.tensor(0.3), torch.tensor(0.35))
    got = ops.cfg_ddim_step(eu.to(cuda), ec.to(cuda), x.to(cuda), 7.5, 0.3, 0.35).float().cpu()
    assert (got - want).abs().max().item() < 0.1 and torch.allclose(got, want, rtol=1e-2, atol=2e-2)


def test_store_accumulate_bit_exact(cuda):
    g = torch.Generator().manual_seed(1)
    sizes = [16 * 256 * 77, 16 * 1024 * 77, 5, 16 * 64 * 64, 3 * 1024 * 1024 + 1]
    dst = [torch.randn(s, generator=g) for s in sizes]
    src = [torch.randn(s, generator=g) for s in sizes]
    want = [d + s for d, s in zip(dst, src)]
    d_dev = [d.to(cuda) for d in dst]
    ops.store_accumulate(d_dev, [s.to(cuda) for s in src])
    for a, b in zip(d_dev, want):
        assert torch.equal(a.cpu(), b)


def test_local_blend(cuda):
    g = torch.Generator().manual_seed(2)
    heads = [8, 8, 8, 8, 8]
    maps = [torch.rand(2 * h, 256, 77, generator=g) ** 4 for h in heads]
    alpha = torch.zeros(2, 1, 1, 1, 1, 77)
    alpha[0, ..., 3] = 1
    alpha[1, ..., 3] = 1
    alpha[1, ..., 4] = 1
    x = torch.randn(2, 4, 64, 64, generator=g)
    want, want_mask = orc.local_blend(x, maps, alpha, 0.3, return_mask=True)
    got, mask = ops.local_blend(x.clone().to(cuda), [m.to(cuda) for m in maps], 2, alpha.reshape(2, 77).to(cuda), 0.3, return_mask=True)
    mism = (mask.cpu() != want_mask).float().mean().item()
    assert mism < 1e-3, f"mask mismatch fraction {mism}"
    if mism == 0:
        assert torch.equal(got.cpu(), want)


# ---------------------------------------------------------------------------------------------------- full-size properties
# At BASELINE.json's full geometries the CPU oracle would need tens of GB of probabilities; parity is checked there through
# size-independent properties of softmax attention (SD-2.1 @768^2: N=9216 d=64 H=5; SDXL @1024^2: N=4096 d=64 H=10; SD-1.5).
FULL = [("sd21_96", 4, 5, 9216, 64), ("sdxl_64", 4, 10, 4096, 64), ("sd15_64", 4, 8, 4096, 40)]


@pytest.mark.parametrize("name,B,H,N,d", FULL)
def test_full_size_properties(cuda, name, B, H, N, d):
    g = torch.Generator().manual_seed(N + d)
    q, k, v = (torch.randn(B, N, H * d, generator=g).to(torch.bfloat16).to(cuda) for _ in range(3))
    scale = d ** -0.5
    src = [0, 0, 2, 2]
    out = ops.attention(q, k, v, H, scale, k_src=src, v_src=src)
    # (1) rows of softmax sum to one: with V = 1 the output is exactly representable and must be 1
    ones = torch.ones_like(v)
    o1 = ops.attention(q, k, ones, H, scale, k_src=src, v_src=src).float()
    assert (o1 - 1).abs().max().item() < 1e-2
    # (2) permuting the keys (and values alike) leaves the output unchanged
    perm = torch.randperm(N, generator=g).to(cuda)
    o2 = ops.attention(q, k[:, perm].contiguous(), v[:, perm].contiguous(), H, scale, k_src=src, v_src=src)
    assert (o2.float() - out.float()).abs().max().item() < TOL
    # (3) linear in V
    v2 = torch.randn(B, N, H * d, generator=g).to(torch.bfloat16).to(cuda)
    o3 = ops.attention(q, k, (v.float() + v2.float()).to(torch.bfloat16), H, scale, k_src=src, v_src=src).float()
    o4 = ops.attention(q, k, v2, H, scale, k_src=src, v_src=src).float()
    assert (o3 - (out.float() + o4)).abs().max().item() < 2 * TOL
    # (4) the source-row indirection equals physically gathering the rows (MasaCtrl rows 1,3 read K,V of rows 0,2)
    o5 = ops.attention(q, k[src].contiguous(), v[src].contiguous(), H, scale)
    assert torch.equal(o5, out)
    # (5) both kernel families agree, and a strip of query rows matches the fp32 oracle
    o6 = ops.attention(q, k, v, H, scale, k_src=src, v_src=src, impl=ops.IEF_IMPL_MMA)
    assert (o6.float() - out.float()).abs().max().item() < TOL
    rows = slice(N - 130, N)
    want = orc.indexed_attention(q[:, rows].cpu(), k.cpu(), v.cpu(), H, scale, k_src=src, v_src=src)
    assert (out[:, rows].float().cpu() - want).abs().max().item() < TOL


def test_attn_workspace_contract(cuda):
    """ief_attn_workspace_bytes: zero for plain calls (the key-norm pre-pass is an A/B option since the unshifted loop replaced
    it: IEF_TC3_SKIPMAX=2), B*H*Nq floats for stored maps behind a tcgen05 launch; a call without workspace, with too small a
    workspace, or with one gives the same softmax."""
    import ctypes as C
    B, H, N, d = 1, 2, 3200, 64
    q, k, v = (t.to(cuda) for t in _qkv(B, N, N, H, d, 41))
    out = torch.empty_like(q)

    def params(qq, kk, vv, dd):
        p = _cabi.AttnParams()
        q4, k4, v4, o4 = (ops._as4(t, H) for t in (qq, kk, vv, out if dd == d else torch.empty_like(qq)))
        p.q, p.k, p.v, p.o = (ops._t4(t, H) for t in (q4, k4, v4, o4))
        p.dtype = _cabi.IEF_BF16 if qq.dtype == torch.bfloat16 else _cabi.IEF_F16
        p.B, p.H, p.Nq, p.Nk, p.d = qq.shape[0], H, qq.shape[1], kk.shape[1], dd
        p.scale, p.impl = dd ** -0.5, _cabi.IEF_IMPL_TCGEN05
        return p

    lib = _cabi.lib()
    p = params(q, k, v, d)
    need = lib.ief_attn_workspace_bytes(C.byref(p))
    forced = os.environ.get("IEF_TC3_SKIPMAX") == "2"
    assert need == (B * H * 25 * 4 if forced else 0)
    need = B * H * 25 * 4   # what the A/B pre-pass would take: offering it must never change the result
    q40 = q[..., :H * 40].contiguous()
    if not forced:
        assert lib.ief_attn_workspace_bytes(C.byref(params(q40, q40, q40, 40))) == 0
    if not forced:
        assert lib.ief_attn_workspace_bytes(C.byref(params(q[:, :1024], k[:, :1024], v[:, :1024], d))) == 0
    qh = q.to(torch.float16)
    assert lib.ief_attn_workspace_bytes(C.byref(params(qh, qh, qh, d))) == 0                # fp16 keeps the exact maximum
    stream = torch.cuda.current_stream().cuda_stream
    results = []
    for ws_bytes in (0, need // 2, need):
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=cuda)
        p.workspace, p.workspace_bytes = (ws.data_ptr(), ws_bytes) if ws_bytes else (None, 0)
        out.zero_()
        _cabi.check("ief_attn_fwd", lib.ief_attn_fwd(C.byref(p), stream))
        torch.cuda.synchronize()
        results.append(out.float().cpu().clone())
    want = orc.indexed_attention(q.cpu(), k.cpu(), v.cpu(), H, d ** -0.5)
    for r in results:
        assert (r - want).abs().max().item() < TOL
    assert torch.equal(results[0], results[1])          # too small a workspace is ignored: identical launches


@pytest.mark.parametrize("shape", [(2, 8, 1024, 77, 40), (2, 4, 300, 77, 80), (1, 2, 256, 77, 160), (2, 2, 4096, 64, 64)])
@pytest.mark.parametrize("with_dprobs", [True, False])
def test_cross_attention_backward(cuda, shape, with_dprobs):
    """ief_cross_attn_bwd through the autograd Function of the Pix2Pix-zero processor: dQ, dK, dV of
    loss = <O, W_o> + <P, W_p> against torch autograd on the fp32 materialised formulation."""
    from image_editing_framework_b200.pix2pix_zero.attention_control import _CrossAttention
    B, H, N, M, d = shape
    q, k, v = _qkv(B, N, M, H, d, 17)
    scale = d ** -0.5
    g = torch.Generator().manual_seed(18)
    wo = torch.randn(B, N, H * d, generator=g)
    wp = torch.randn(B * H, N, M, generator=g) if with_dprobs else None
    # reference: fp32 autograd on the bf16-rounded inputs
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    p_ref = orc.attention_probs(qf, kf, H, scale)
    o_ref = orc.apply_probs(p_ref, vf, H)
    loss = (o_ref * wo).sum() + ((p_ref * wp).sum() if with_dprobs else 0.0)
    loss.backward()
    # kernels
    qc, kc, vc = (t.to(cuda).clone().requires_grad_(True) for t in (q, k, v))
    out, probs = _CrossAttention.apply(qc, kc, vc, H, scale)
    loss_c = (out.float() * wo.to(cuda)).sum() + ((probs * wp.to(cuda)).sum() if with_dprobs else 0.0)
    loss_c.backward()
    torch.cuda.synchronize()
    assert (probs.cpu() - p_ref.detach()).abs().max().item() < 5e-3
    for name, got, want in (("dq", qc.grad, qf.grad), ("dk", kc.grad, kf.grad), ("dv", vc.grad, vf.grad)):
        err = (got.float().cpu() - want).abs().max().item()
        ref = want.abs().max().item()
        assert err <= 2e-2 * ref + 1e-3, f"{name}: max abs err {err} vs max |grad| {ref}"
    # dQ alone (K, V detached): no dS output requested
    qd = q.to(cuda).clone().requires_grad_(True)
    out2, _ = _CrossAttention.apply(qd, k.to(cuda), v.to(cuda), H, scale)
    (out2.float() * wo.to(cuda)).sum().backward()
    qf2 = q.float().clone().requires_grad_(True)
    (orc.apply_probs(orc.attention_probs(qf2, k.float(), H, scale), v.float(), H) * wo).sum().backward()
    assert (qd.grad.float().cpu() - qf2.grad).abs().max().item() <= 2e-2 * qf2.grad.abs().max().item() + 1e-3


@pytest.mark.parametrize("accum", [False, True])
def test_attn_probs_out_two_key_blocks_odd_key_count(cuda, accum):
    """Stored maps of a call with a second key/value block and an ODD key count (the second block then starts on an odd column
    of the map rows): found by tools/fuzz_attn.py as a misaligned 8-byte access in the log-sum-exp sweep."""
    B, H, N, M, d = 2, 2, 640, 333, 64
    q, k, v = _qkv(B, N, M, H, d, 23)
    scale = d ** -0.5
    kw = dict(k_src=[0, 0], v_src=[0, 0], k_src2=[0, 1], v_src2=[0, 1])
    qh, kh = orc.head_to_batch(q.float(), H), orc.head_to_batch(torch.cat([k[[0, 0]], k[[0, 1]]], 1).float(), H)
    want_p = (torch.bmm(qh, kh.transpose(1, 2)) * scale).softmax(-1)
    want_o = orc.indexed_attention(q, k, v, H, scale, **kw)
    probs = torch.full((B * H, N, 2 * M), 0.25 if accum else 7.0, device=cuda)
    got = ops.attention(q.to(cuda), k.to(cuda), v.to(cuda), H, scale, probs_out=probs, probs_accum=accum, **kw)
    torch.cuda.synchronize()
    if os.environ.get("IEF_PROBS_VIA_LSE", "1") != "0":      # the switch that sends stored maps back to the two-sweep mma kernel
        assert _cabi.last_attn_impl() == "tcgen05+probs"
    assert (got.float().cpu() - want_o).abs().max().item() < TOL
    assert (probs.cpu() - (0.25 if accum else 0.0) - want_p).abs().max().item() < 5e-3


# ---------------------------------------------------------------------------------------------------- the documented binding
def test_integration_md_stub_runs_the_kernel(cuda):
    """INTEGRATION.md §2's ctypes stub — the binding a reference maintainer would paste — drives ief_attn_fwd on its own and
    reproduces MasaCtrl's mutual self-attention (masactrl/model/attention_control.py:51-58: targets read the sources' K/V)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    stub = re.search(r"```python\n(# p2p/model/_ief\.py.*?)```", doc, re.S).group(1)
    ns = {}
    exec(compile(stub.replace('C.CDLL("libief_b200.so")', f'C.CDLL({_cabi.LIB_PATH!r})'), "INTEGRATION.md", "exec"), ns)
    B, N, H, d = 4, 1024, 8, 40
    q, k, v = _qkv(B, N, N, H, d, seed=77)
    src = [0, 0, 2, 2]
    before = _cabi.launch_count()
    got = ns["attn_fwd"](q.to(cuda), k.to(cuda), v.to(cuda), H, d ** -0.5, k_src=src, v_src=src)
    torch.cuda.synchronize()
    assert _cabi.launch_count() > before
    want = orc.indexed_attention(q, k, v, H, d ** -0.5, k_src=src, v_src=src)
    assert (got.float().cpu() - want).abs().max().item() < TOL


def test_two_devices_driven_from_two_threads_of_one_process():
    """The library keeps no per-process launch state: per-device (kernel, device) configuration, the caller's stream, a thread-local
    error string. cuda:0 first, then cuda:1 (the order that used to leave cuda:1 without its > 48 KB shared-memory opt-in), then both
    concurrently from two threads, for a tcgen05 self-attention, an edited cross-attention and the mma.sync path each."""
    import threading
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    B, H, N, d = 4, 8, 1024, 64
    q, k, v = _qkv(B, N, N, H, d, 5)
    kc, vc = (torch.randn(B, 77, H * d, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16) for _ in range(2))
    src = [0, 0, 2, 2]
    want_self = orc.indexed_attention(q, k, v, H, d ** -0.5, k_src=src, v_src=src)
    want_cross = orc.plain_attention(q, kc, vc, H, d ** -0.5)
    errors = []

    def work(index, rounds):
        try:
            dev = torch.device("cuda", index)
            with torch.cuda.device(dev):
                qd, kd, vd, kcd, vcd = (t.to(dev) for t in (q, k, v, kc, vc))
                for _ in range(rounds):
                    for impl in (ops.IEF_IMPL_TCGEN05, ops.IEF_IMPL_MMA):
                        got = ops.attention(qd, kd, vd, H, d ** -0.5, k_src=src, v_src=src, impl=impl)
                        assert (got.float().cpu() - want_self).abs().max().item() < TOL
                    got = ops.cross_attention_edit(qd, kcd, vcd, H, d ** -0.5)
                    assert (got.float().cpu() - want_cross).abs().max().item() < TOL
        except Exception as e:   # noqa: BLE001 - reported to the main thread
            errors.append((index, repr(e)))

    work(0, 1)
    work(1, 1)
    assert not errors, errors
    threads = [threading.Thread(target=work, args=(i, 5)) for i in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
