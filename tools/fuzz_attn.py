"""Randomised parity sweep of ief_attn_fwd / ief_cross_attn_edit_fwd against an fp32 materialised reference computed with torch on the
same GPU (shapes, ragged sizes, row sources, second key block, key bias, probability output, both dtypes, all impls).
Not part of the test suite (takes a minute); run it under the kernel-selection switches of DESIGN.md section 3.1b."""
import os
import random
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops, _cabi

dev = torch.device("cuda:0")
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 80
TOL = 2e-2


def ref_attention(q, k, v, H, scale, q_src, k_src, v_src, k_src2, v_src2, bias):
    B, N, C = q.shape
    d = C // H
    def heads(t):
        return t.float().reshape(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3)
    qq, kk, vv = heads(q)[q_src], heads(k)[k_src], heads(v)[v_src]
    if k_src2 is not None:
        kk, vv = torch.cat([kk, heads(k)[k_src2]], 2), torch.cat([vv, heads(v)[v_src2]], 2)
    s = torch.einsum("bhnd,bhmd->bhnm", qq, kk) * scale
    if bias is not None:
        s = s + bias[:, None, None, :]
    p = s.softmax(-1)
    return torch.einsum("bhnm,bhmd->bhnd", p, vv).permute(0, 2, 1, 3).reshape(B, N, C), p


worst = 0.0
for case in range(n_cases):
    dtype = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    d = rng.choice([8, 16, 32, 40, 48, 56, 64, 80, 96, 128, 160])
    H = rng.choice([1, 2, 5, 8])
    B = rng.choice([1, 2, 4])
    N = rng.choice([64, 100, 128, 256, 300, 512, 640, 1000, 1024, 1536, 2048, 3200])
    cross = rng.random() < 0.2
    M = 77 if cross else rng.choice([N, N, max(64, N // 2), N + 37])
    if N * M * B * H > 2 ** 26:
        N = 512
        M = 512 if not cross else 77
    q = torch.randn(B, N, H * d, device=dev).to(dtype)
    k = torch.randn(B, M, H * d, device=dev).to(dtype)
    v = torch.randn(B, M, H * d, device=dev).to(dtype)
    scale = d ** -0.5 * rng.choice([1.0, 1.0, 3.0])
    ident = list(range(B))
    src = lambda: [rng.randrange(B) for _ in range(B)] if rng.random() < 0.5 else ident
    q_src, k_src, v_src = src(), src(), src()
    use2 = (not cross) and rng.random() < 0.2
    k2 = [rng.randrange(B) for _ in range(B)] if use2 else None
    use_bias = (not cross) and (not use2) and rng.random() < 0.25
    bias_t = None
    kw = {}
    if use_bias:
        bt = torch.randn(2, M, device=dev)
        bt[0, ::3] = torch.finfo(torch.float32).min
        sel = [rng.choice([-1, 0, 1]) for _ in range(B)]
        kw.update(key_bias=bt.contiguous(), bias_sel=sel)
        bias_t = torch.stack([bt[s_] if s_ >= 0 else torch.zeros(M, device=dev) for s_ in sel])
    want_probs = (not use_bias) and rng.random() < 0.25
    impls = [ops.IEF_IMPL_AUTO, ops.IEF_IMPL_MMA] + ([ops.IEF_IMPL_TCGEN05] if d % 8 == 0 and not want_probs and (not use_bias or d <= 64) else [])
    want, p_ref = ref_attention(q, k, v, H, scale, q_src, k_src, v_src, k2, k2, bias_t)
    for impl in impls:
        print(f"case {case} impl {impl}: dtype={dtype} B={B} H={H} N={N} M={M} d={d} src={q_src},{k_src},{v_src} k2={k2} bias={use_bias} probs={want_probs}", file=sys.stderr, flush=True)
        probs = None
        if want_probs and impl != ops.IEF_IMPL_TCGEN05:
            probs = torch.zeros(B * H, N, M * (2 if use2 else 1), device=dev)
        if cross and impl == ops.IEF_IMPL_AUTO and not use_bias and q_src == ident and k_src == ident and v_src == ident:
            got = ops.cross_attention_edit(q, k, v, H, scale, probs_out=probs)
        else:
            got = ops.attention(q, k, v, H, scale, q_src=q_src, k_src=k_src, v_src=v_src, k_src2=k2, v_src2=k2, impl=impl, probs_out=probs, **kw)
        torch.cuda.synchronize()
        err = (got.float() - want).abs().max().item()
        perr = (probs - p_ref.reshape(B * H, N, -1)).abs().max().item() if probs is not None else 0.0
        worst = max(worst, err)
        tag = f"case {case}: dtype={dtype} B={B} H={H} N={N} M={M} d={d} scale={scale:.3f} src={q_src},{k_src},{v_src} k2={k2} bias={use_bias} probs={want_probs} impl={impl} -> {_cabi.last_attn_impl()}"
        tol = TOL * max(1.0, scale * d ** 0.5)   # the 2e-2 gate is stated for scale = d^-0.5; tripled logits triple the bf16 error
        if not (err < tol and perr < 1e-2) or not torch.isfinite(got).all():
            print("FAIL", tag, "err", err, "probs err", perr)
            sys.exit(1)
print(f"{n_cases} random cases ok, worst max-abs error {worst:.4f}")
