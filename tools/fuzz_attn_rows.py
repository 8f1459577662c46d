"""Second randomised sweep of ief_attn_fwd: the arguments fuzz_attn.py leaves alone — row masks (`rows`, untouched outputs),
`probs_slot` (rows whose maps are dropped or redirected), `probs_accum`, strided inputs (head-major '(b h) n d' views, the slices of
a fused QKV projection), and very small token counts. Reference: fp32 materialised attention with torch on the same GPU.
Usage: python tools/fuzz_attn_rows.py SEED N_CASES (run it under the switches of DESIGN.md section 3.1b too)."""
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
SENTINEL = 7.0


def heads_of(t, H):
    B, N, C = t.shape
    return t.float().reshape(B, N, H, C // H).permute(0, 2, 1, 3)


def make(layout, B, N, H, d, dtype):
    """A [B, N, H*d]-shaped (or 4-D) operand whose storage follows `layout`; returns (what ops.attention gets, dense [B,N,H*d])."""
    x = torch.randn(B, N, H * d, device=dev).to(dtype)
    if layout == "dense":
        return x, x
    if layout == "fused":          # a slice of a wider projection: token stride 3*H*d
        wide = torch.randn(B, N, 3 * H * d, device=dev).to(dtype)
        j = rng.randrange(3)
        wide[:, :, j * H * d:(j + 1) * H * d] = x
        return wide[:, :, j * H * d:(j + 1) * H * d], x
    if layout == "head_major":     # the reference's '(b h) n d' storage, handed over as a permuted 4-D view
        hm = x.reshape(B, N, H, d).permute(0, 2, 1, 3).contiguous()      # [B, H, N, d]
        return hm.permute(0, 2, 1, 3), x
    raise ValueError(layout)


worst = 0.0
for case in range(n_cases):
    dtype = rng.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    d = rng.choice([8, 16, 32, 40, 64, 80, 128, 160])
    H = rng.choice([1, 2, 5, 8])
    B = rng.choice([1, 2, 3, 4, 6])
    N = rng.choice([1, 3, 16, 17, 64, 64, 100, 256, 256, 640, 1024, 1024, 1536])
    M = rng.choice([N, N, N, 1, 5, 64, 129, N + 1])
    scale = d ** -0.5
    layouts = [rng.choice(["dense", "dense", "fused", "head_major"]) for _ in range(3)]
    (q_in, q), (k_in, k), (v_in, v) = [make(l, B, n, H, d, dtype) for l, n in zip(layouts, (N, M, M))]
    ident = list(range(B))
    src = lambda: [rng.randrange(B) for _ in range(B)] if rng.random() < 0.5 else ident
    q_src, k_src, v_src = src(), src(), src()
    rows = sorted(rng.sample(ident, rng.randint(1, B))) if rng.random() < 0.5 else None
    active = rows if rows is not None else ident
    want_probs = rng.random() < 0.5
    slot, n_slots, accum = None, B, False
    if want_probs:
        accum = rng.random() < 0.4
        if rng.random() < 0.6:      # some rows dropped, the others packed into a smaller store
            kept = [b for b in ident if rng.random() < 0.6]
            slot = [kept.index(b) if b in kept else -1 for b in ident]
            n_slots = max(1, len(kept))
    qq, kk, vv = heads_of(q, H)[q_src], heads_of(k, H)[k_src], heads_of(v, H)[v_src]
    p_ref = (torch.einsum("bhnd,bhmd->bhnm", qq, kk) * scale).softmax(-1)
    o_ref = torch.einsum("bhnm,bhmd->bhnd", p_ref, vv).permute(0, 2, 1, 3).reshape(B, N, H * d)
    impls = [ops.IEF_IMPL_AUTO, ops.IEF_IMPL_MMA] + ([ops.IEF_IMPL_TCGEN05] if not want_probs else [])
    for impl in impls:
        tag = (f"case {case}: dtype={dtype} B={B} H={H} N={N} M={M} d={d} layouts={layouts} src={q_src},{k_src},{v_src} rows={rows} "
               f"probs={want_probs} slot={slot} accum={accum} impl={impl}")
        print(tag, file=sys.stderr, flush=True)
        out = torch.full((B, N, H * d), SENTINEL, device=dev, dtype=dtype)
        probs = prior = None
        if want_probs:
            prior = torch.rand(n_slots, H, N, M, device=dev) if accum else torch.full((n_slots, H, N, M), SENTINEL, device=dev)
            probs = prior.clone()
        try:
            got = ops.attention(q_in, k_in, v_in, H, scale, q_src=q_src, k_src=k_src, v_src=v_src, impl=impl, rows=rows, out=out,
                                probs_out=probs, probs_accum=accum, probs_slot=slot)
        except _cabi.IefError as e:
            if impl == ops.IEF_IMPL_TCGEN05 and e.code == -2:
                continue            # a forced tcgen05 launch may refuse a shape; AUTO must not
            print("FAIL", tag, "raised", e)
            sys.exit(1)
        torch.cuda.synchronize()
        tag += f" -> {_cabi.last_attn_impl()}"
        for b in ident:
            if b in active:
                err = (got[b].float() - o_ref[b]).abs().max().item()
                worst = max(worst, err)
                ok = err < TOL and bool(torch.isfinite(got[b]).all())
            else:
                err, ok = 0.0, bool((got[b] == SENTINEL).all())
            if not ok:
                print("FAIL", tag, f"row {b} ({'active' if b in active else 'masked'}) err {err}")
                sys.exit(1)
        if want_probs:
            expect = prior.clone()
            for b in active:
                s = b if slot is None else slot[b]
                if s >= 0:
                    expect[s] = p_ref[b] + (prior[s] if accum else 0.0)
            perr = (probs - expect).abs().max().item()
            if not perr < 1e-2:
                print("FAIL", tag, "probs err", perr)
                sys.exit(1)
print(f"{n_cases} random cases ok, worst max-abs error {worst:.4f}")
