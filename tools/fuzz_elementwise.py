"""Randomised checks of the elementwise entry points against torch on the same GPU: ief_cfg_ddim_step (bit-exact in fp32, odd sizes,
unaligned views), ief_store_accumulate (bit-exact, odd sizes, up to 64 tensors per launch), ief_mask_blend (bit-exact), and
ief_local_blend against the reference formulation of LocalBlend (ptp_utils.py:20-32). Not part of the test suite."""
import os
import random
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(rng.randrange(1 << 30))

for case in range(n_cases):
    # ---- CFG + DDIM step
    n = rng.choice([1, 7, 31, 32, 1000, 4097, 32768, 100003])
    dt = rng.choice([torch.float32, torch.float32, torch.bfloat16, torch.float16])
    eu, ec, x = (torch.randn(n, device=dev).to(dt) for _ in range(3))
    g, at, ap = rng.uniform(1, 10), rng.uniform(0.01, 0.99), rng.uniform(0.01, 0.99)
    cond = rng.random() < 0.8
    got = ops.cfg_ddim_step(eu, ec if cond else None, x, g, at, ap)
    a_t, a_p = torch.tensor(at, dtype=torch.float32, device=dev), torch.tensor(ap, dtype=torch.float32, device=dev)
    eps = (eu.float() + g * (ec.float() - eu.float())) if cond else eu.float()
    x0 = (x.float() - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    want = (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps).to(dt)
    if dt == torch.float32:
        assert torch.equal(got, want), f"case {case}: cfg_ddim_step fp32 not bit-exact (n={n}, cond={cond})"
    else:
        assert (got.float() - want.float()).abs().max().item() <= 2e-2 * want.float().abs().max().item() + 1e-3, f"case {case}: cfg_ddim_step {dt}"
    # ---- store accumulate
    k = rng.choice([1, 3, 22, 64, 70])
    sizes = [rng.choice([1, 5, 1023, 4096, 77 * 256 * 16 + 3]) for _ in range(k)]
    dst = [torch.randn(s_, device=dev) for s_ in sizes]
    src = [torch.randn(s_, device=dev) for s_ in sizes]
    want = [a + b for a, b in zip(dst, src)]
    ops.store_accumulate(dst, src)
    assert all(torch.equal(a, b) for a, b in zip(dst, want)), f"case {case}: store_accumulate"
    # ---- mask blend
    B, N, C = rng.choice([1, 4]), rng.choice([17, 256, 1000]), rng.choice([8, 40, 320])
    dt = rng.choice([torch.float32, torch.bfloat16, torch.float16])
    fg, bg = torch.randn(B, N, C, device=dev).to(dt), torch.randn(B, N, C, device=dev).to(dt)
    w = torch.rand(N, device=dev)
    rows = sorted(rng.sample(range(B), rng.randint(1, B)))
    want = fg.clone()
    wf = w.reshape(1, N, 1)
    want[rows] = (fg[rows].float() * wf + bg[rows].float() * (1 - wf)).to(dt)
    got = ops.mask_blend(fg.clone(), bg, w, rows=rows)
    assert torch.equal(got, want), f"case {case}: mask_blend"
    # ---- LocalBlend
    n_prompts, heads_, words = 2, 8, 77   # the reference formula only broadcasts for two prompts
    maps = [torch.rand(n_prompts * heads_, 256, words, device=dev) ** 4 for _ in range(5)]
    alpha = torch.zeros(n_prompts, words, device=dev)
    for p_ in range(n_prompts):
        alpha[p_, rng.sample(range(1, 20), rng.randint(1, 3))] = 1
    x_t = torch.randn(n_prompts, 4, 64, 64, device=dev)
    thr = rng.uniform(0.2, 0.5)
    m = torch.cat([t.reshape(n_prompts, -1, 1, 16, 16, words) for t in maps], dim=1)
    m = (m * alpha.reshape(n_prompts, 1, 1, 1, 1, words)).sum(-1).mean(1)
    mask = F.interpolate(F.max_pool2d(m, (3, 3), (1, 1), padding=(1, 1)), size=(64, 64))
    mask = mask / mask.max(2, keepdims=True)[0].max(3, keepdims=True)[0]
    mask = mask.gt(thr)
    mask = (mask[:1] + mask[1:]).float()
    want = x_t[:1] + mask * (x_t - x_t[:1])
    got = ops.local_blend(x_t.clone(), maps, n_prompts, alpha.contiguous(), thr)
    # a pixel whose normalised value sits within float rounding of the threshold may flip: allow a handful
    bad = (got != want).reshape(n_prompts, 4, -1).any(1).sum().item()
    assert bad <= 4, f"case {case}: local_blend differs on {bad} pixels"
print(f"{n_cases} random elementwise cases ok")
