"""Diagnostic for test_attn_tcgen05_unshifted_softmax_second_pass[overflow-shape3]: where is the error and which path makes it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from image_editing_framework_b200 import ops, _cabi
from oracle import controlled_attention as orc

B, H, N, M, d = 2, 10, 2304, 2304, 64
g = torch.Generator().manual_seed(57 + N)
q, k, v = (torch.randn(B, n, H * d, generator=g).to(torch.bfloat16) for n in (N, M, M))
mult = float(sys.argv[1]) if len(sys.argv) > 1 else 7.0
q, k = (q.float() * mult).to(torch.bfloat16), (k.float() * mult).to(torch.bfloat16)
scale = d ** -0.5
want = orc.indexed_attention(q, k, v, H, scale)
dev = torch.device("cuda:0")
for dt in (torch.bfloat16, torch.float16):
    for impl in (ops.IEF_IMPL_TCGEN05, ops.IEF_IMPL_MMA):
        got = ops.attention(q.to(dev).to(dt), k.to(dev).to(dt), v.to(dev).to(dt), H, scale, impl=impl)
        torch.cuda.synchronize()
        e = (got.float().cpu() - want).abs()
        idx = e.argmax().item()
        b, n, c = idx // (N * H * d), (idx // (H * d)) % N, idx % (H * d)
        h = c // d
        s = (q[b, n, h * d:(h + 1) * d].float() @ k[b, :, h * d:(h + 1) * d].float().T) * scale
        top = s.topk(3)
        print(f"{dt} impl={_cabi.last_attn_impl()} NOMAX={os.environ.get('IEF_TC3_NOMAX')} max err {e.max().item():.4f} at b={b} n={n} h={h} ch={c % d} "
              f"top logits {top.values.tolist()} keys {top.indices.tolist()} v at top {v[b, top.indices, c].tolist()} want {want[b, n, c]:.4f} got {got[b, n, c].item():.4f}; rows with err>0.02: {(e.amax(-1) > 0.02).sum().item()}")
