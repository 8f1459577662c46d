"""Tiny target for `ncu --set full`: the dominant kernel shape only (B=4,H=8,N=4096,d=40 MasaCtrl-style sources)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops

B, H, N, d = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (4, 8, 4096, 40)))
impl = {"tc": ops.IEF_IMPL_TCGEN05, "mma": ops.IEF_IMPL_MMA}[sys.argv[5] if len(sys.argv) > 5 else "tc"]
dev = torch.device("cuda:0")
q, k, v = (torch.randn(B, N, H * d, device=dev).to(torch.bfloat16) for _ in range(3))
src = [2 * (b // 2) for b in range(B)]
for _ in range(8):
    o = ops.attention(q, k, v, H, d ** -0.5, k_src=src, v_src=src, impl=impl)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
