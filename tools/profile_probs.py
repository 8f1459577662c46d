"""Target for `ncu -k regex:probs_from_lse`: the stored-maps path (tcgen05 O + row log-sum-exp, then the map-writing sweep) at SD-1.5's 32x32 layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops
dev = torch.device("cuda:0")
B, H, N, d = 4, 8, 1024, 80
q, k, v = (torch.randn(B, N, H * d, device=dev).to(torch.bfloat16) for _ in range(3))
probs = torch.zeros(2 * H, N, N, device=dev)
for _ in range(4):
    ops.attention(q, k, v, H, d ** -0.5, probs_out=probs, probs_accum=True, probs_slot=[-1, -1, 0, 1])
torch.cuda.synchronize()
print("ok")
