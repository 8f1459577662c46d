"""Launch the cross-attention kernel flavours a few times at SD-1.5 shapes (target for `ncu -k regex:cross`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")
B, H = 4, 8
N, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 40)
mapper = torch.eye(77, device=dev)[None].contiguous()
alpha = torch.ones(1, 77, device=dev)
edit = ops.CrossEdit(ops.IEF_EDIT_REPLACE, 1, mapper=mapper)
q = torch.randn(B, N, H * d, device=dev).to(torch.bfloat16)
k, v = (torch.randn(B, 77, H * d, device=dev).to(torch.bfloat16) for _ in range(2))
o = torch.empty_like(q)
for _ in range(3):
    ops.cross_attention_edit(q, k, v, H, d ** -0.5, out=o)
    ops.cross_attention_edit(q, k, v, H, d ** -0.5, edit=edit, step_alpha=alpha, base_row=[-1, -1, -1, 2], edit_slot=[0] * 4, out=o)
torch.cuda.synchronize()
print("ok")
