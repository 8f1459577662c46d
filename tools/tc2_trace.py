"""Per-phase clock64 trace of CTA (0,0,0) of the v2 tcgen05 attention kernel (debug hook ief_debug_set_trace_buffer).
Needs a traced build: IEF_EXTRA_NVCC_FLAGS="-DIEF_TC2_TRACE=1" bash image_editing_framework_b200/csrc/build.sh, and
IEF_TC_VERSION=2 to select the generation-2 kernel for head_dim <= 64."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops, _cabi

B, H, N, d = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (4, 8, 4096, 40)))
dev = torch.device("cuda:0")
q, k, v = (torch.randn(B, N, H * d, device=dev).to(torch.bfloat16) for _ in range(3))
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
lib = _cabi.lib()
lib.ief_debug_set_trace_buffer.argtypes = [C.c_void_p]
for _ in range(3):
    ops.attention(q, k, v, H, d ** -0.5, impl=ops.IEF_IMPL_TCGEN05)
lib.ief_debug_set_trace_buffer(buf.data_ptr())
ops.attention(q, k, v, H, d ** -0.5, impl=ops.IEF_IMPL_TCGEN05)
torch.cuda.synchronize()
lib.ief_debug_set_trace_buffer(None)
t = buf.cpu()
sm = t[:1024].view(2, 64, 8)
mm = t[1024:1536].view(2, 64, 4)
t0 = int(sm[0, 0, 0])
nt = min(N // 128, 64)
print("softmax WG phases (cycles): wait_S | ld | max | exp+st-issue | wait_st+arrive   || tile period")
for wg in range(2):
    for j in range(2, min(nt, 14)):
        r = sm[wg, j]
        period = int(sm[wg, j, 0] - sm[wg, j - 1, 0])
        print(f"wg{wg} j={j:2d} start={int(r[0]) - t0:7d}  wait_S={int(r[1] - r[0]):5d} ld={int(r[2] - r[1]):4d} max={int(r[3] - r[2]):4d} exp={int(r[4] - r[3]):5d} st+arr={int(r[5] - r[4]):4d} | period={period}")
print("offset wg1.start - wg0.start per tile:", [int(sm[1, j, 0] - sm[0, j, 0]) for j in range(0, nt)])
print("exp start wg0/wg1 rel:", [(int(sm[0, j, 3]) - t0, int(sm[1, j, 3]) - t0) for j in range(0, min(nt, 8))])
print("MMA warp: wait_P(+K) | issue PV+QK+commits")
for tt in range(2):
    for j in range(2, min(nt - 1, 14)):
        r = mm[tt, j]
        print(f"t{tt} j={j:2d} at={int(r[0]) - t0:7d} wait_P={int(r[1] - r[0]):5d} issue={int(r[3] - r[1]):4d}")
c = t[1536:1542]
if int(c[0]) != 0:
    k0 = int(c[0])
    print(f"CTA phases (cycles from kernel entry): setup done={int(c[1]) - k0}  first S wait start={t0 - k0}  first S ready={int(sm[0, 0, 1]) - k0}  "
          f"loop end={int(c[2]) - k0}  last PV done={int(c[3]) - k0}  epilogue done={int(c[4]) - k0}  CTA exit={int(c[5]) - k0}")
    print("first steps wg0:", [(int(sm[0, j, 0]) - k0, int(sm[0, j, 5]) - k0) for j in range(0, 3)])
tot = int(sm[0, nt - 1, 5] - sm[0, 0, 0])
print(f"total cycles for {nt} tiles (wg0): {tot}  -> {tot / nt:.0f} per tile-pair")
