"""Per-phase clock64 trace of CTA 0 of the generation-3 tcgen05 attention kernel (debug hook ief_debug_set_trace_buffer):
one softmax thread (row 0, column half 0) of each stream stamps every tile.
Needs a traced build:  IEF_EXTRA_NVCC_FLAGS="-DIEF_TC3_TRACE=2" bash image_editing_framework_b200/csrc/build.sh
(the default library has the stamps compiled out)."""
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
if int(t[:1024].abs().sum()) == 0:
    sys.exit("no stamps recorded: rebuild the library with -DIEF_TC3_TRACE=2 (see the docstring)")
sm = t[:1024].view(2, 64, 8)
t0 = int(sm[0, 0, 0])
nt = min(N // 128, 64)
s2 = t[1024:1536].view(2, 64, 4)
print("stream tile | start | wait_S  ld_S  [max  xchg-barrier  decide  scale]  wait_PV(j-1)  wait_turn  exp+st  st_wait | period")
for j in range(4, min(nt, 12)):
    for s_ in range(2):
        r = [int(x) for x in sm[s_, j]]
        period = r[0] - int(sm[s_, j - 1, 0])
        if int(s2[s_, j, 0]) == 0:   # unshifted loop (MAXMODE 2): no maximum / exchange / decision sub-phases
            print(f"  t{s_} j={j:2d} | {r[0] - t0:7d} | {r[1] - r[0]:6d} {r[2] - r[1]:5d} [   -      -      - {r[6] - r[2]:6d}] {r[7] - r[6]:8d} {r[3] - r[7]:10d} {r[4] - r[3]:8d} {r[5] - r[4]:7d} | {period}")
            continue
        print(f"  t{s_} j={j:2d} | {r[0] - t0:7d} | {r[1] - r[0]:6d} {r[2] - r[1]:5d} [{int(s2[s_, j, 0]) - r[2]:4d} {int(s2[s_, j, 1] - s2[s_, j, 0]):6d} {int(s2[s_, j, 2] - s2[s_, j, 1]):6d} {r[6] - int(s2[s_, j, 2]):6d}] {r[7] - r[6]:8d} {r[3] - r[7]:10d} {r[4] - r[3]:8d} {r[5] - r[4]:7d} | {period}")
tot = int(sm[0, nt - 1, 5] - sm[0, 0, 0])
print(f"total cycles for {nt} tiles (stream 0): {tot} -> {tot / nt:.0f} per tile")
c = t[1536:1542]
if int(c[0]) != 0:
    k0 = int(c[0])
    print(f"CTA: setup {int(c[1]) - k0}, first tile start {t0 - k0}, first S ready {int(sm[0, 0, 1]) - k0}, loop end {int(c[2]) - k0}, last PV {int(c[3]) - k0}, epilogue {int(c[4]) - k0}, exit {int(c[5]) - k0}")
it = t[1600:1728].view(16, 2, 4)
if int(it[0, 0, 0]) != 0:
    print("persistent CTA 0, per item and stream: start | loop end (last PV done) | epilogue end   [cycles since CTA start; item length = start-to-start]")
    for k_ in range(16):
        if int(it[k_, 0, 0]) == 0:
            break
        row_ = []
        for s_ in range(2):
            a0, a1, a2 = (int(x) - int(c[0]) for x in it[k_, s_, :3])
            top = int(it[k_, s_, 3]) - int(c[0])   # top of the item loop (before the item decode)
            nxt = int(it[k_ + 1, s_, 0]) - int(c[0]) if k_ + 1 < 16 and int(it[k_ + 1, s_, 0]) != 0 else None
            row_.append(f"t{s_}: {a0:7d} {a1:7d} {a2:7d} (decode {a0 - top}, loop {a1 - a0}, epilogue {a2 - a1}" + (f", to next start {nxt - a2}, item {nxt - a0})" if nxt is not None else ")"))
        print(f"  item {k_:2d} | " + " | ".join(row_))
