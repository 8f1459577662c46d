import ctypes as C, os, sys
sys.path.insert(0, '/root/repo')
import torch
from image_editing_framework_b200 import ops, _cabi
B, H, N, d = (int(x) for x in sys.argv[1:5])
dev = torch.device("cuda:0")
q, k, v = (torch.randn(B, N, H * d, device=dev).to(torch.bfloat16) for _ in range(3))
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
lib = _cabi.lib(); lib.ief_debug_set_trace_buffer.argtypes = [C.c_void_p]
for _ in range(3): ops.attention(q, k, v, H, d ** -0.5, impl=ops.IEF_IMPL_TCGEN05)
lib.ief_debug_set_trace_buffer(buf.data_ptr()); ops.attention(q, k, v, H, d ** -0.5, impl=ops.IEF_IMPL_TCGEN05); torch.cuda.synchronize(); lib.ief_debug_set_trace_buffer(None)
t = buf.cpu(); k0 = int(t[1536])
names = {1539: "last PV", 1542: "row sums", 1543: "decision", 1546: "w4 before ld", 1547: "w4 after ld", 1548: "w4 after stores (first chunk)", 1550: "w19 before ld", 1551: "w19 after ld", 1552: "w19 after stores", 1553: "w19 loop done", 1540: "w4 epilogue end", 1541: "exit"}
for i, n in sorted(names.items(), key=lambda x: int(t[x[0]])): print(f"{n:32s} {int(t[i]) - k0}")
