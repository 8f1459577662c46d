"""Plain cross-attention (77 keys): device time per launch from 50 launches replayed from a CUDA graph (warm caches, no host gaps) and from single launches with the L2 flushed. IEF_CROSS_TC=0 selects the mma.sync kernel."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B, H = 4, 8
for N, d in ((4096, 40), (1024, 80), (256, 160), (4096, 64)):
    q = torch.randn(B, N, H * d, device=dev).to(torch.bfloat16)
    k, v = (torch.randn(B, 77, H * d, device=dev).to(torch.bfloat16) for _ in range(2))
    o = torch.empty_like(q)
    fn = lambda: ops.cross_attention_edit(q, k, v, H, d ** -0.5, out=o)
    for _ in range(5):
        fn()
    # 50 launches replayed from a CUDA graph: device time per launch with warm caches and no host gaps
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        for _ in range(50):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    e.synchronize()
    warm = s.elapsed_time(e) / 50
    ts = []
    for _ in range(20):
        flush.zero_()
        torch.cuda._sleep(400000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    nbytes = 2 * q.numel() * 2 + 2 * k.numel() * 2
    print(json.dumps(dict(N=N, d=d, kernel="tcgen05" if os.environ.get("IEF_CROSS_TC", "1") != "0" else "mma.sync", us_back_to_back=round(warm * 1e3, 2),
                          us_cold_l2=round(ts[len(ts) // 2] * 1e3, 2), GBps_cold=round(nbytes / ts[len(ts) // 2] / 1e6, 1))), flush=True)
