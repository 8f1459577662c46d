"""Why was PnP on the SD-2.1 / SDXL stand-ins 60 % slower and 8 GB bigger than plain sampling (profiles/r01_methods_sd21.jsonl)?
PnP hooks 8 self-attention layers; the other 24 run the stand-in's own AttnProcessor (F.scaled_dot_product_attention when use_sdpa).
Times that call per layer shape against this library's kernels, with the peak memory each allocates."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / n, (torch.cuda.max_memory_allocated() - base) / 2 ** 20


for name, B, H, N, d, M in (("sd21 96^2 self", 4, 5, 9216, 64, 9216), ("sd21 96^2 cross", 4, 5, 9216, 64, 77), ("sd21 48^2 self", 4, 10, 2304, 64, 2304),
                            ("sd21 48^2 cross", 4, 10, 2304, 64, 77), ("sd15 64^2 self", 4, 8, 4096, 40, 4096), ("sd15 64^2 cross", 4, 8, 4096, 40, 77),
                            ("sdxl 64^2 self", 4, 10, 4096, 64, 4096), ("sdxl 64^2 cross", 4, 10, 4096, 64, 77)):
    q = torch.randn(B, N, H * d, device=dev, dtype=torch.bfloat16)
    k, v = (torch.randn(B, M, H * d, device=dev, dtype=torch.bfloat16) for _ in range(2))

    def sdpa():
        q4, k4, v4 = (t.view(t.shape[0], t.shape[1], H, d).transpose(1, 2) for t in (q, k, v))
        return F.scaled_dot_product_attention(q4, k4, v4, scale=d ** -0.5).transpose(1, 2).reshape(B, N, -1)

    def ours():
        return ops.cross_attention_edit(q, k, v, H, d ** -0.5) if M <= 80 else ops.attention(q, k, v, H, d ** -0.5)

    ms_s, mb_s = timed(sdpa)
    ms_o, mb_o = timed(ours)
    print(json.dumps(dict(layer=name, sdpa_ms=round(ms_s, 4), sdpa_peak_MiB=round(mb_s, 1), ours_ms=round(ms_o, 4), ours_peak_MiB=round(mb_o, 1))), flush=True)
