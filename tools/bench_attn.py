"""Kernel-only timing of ief_attn_fwd at the reference's attention geometries (CUDA events, L2 flushed between reps)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def time_call(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    shapes = [  # name, B, H, N, d
        ("sd15_64", 4, 8, 4096, 40), ("sd15_32", 4, 8, 1024, 80), ("sd15_16", 4, 8, 256, 160), ("sd15_8", 4, 8, 64, 160),
        ("sd21_96", 4, 5, 9216, 64), ("sd21_48", 4, 10, 2304, 64), ("sdxl_64", 4, 10, 4096, 64), ("sdxl_32", 4, 20, 1024, 64),
        ("big_d64", 8, 16, 4096, 64), ("big_d40", 16, 8, 4096, 40),
    ]
    impls = [("tcgen05", ops.IEF_IMPL_TCGEN05), ("mma", ops.IEF_IMPL_MMA)]
    args = sys.argv[1:]
    if any(a in ("tcgen05", "mma") for a in args):
        impls = [i for i in impls if i[0] in args]
    if "big" in args:   # the six large shapes only
        shapes = [s_ for s_ in shapes if s_[0] in ("sd15_64", "sd21_96", "sd21_48", "sdxl_64", "sdxl_32", "big_d64", "big_d40")]
    no_sdpa = "nosdpa" in args
    out = []
    for name, B, H, N, d in shapes:
        q, k, v = (torch.randn(B, N, H * d, device=dev).to(torch.bfloat16) for _ in range(3))
        o = torch.empty_like(q)
        flops = 4.0 * B * H * N * N * d
        KS = [2 * (b // 2) for b in range(B)]  # MasaCtrl-style: every odd row reads the K,V of the even row before it
        for iname, impl in impls:
            med, best = time_call(lambda: ops.attention(q, k, v, H, d ** -0.5, impl=impl, k_src=KS,
                                                       v_src=KS, out=o))
            rec = dict(shape=name, B=B, H=H, N=N, d=d, impl=iname, ms_median=round(med, 4), ms_best=round(best, 4),
                       tflops_median=round(flops / med / 1e9, 1), tflops_best=round(flops / best / 1e9, 1))
            print(json.dumps(rec), flush=True)
            out.append(rec)
        # torch SDPA (library flash attention) for context only
        q4, k4, v4 = (t.view(B, N, H, d).transpose(1, 2) for t in (q, k, v))
        if no_sdpa:
            continue
        try:
            med, best = time_call(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4))
            print(json.dumps(dict(shape=name, impl="torch_sdpa", ms_median=round(med, 4), tflops_median=round(flops / med / 1e9, 1))), flush=True)
        except Exception as ex:  # noqa
            print("sdpa failed", ex)
    return out


if __name__ == "__main__":
    main()
