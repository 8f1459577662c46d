"""CUDA-event timings of the HBM / latency-bound kernels at SD-1.5 shapes, with algorithmic bytes and GB/s against the measured
copy bandwidth (MEASURED_PEAKS.json hbm_gbs). L2 is flushed between repetitions and the call is enqueued behind a
spin kernel, so the events bracket device time only (not the Python/ctypes call overhead)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def time_call(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(400000)  # keep the GPU busy while the host builds the call, so the events bracket the kernel only
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes, extra=""):
    gbs = nbytes / ms / 1e6
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), algorithmic_MB=round(nbytes / 1e6, 2), GBps=round(gbs, 1), frac_of_measured_hbm=round(gbs / PEAK, 3), note=extra)), flush=True)


def main():
    B, H = 4, 8
    mapper = torch.eye(77, device=dev)[None].contiguous()
    alpha = torch.ones(1, 77, device=dev)
    edit = ops.CrossEdit(ops.IEF_EDIT_REPLACE, 1, mapper=mapper)
    for N, d in ((4096, 40), (1024, 80), (256, 160)):
        q = torch.randn(B, N, H * d, device=dev).to(torch.bfloat16)
        k, v = (torch.randn(B, 77, H * d, device=dev).to(torch.bfloat16) for _ in range(2))
        o = torch.empty_like(q)
        qo = 2 * q.numel() * 2 + 2 * k.numel() * 2
        ms = time_call(lambda: ops.cross_attention_edit(q, k, v, H, d ** -0.5, out=o))
        report(f"cross_attn plain N={N} d={d}", ms, qo)
        ms = time_call(lambda: ops.cross_attention_edit(q, k, v, H, d ** -0.5, edit=edit, step_alpha=alpha, base_row=[-1, -1, -1, 2], edit_slot=[0] * 4, out=o))
        report(f"cross_attn P2P replace edit N={N} d={d}", ms, qo + q.numel() // B * 2, "row 3 also reads row 2's Q tile")
        if N <= 1024:
            store = torch.zeros(2 * H, N, 77, device=dev)
            ms = time_call(lambda: ops.cross_attention_edit(q, k, v, H, d ** -0.5, edit=edit, step_alpha=alpha, base_row=[-1, -1, -1, 2], edit_slot=[0] * 4,
                                                            probs_out=store, probs_accum=True, store_slot=[-1, -1, 0, 1], out=o))
            report(f"cross_attn edit + store accumulate N={N} d={d}", ms, qo + q.numel() // B * 2 + 2 * store.numel() * 4, "store: read + write fp32 maps of the cond half")
    # self-attention with probability output (AttentionStore self maps, N <= 1024)
    for N, d in ((1024, 80), (256, 160)):
        q, k, v = (torch.randn(B, N, H * d, device=dev).to(torch.bfloat16) for _ in range(3))
        probs = torch.zeros(2 * H, N, N, device=dev)
        ms = time_call(lambda: ops.attention(q, k, v, H, d ** -0.5, probs_out=probs, probs_slot=[-1, -1, 0, 1]))
        report(f"self-attn + probs store N={N} d={d}", ms, 4 * q.numel() * 2 + probs.numel() * 4, "fp32 maps of the cond half written once")
        ms = time_call(lambda: ops.attention(q, k, v, H, d ** -0.5, probs_out=probs, probs_accum=True, probs_slot=[-1, -1, 0, 1]))
        report(f"self-attn + probs accumulate N={N} d={d}", ms, 4 * q.numel() * 2 + 2 * probs.numel() * 4)
    # AttentionStore.between_steps as the reference does it (separate += pass) — kept for masactrl's store
    sizes = [16 * 1024 * 1024] * 5 + [16 * 256 * 256] * 5 + [16 * 64 * 64] + [16 * 1024 * 77] * 5 + [16 * 256 * 77] * 5 + [16 * 64 * 77]
    dst = [torch.zeros(s, device=dev) for s in sizes]
    src = [torch.ones(s, device=dev) for s in sizes]
    ms = time_call(lambda: ops.store_accumulate(dst, src))
    report("store_accumulate (SD-1.5 step store, 22 tensors, 1 launch)", ms, 3 * 4 * sum(sizes))
    ms = time_call(lambda: [d.add_(s) for d, s in zip(dst, src)])
    report("  same with torch `+=` loop (reference's between_steps)", ms, 3 * 4 * sum(sizes))
    # CFG + DDIM step
    for dt in (torch.float32, torch.bfloat16):
        eu, ec, x = (torch.randn(2 * 4 * 64 * 64, device=dev).to(dt) for _ in range(3))
        out = torch.empty_like(x)
        ms = time_call(lambda: ops.cfg_ddim_step(eu, ec, x, 7.5, 0.5, 0.6, out=out))
        report(f"cfg_ddim_step {dt}".replace("torch.", ""), ms, 4 * x.numel() * x.element_size(), "launch-latency bound: 32K elements")
    a_t, a_p = torch.tensor(0.5, device=dev), torch.tensor(0.6, device=dev)
    eu, ec, x = (torch.randn(2 * 4 * 64 * 64, device=dev) for _ in range(3))

    def torch_step():
        eps = eu + 7.5 * (ec - eu)
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps
    ms = time_call(torch_step)
    report("  same with torch eager ops (reference)", ms, 4 * x.numel() * 4)
    # LocalBlend
    maps = [torch.rand(16, 256, 77, device=dev) for _ in range(5)]
    wa = torch.zeros(2, 77, device=dev)
    wa[:, 3] = 1
    xt = torch.randn(2, 4, 64, 64, device=dev)
    ms = time_call(lambda: ops.local_blend(xt, maps, 2, wa, 0.3))
    report("local_blend (5 maps of 16x16, 2 launches)", ms, sum(m.numel() for m in maps) * 4 // 77 * 32 + 2 * xt.numel() * 4, "only the selected words' sectors are read")


if __name__ == "__main__":
    main()
