"""Counts CUDA kernel launches of one UNet forward (B=1 and B=4) of the bench workload with torch.profiler."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import io, contextlib
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from image_editing_framework_b200 import masactrl
from image_editing_framework_b200.editing import encode_prompts

dev = torch.device("cuda:0")
pipe, cfg = bench.build_pipeline("sd15", dev, torch.bfloat16)
if "--nchw" not in sys.argv:
    pipe.unet.to(memory_format=torch.channels_last)  # as bench.py runs it
ctx = encode_prompts(pipe, bench.PROMPTS)
with contextlib.redirect_stdout(io.StringIO()):
    ed = masactrl.MutualSelfAttentionControl(4, 10, total_steps=50)
masactrl.regiter_attention_editor_diffusers(pipe, ed)
ed.cur_step = 10
lat = torch.randn(1, 4, 64, 64, device=dev, dtype=torch.bfloat16)
for B in (1, 4):
    x = torch.cat([lat] * B)
    c = ctx[2:3] if B == 1 else ctx
    with torch.no_grad():
        for _ in range(3):
            pipe.unet(x, 501, encoder_hidden_states=c)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            pipe.unet(x, 501, encoder_hidden_states=c)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(e.device_time for e in evs)
    ours = [e for e in evs if "attn_tc" in e.name or "attn_mma" in e.name or "cross_attn" in e.name or "cfg_ddim" in e.name]
    print(f"B={B}: {len(evs)} kernels, {tot/1e3:.2f} ms device time; ours: {len(ours)} kernels {sum(e.device_time for e in ours)/1e3:.2f} ms")
    agg = {}
    for e in evs:
        agg[e.name[:60]] = agg.get(e.name[:60], 0) + e.device_time
    for n, t in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
        print(f"   {t/1e3:8.3f} ms  {n}")
