"""Full-size run (SD-1.5 geometry: 64x64 latents, 50 DDIM steps, bf16 stand-in UNet) of the pipeline-level classes that carry the
reference's names — the call a user of the reference's `*/model/sd_utils.py` makes: text conditioning -> controlled denoising loop ->
VAE decode -> uint8 images on the host. Wall-clock per call, CUDA-synchronised; eager (the classes' default) and with their opt-in
graphs=True on a kept instance + controller (captures happen in the warm-up calls).
    python tools/bench_pipeline_classes.py [ddim_steps]"""
import contextlib
import io
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import p2p, masactrl, pnp, pix2pix_zero, _cabi
from image_editing_framework_b200.ddim import ddim_inversion
from image_editing_framework_b200.standin import make_pipeline
from image_editing_framework_b200.standin.unet import sd15_config, AttnProcessor

dev = torch.device("cuda:0")
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 50
PROMPTS = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]


def main():
    AttnProcessor.use_sdpa = True   # un-hooked layers run what diffusers runs on a GPU
    cfg = sd15_config()
    with torch.device(dev):
        pipe = make_pipeline(cfg, seed=0, device=dev, dtype=torch.bfloat16)
    pipe.unet.to(memory_format=torch.channels_last)
    lat = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(0)).to(dev).to(torch.bfloat16)

    kept = {}

    def keep(key, make):
        key = (key, GRAPHS[0])
        if key not in kept:
            kept[key] = make()
        return kept[key]

    def case_p2p():
        ctrl = keep("p2p_ctrl", lambda: p2p.AttentionReplace(PROMPTS, pipe.tokenizer, STEPS, 0.8, 0.6, device=dev))
        ctrl.reset()
        try:
            return keep("p2p", lambda: p2p.P2P(pipe, STEPS, graphs=GRAPHS[0])).text2image_ldm_stable(pipe, PROMPTS, ctrl, num_inference_steps=STEPS, guidance_scale=7.5, latent=lat)[0]
        finally:
            p2p.unregister_attention_control(pipe, ctrl)

    def case_masactrl():
        pipe.scheduler.set_timesteps(STEPS)
        inv = ddim_inversion()
        inv.graphs = GRAPHS[0]
        trajectory, _ = inv.ddim_inversion_loop(pipe, lat, PROMPTS[:1])       # B=1 inversion, as masactrl/edit_real.py does
        editor = keep("masa_editor", lambda: masactrl.MutualSelfAttentionControl(4, 10, total_steps=STEPS))
        editor.reset()
        masactrl.regiter_attention_editor_diffusers(pipe, editor)
        try:
            start = trajectory[-1].expand(2, -1, -1, -1)
            return keep("masa", lambda: masactrl.MasaCtrl(pipe, STEPS, graphs=GRAPHS[0]))(PROMPTS, num_inference_steps=STEPS, guidance_scale=7.5, latents=start,
                                                  ref_intermediate_latents=trajectory)[0]
        finally:
            masactrl.unregister_attention_control(pipe, editor)

    def case_pnp():
        return keep("pnp", lambda: pnp.PnP(pipe, STEPS, graphs=GRAPHS[0]))(PROMPTS, num_inference_steps=STEPS, guidance_scale=7.5, latents=lat, pnp_attn_t=0.5, pnp_f_t=0.8)

    def case_p2z():
        editor = keep("p2z", lambda: pix2pix_zero.P2P_Zero(pipe, STEPS, graphs=GRAPHS[0]))
        try:
            return editor(PROMPTS, num_inference_steps=STEPS, guidance_scale=7.5, latents=lat.clone(), guidance_amount=0.1)[1]
        finally:
            pix2pix_zero.restore_original_processors(pipe.unet, editor.original_processors)

    out = {}
    GRAPHS = [False]
    for graphs in (False, True):
        GRAPHS[0] = graphs
        for name, fn in (("P2P (AttentionReplace)", case_p2p), ("MasaCtrl (+ DDIM inversion)", case_masactrl), ("PnP", case_pnp),
                         ("P2P_Zero (sample + guided edit)", case_p2z)):
            with contextlib.redirect_stdout(io.StringIO()):
                for _ in range(2 if graphs else 1):    # warm-up (cuDNN autotune, allocator; with graphs: first eager pass + captures)
                    fn()
                torch.cuda.synchronize()
                l0, t0 = _cabi.launch_count(), time.perf_counter()
                image = fn()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            assert image.dtype.name == "uint8" and image.shape[1:] == (512, 512, 3), (image.dtype, image.shape)
            name = name + (" graphs=True" if graphs else "")
            out[name] = {"s_per_call": round(dt, 3), "calls_per_s": round(1 / dt, 3), "ief_launches_outside_graphs": _cabi.launch_count() - l0}
            print(name, out[name], flush=True)
    print(json.dumps({"model": "sd15 stand-in, bf16, channels_last", "ddim_steps": STEPS, "results": out}))


if __name__ == "__main__":
    main()
