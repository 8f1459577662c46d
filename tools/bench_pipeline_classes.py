"""Full-size run (SD-1.5 geometry: 64x64 latents, 50 DDIM steps, bf16 stand-in UNet) of the pipeline-level classes that carry the
reference's names — the call a user of the reference's `*/model/sd_utils.py` makes: text conditioning -> controlled denoising loop ->
VAE decode -> uint8 images on the host. Eager mode (these classes do not use graph replay). Wall-clock per call, CUDA-synchronised.
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

    def case_p2p():
        ctrl = p2p.AttentionReplace(PROMPTS, pipe.tokenizer, STEPS, 0.8, 0.6, device=dev)
        try:
            return p2p.P2P(pipe, STEPS).text2image_ldm_stable(pipe, PROMPTS, ctrl, num_inference_steps=STEPS, guidance_scale=7.5, latent=lat)[0]
        finally:
            p2p.unregister_attention_control(pipe, ctrl)

    def case_masactrl():
        pipe.scheduler.set_timesteps(STEPS)
        trajectory, _ = ddim_inversion().ddim_inversion_loop(pipe, lat, PROMPTS[:1])       # B=1 inversion, as masactrl/edit_real.py does
        editor = masactrl.MutualSelfAttentionControl(4, 10, total_steps=STEPS)
        masactrl.regiter_attention_editor_diffusers(pipe, editor)
        try:
            start = trajectory[-1].expand(2, -1, -1, -1)
            return masactrl.MasaCtrl(pipe, STEPS)(PROMPTS, num_inference_steps=STEPS, guidance_scale=7.5, latents=start,
                                                  ref_intermediate_latents=trajectory)[0]
        finally:
            masactrl.unregister_attention_control(pipe, editor)

    def case_pnp():
        return pnp.PnP(pipe, STEPS)(PROMPTS, num_inference_steps=STEPS, guidance_scale=7.5, latents=lat, pnp_attn_t=0.5, pnp_f_t=0.8)

    def case_p2z():
        editor = pix2pix_zero.P2P_Zero(pipe, STEPS)
        try:
            return editor(PROMPTS, num_inference_steps=STEPS, guidance_scale=7.5, latents=lat.clone(), guidance_amount=0.1)[1]
        finally:
            pix2pix_zero.restore_original_processors(pipe.unet, editor.original_processors)

    out = {}
    for name, fn in (("P2P (AttentionReplace)", case_p2p), ("MasaCtrl (+ DDIM inversion)", case_masactrl), ("PnP", case_pnp),
                     ("P2P_Zero (sample + guided edit)", case_p2z)):
        with contextlib.redirect_stdout(io.StringIO()):
            fn()                                   # warm-up (cuDNN autotune, allocator)
            torch.cuda.synchronize()
            l0, t0 = _cabi.launch_count(), time.perf_counter()
            image = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        assert image.dtype.name == "uint8" and image.shape[1:] == (512, 512, 3), (image.dtype, image.shape)
        out[name] = {"s_per_call": round(dt, 3), "calls_per_s": round(1 / dt, 3), "ief_launches": _cabi.launch_count() - l0}
        print(name, out[name], flush=True)
    print(json.dumps({"model": "sd15 stand-in, bf16, channels_last, eager", "ddim_steps": STEPS, "results": out}))


if __name__ == "__main__":
    main()
