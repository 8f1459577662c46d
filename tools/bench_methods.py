"""Full-size (SD-1.5 geometry, 64x64 latents, UNet batch 4) run of every editing method through the public API, eager mode:
ms per 50-step edit pass, eager and with a persistent graphs.GraphedUNet runner (phase-keyed CUDA-graph replay; captures amortised: they happen in the warm-up edits). Secondary numbers next to bench.py's headline (MasaCtrl);
they also show that every controller works at the real sizes (the parity tests use small stand-ins).
    python tools/bench_methods.py [ddim_steps]"""
import contextlib
import io
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import p2p, masactrl, pnp, pix2pix_zero, editing, _cabi
from image_editing_framework_b200.standin import make_pipeline
from image_editing_framework_b200.standin.unet import sd15_config, sd21_config, sdxl_config

dev = torch.device("cuda:0")
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 50
MODEL = sys.argv[2] if len(sys.argv) > 2 else "sd15"   # sd15 (512^2) | sd21 (768^2, BASELINE config 3) | sdxl (1024^2, config 4)
PROMPTS = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def main():
    from image_editing_framework_b200.standin.unet import AttnProcessor
    AttnProcessor.use_sdpa = True   # un-hooked layers (PnP patches 8 of 32) run what diffusers 0.27 runs on a GPU
    cfg = {"sd15": sd15_config, "sd21": sd21_config, "sdxl": sdxl_config}[MODEL]()
    hw = cfg.sample_size
    with torch.device(dev):
        pipe = make_pipeline(cfg, seed=0, device=dev, dtype=torch.bfloat16)
    pipe.unet.to(memory_format=torch.channels_last)
    tok = pipe.tokenizer
    context = editing.encode_prompts(pipe, PROMPTS)
    g = torch.Generator().manual_seed(0)
    lat1 = torch.randn(1, 4, hw, hw, generator=g).to(dev).to(torch.bfloat16)
    lat2 = torch.cat([lat1, lat1])
    common = dict(prompts=PROMPTS, tokenizer=tok, num_steps=STEPS, cross_replace_steps=0.8, self_replace_steps=0.6, device=dev)

    from image_editing_framework_b200.graphs import GraphedUNet
    GRAPHS = [False]
    runners = {}

    def runner_for(name, ctrl, key_fn=None):
        # one persistent runner per case: captures happen in the warm-up edit, the timed edit only replays
        if not GRAPHS[0]:
            return False
        if name not in runners:
            runners[name] = GraphedUNet(pipe.unet, ctrl, key_fn, launch_counter=_cabi.launch_count)
        return runners[name]

    def p2p_run(make):
        box = {}

        def run():
            ctrl = box.setdefault("c", quiet(make))
            ctrl.reset()
            try:
                return editing.p2p_edit(pipe, PROMPTS, ctrl, lat1, STEPS, 7.5, context=context, graphs=runner_for(id(box), ctrl))
            finally:
                p2p.unregister_attention_control(pipe, ctrl)
        return run

    def masa_run(make):
        box = {}

        def run():
            ed = box.setdefault("c", quiet(make))
            ed.reset()
            masactrl.regiter_attention_editor_diffusers(pipe, ed)
            try:
                return editing.masactrl_edit(pipe, PROMPTS, lat2, STEPS, 7.5, context=context, graphs=runner_for(id(box), ed), editor=ed)
            finally:
                masactrl.unregister_attention_control(pipe, ed)
        return run

    def p2z_run():
        unet, originals = pix2pix_zero.prep_unet(pipe.unet)
        try:
            with torch.no_grad():
                pipe.scheduler.set_timesteps(STEPS)
                x = lat1
                ctx2 = torch.cat([context[:1], context[2:3]])
                for t in pipe.scheduler.timesteps.tolist():  # the map-collection pass of sd_utils.py:92-122 (maps stay on the device)
                    unet(torch.cat([x] * 2), t, encoder_hidden_states=ctx2)
                    _ = [m.attn_probs for n, m in unet.named_modules() if type(m).__name__ == "Attention" and "attn2" in n]
        finally:
            pix2pix_zero.restore_original_processors(unet, originals)
            for p_ in unet.parameters():
                p_.requires_grad = False

    def _lean(ctrl):
        ctrl.lean_store = True
        return ctrl

    eq = p2p.seq_aligner.get_equalizer(tok, PROMPTS[1], ("dog",), (3.0,))
    cases = [
        ("p2p EmptyControl (plain sampling)", p2p_run(lambda: p2p.EmptyControl(False))),
        ("p2p AttentionReplace 0.8/0.6", p2p_run(lambda: p2p.AttentionReplace(**common))),
        ("p2p AttentionRefine", p2p_run(lambda: p2p.AttentionRefine(**common))),
        ("p2p AttentionReweight(Replace)", p2p_run(lambda: p2p.AttentionReweight(equalizer=eq, controller=p2p.AttentionReplace(**common), **common))),
        ("p2p AttentionStore", p2p_run(lambda: p2p.AttentionStore(False))),
        ("p2p AttentionReplace + LocalBlend (store on)", p2p_run(lambda: p2p.AttentionReplace(
            local_blend=p2p.LocalBlend(tok, PROMPTS, [["cat"], ["dog"]], device=dev), **common))),
        ("p2p AttentionReplace + LocalBlend, lean_store (opt-in)", p2p_run(lambda: _lean(p2p.AttentionReplace(
            local_blend=p2p.LocalBlend(tok, PROMPTS, [["cat"], ["dog"]], device=dev), **common)))),
        ("masactrl MutualSelfAttentionControl(4, 10)", masa_run(lambda: masactrl.MutualSelfAttentionControl(4, 10, total_steps=STEPS))),
        ("masactrl Union", masa_run(lambda: masactrl.MutualSelfAttentionControlUnion(4, 10, total_steps=STEPS))),
        ("masactrl MaskAuto", masa_run(lambda: masactrl.MutualSelfAttentionControlMaskAuto(4, 10, total_steps=STEPS, ref_token_idx=[5], cur_token_idx=[5]))),
        ("pnp attn 0.5 / feature 0.8", lambda: editing.pnp_edit(pipe, PROMPTS, lat2, STEPS, 7.5, context=context, graphs=runner_for("pnp", None), xl=MODEL == "sdxl")),
        ("pix2pix-zero map-collection pass (B=2)", p2z_run),
    ]
    if MODEL != "sd15":  # LocalBlend / MaskAuto read 16x16 maps of the SD-1.5 topology; MasaCtrl's layer table is model specific
        keep = ("p2p EmptyControl", "p2p AttentionReplace 0.8/0.6", "p2p AttentionRefine", "pnp")
        cases = [c for c in cases if c[0].startswith(keep)]
        mt = "SDXL" if MODEL == "sdxl" else "SD"
        sl = 64 if MODEL == "sdxl" else 10   # SDXL: the decoder's self-attention layers 64..69 (N=4096)
        cases.append((f"masactrl MutualSelfAttentionControl(4, {sl}, model_type={mt})",
                      masa_run(lambda: masactrl.MutualSelfAttentionControl(4, sl, total_steps=STEPS, model_type=mt))))
    for name, run in cases:
        row = dict(method=name, model=MODEL, latent=hw, ddim_steps=STEPS)
        for graphs in (False, True):
            if graphs and name.startswith("pix2pix"):
                continue
            GRAPHS[0] = graphs
            run()  # warm-up (allocator, cuDNN autotune; with graphs: the captures)
            if graphs:
                run()  # phases seen once per edit (first stored step) are captured in the second edit
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            row["graphs_ms_per_edit_pass" if graphs else "eager_ms_per_edit_pass"] = round(dt * 1e3, 1)
        row["peak_mem_GB"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
        print(json.dumps(row), flush=True)
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
