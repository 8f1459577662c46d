"""The golden scenarios (tests/golden/make_goldens.py) replayed through THIS repo's register / controller / driver code.
Shared by the CPU host-logic tests (oracle-backed fake ops) and the GPU parity tests (real kernels)."""
import os

import torch

from image_editing_framework_b200 import p2p, masactrl, pnp, pix2pix_zero, editing
from image_editing_framework_b200.ddim import FusedDDIM
from image_editing_framework_b200.standin import make_pipeline, tiny_config
from image_editing_framework_b200.standin.unet import UNetConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name), map_location="cpu", weights_only=False)


class Recorder:
    def __init__(self, unet, keep):
        self.keep, self.step, self.records = set(keep), 0, {}
        for m in unet.modules():
            if type(m).__name__ == "Attention":
                m.forward = self._wrap(m.forward)

    def _wrap(self, inner):
        def fwd(*a, **kw):
            out = inner(*a, **kw)
            if self.step in self.keep:
                self.records.setdefault(self.step, []).append(out.detach().float().cpu())
            return out
        return fwd


def latent(seed, shape, device):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).to(device)


def psnr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    mse = ((a - b) ** 2).mean().item()
    peak = (b.max() - b.min()).item()
    return float("inf") if mse == 0 else 10 * torch.log10(torch.tensor(peak * peak / mse)).item()


def run_p2p(kind, g, device):
    pipe = make_pipeline(tiny_config(), seed=g["pipe_seed"], device=device)
    prompts, steps = g["prompts"], g["steps"]
    common = dict(prompts=prompts, tokenizer=pipe.tokenizer, num_steps=steps, cross_replace_steps=0.8, self_replace_steps=0.6, device=device)
    if kind == "replace":
        ctrl = p2p.AttentionReplace(**common)
    elif kind == "refine":
        ctrl = p2p.AttentionRefine(**common)
    elif kind == "reweight":
        eq = p2p.seq_aligner.get_equalizer(pipe.tokenizer, prompts[1], ("dog",), (3.0,))
        ctrl = p2p.AttentionReweight(equalizer=eq, controller=p2p.AttentionReplace(**common), **common)
    elif kind == "store":
        ctrl = p2p.AttentionStore(False)
    else:
        ctrl = p2p.EmptyControl(False)
    pipe.scheduler.set_timesteps(steps)
    p2p.register_attention_control(pipe, ctrl)
    rec = Recorder(pipe.unet, g["keep"])
    context = editing.encode_prompts(pipe, prompts)
    hw = g["latent_hw"]
    latents = latent(g["latent_seed"], (1, 4, hw, hw), device).expand(2, 4, hw, hw).contiguous()
    fused = FusedDDIM(pipe.scheduler)
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(pipe.scheduler.timesteps.tolist()):
            rec.step = i
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context)["sample"]
            latents = ctrl.step_callback(fused.step(noise, t, latents, g["guidance"]))
            per_step.append(latents.float().cpu())
    return ctrl, rec.records, per_step


def run_masactrl(g, device):
    pipe = make_pipeline(tiny_config(), seed=g["pipe_seed"], device=device)
    steps = g["steps"]
    pipe.scheduler.set_timesteps(steps)
    ctrl = masactrl.MutualSelfAttentionControl(g["start_step"], g["start_layer"], total_steps=steps)
    masactrl.regiter_attention_editor_diffusers(pipe, ctrl)
    rec = Recorder(pipe.unet, g["keep"])
    context = editing.encode_prompts(pipe, g["prompts"])
    hw = g["latent_hw"]
    init = latent(g["latent_seed"], (1, 4, hw, hw), device)
    latents = torch.cat([init, init])
    fused = FusedDDIM(pipe.scheduler)
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(pipe.scheduler.timesteps.tolist()):
            rec.step = i
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            latents = fused.step(noise, t, latents, g["guidance"])
            per_step.append(latents.float().cpu())
    return ctrl, rec.records, per_step


def run_masactrl_masks(g, device, which):
    """which: 'mask' | 'mask_auto' | 'mutual' — the masked MasaCtrl variants on the 32x32-latent stand-in of masactrl_masks.pt."""
    pipe = make_pipeline(UNetConfig(**g["config"]), seed=g["pipe_seed"], device=device)
    steps = g["steps"]
    pipe.scheduler.set_timesteps(steps)
    if which == "mask":
        ctrl = masactrl.MutualSelfAttentionControlMask(g["start_step"], g["start_layer"], total_steps=steps, mask_s=g["mask_s"], mask_t=g["mask_t"])
    elif which == "mask_auto":
        ctrl = masactrl.MutualSelfAttentionControlMaskAuto(g["start_step"], g["start_layer"], total_steps=steps, thres=g["thres"],
                                                          ref_token_idx=g["ref_token_idx"], cur_token_idx=g["cur_token_idx"])
    else:
        ctrl = masactrl.MutualSelfAttentionControl(g["start_step"], g["start_layer"], total_steps=steps)
    masactrl.regiter_attention_editor_diffusers(pipe, ctrl)
    context = editing.encode_prompts(pipe, g["prompts"])
    hw = g["latent_hw"]
    latents = torch.cat([latent(g["latent_seed"], (1, 4, hw, hw), device), latent(g["latent_seed"] + 1, (1, 4, hw, hw), device)])
    fused = FusedDDIM(pipe.scheduler)
    rec = RowRecorder(pipe.unet, (steps - 1,))
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(pipe.scheduler.timesteps.tolist()):
            rec.step = i
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            latents = fused.step(noise, t, latents, g["guidance"])
            per_step.append(latents.float().cpu())
    ctrl.layer_rows = rec.records[steps - 1]
    return ctrl, per_step


def run_pnp(g, device, xl=False):
    pipe = make_pipeline(UNetConfig(**g["config"]) if xl else tiny_config(), seed=g["pipe_seed"], device=device)
    steps = g["steps"]
    pipe.scheduler.set_timesteps(steps)
    ts = pipe.scheduler.timesteps
    if xl:
        pnp.register_attention_control_efficient_xl(pipe, ts[:int(steps * g["pnp_attn_t"])])
        pnp.register_conv_control_efficient_xl(pipe, ts[:int(steps * g["pnp_f_t"])])
    else:
        pnp.register_attention_control_efficient(pipe, ts[:int(steps * g["pnp_attn_t"])])
        pnp.register_conv_control_efficient(pipe, ts[:int(steps * g["pnp_f_t"])])
    rec = Recorder(pipe.unet, g["keep"])
    context = editing.encode_prompts(pipe, g["prompts"])
    hw = g["latent_hw"]
    init = latent(g["latent_seed"], (1, 4, hw, hw), device)
    latents = torch.cat([init, init])
    fused = FusedDDIM(pipe.scheduler)
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(ts.tolist()):
            rec.step = i
            (pnp.register_time_xl if xl else pnp.register_time)(pipe, t)
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            latents = fused.step(noise, t, latents, g["guidance"])
            per_step.append(latents.float().cpu())
    return rec.records, per_step


def run_pix2pix_zero(g, device):
    pipe = make_pipeline(tiny_config(), seed=g["pipe_seed"], device=device)
    unet, originals = pix2pix_zero.prep_unet(pipe.unet)
    rec = Recorder(unet, (0,))
    context = editing.encode_prompts(pipe, g["prompts"])
    hw = g["latent_hw"]
    x = latent(g["latent_seed"], (1, 4, hw, hw), device)
    with torch.no_grad():
        out = unet(torch.cat([x] * 2), g["t"], encoder_hidden_states=context).sample
    probs = {n: m.attn_probs.float().cpu() for n, m in unet.named_modules() if type(m).__name__ == "Attention" and "attn2" in n}
    return out.float().cpu(), rec.records[0], probs, (unet, originals)


def run_pix2pix_zero_loop(g, device, guidance_amount=None, graphs=False):
    pipe = make_pipeline(tiny_config(), seed=g["pipe_seed"], device=device)
    unet, originals = pix2pix_zero.prep_unet(pipe.unet)
    emb_src = editing.encode_prompts(pipe, g["prompts"][:1])
    emb_edit = editing.encode_prompts(pipe, g["prompts"][1:])
    hw = g["latent_hw"]
    lat0 = latent(g["latent_seed"], (1, 4, hw, hw), device)
    rec, edit = editing.pix2pix_zero_edit(pipe, emb_src, emb_edit, lat0, g["steps"], g["guidance"],
                                          g["guidance_amount"] if guidance_amount is None else guidance_amount, graphs=graphs)
    pix2pix_zero.restore_original_processors(unet, originals)
    return rec.float().cpu(), edit.float().cpu()


def run_p2p_localblend(g, device, lean_store=False):
    cfg = UNetConfig(**g["config"])
    pipe = make_pipeline(cfg, seed=g["pipe_seed"], device=device)
    prompts, steps = g["prompts"], g["steps"]
    lb = p2p.LocalBlend(pipe.tokenizer, prompts, g["blend_words"], threshold=g["threshold"], device=device)
    lb.masks = []
    ctrl = p2p.AttentionReplace(prompts, pipe.tokenizer, steps, 0.8, 0.6, lb, device=device)
    ctrl.lean_store = lean_store
    pipe.scheduler.set_timesteps(steps)
    p2p.register_attention_control(pipe, ctrl)
    context = editing.encode_prompts(pipe, prompts)
    latents = torch.cat([latent(s_, (1, 4, 64, 64), device) for s_ in g["latent_seeds"]])
    fused = FusedDDIM(pipe.scheduler)
    per_step = []
    with torch.no_grad():
        for t in pipe.scheduler.timesteps.tolist():
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context)["sample"]
            latents = ctrl.step_callback(fused.step(noise, t, latents, g["guidance"]))
            per_step.append(latents.float().cpu())
    maps = ctrl.attention_store["down_cross"][2:4] + ctrl.attention_store["up_cross"][:3]
    ctrl.blend_masks = [m.float().cpu() for m in lb.masks]
    return ctrl, per_step, [m.float().cpu() for m in maps]


# --------------------------------------------------------------------------------- pipeline-level classes (`*/model/sd_utils.py`)
PIPELINE_PROMPTS = ["a photo of a cat sitting on the bench", "a photo of a dog sitting on the bench"]
PIPELINE_CASES = [(family, base + suffix) for family, base in (("p2p", "P2P"), ("masactrl", "MasaCtrl"), ("pnp", "PnP"), ("pix2pix-zero", "P2P_Zero"))
                  for suffix in ("", "_NTI", "_XL", "_XL_NTI")]
TINY_XL = dict(sample_size=8, block_out_channels=(32, 64, 64), transformer_layers=(0, 2, 3), num_heads=(2, 2, 2),
               cross_attention_dim=32, norm_num_groups=8, use_linear_projection=True, name="tiny_xl")


class XLPipelineDouble:
    """The StableDiffusionXLPipeline members the XL drivers touch, over the stand-in pipeline. added_cond_kwargs are recorded (the
    stand-in UNet has no add-embedding); the pooled embedding is the token mean."""

    def __init__(self, seed, config, device):
        self._p = make_pipeline(config, seed=seed, device=device)
        for name in ("unet", "scheduler", "vae", "tokenizer", "text_encoder"):
            setattr(self, name, getattr(self._p, name))
        self.added = []
        inner = self.unet.forward

        def forward(sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None, **kw):
            self.added.append(added_cond_kwargs)
            return inner(sample, timestep, encoder_hidden_states)
        self.unet.forward = forward

    device = _execution_device = property(lambda self: self.unet.device)

    def __getattr__(self, name):             # prepare_latents, progress_bar, vae_scale_factor, ... come from the stand-in pipeline
        return getattr(self.__dict__["_p"], name)

    def encode_prompt(self, prompt, device, do_classifier_free_guidance=True, **kw):
        pe, ne = self._p.encode_prompt(prompt, device)
        return pe, ne, pe.mean(1), ne.mean(1)

    def _get_add_time_ids(self, original_size, crops, target_size, dtype):
        return torch.tensor([list(original_size) + list(crops) + list(target_size)], dtype=dtype)


def null_text_embeddings(steps, seed, device, dim=32):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(1, 77, dim, generator=g) * 0.1).to(device) for _ in range(steps)]


def run_pipeline_case(family, name, api, device):
    """One fixed scenario per pipeline-level class, written against `api` = the reference's modules of that family
    (oracle.reference_loader) or this package's mirror of them — the class and hook names are the same on both sides.
    Returns the uint8 images the class produces (a list) and the pipeline it ran on."""
    xl, nti = "XL" in name, "NTI" in name
    # the *_xl PnP hook tables address SDXL's block topology (3 blocks, no attention in the first)
    cfg = UNetConfig(**TINY_XL) if (xl and family == "pnp") else tiny_config()
    seed = {"p2p": 7, "masactrl": 9, "pnp": 13, "pix2pix-zero": 17}[family]
    pipe = XLPipelineDouble(seed, cfg, device) if xl else make_pipeline(cfg, seed=seed, device=device)
    g = torch.Generator().manual_seed(100 + seed)
    lat = torch.randn(1, 4, 8, 8, generator=g).to(device)
    cls = getattr(api.sd_utils, name)
    if family == "p2p":
        steps = 4
        ctrl = api.attention_control.AttentionReplace(prompts=PIPELINE_PROMPTS, tokenizer=pipe.tokenizer, num_steps=steps,
                                                      cross_replace_steps=0.8, self_replace_steps=0.5, device=device)
        editor = cls(pipe, steps)
        # both sides hard-code 512^2 / 1024^2: keep their loop but start from an 8x8 latent so a CPU run stays small
        editor.init_latent = lambda latent, model, h, w, gen, bs: (latent, latent.expand(bs, 4, 8, 8))
        kw = {"uncond_embeddings_list": null_text_embeddings(steps, 12, device)} if nti else {}
        image, x_t = editor.text2image_ldm_stable(pipe, PIPELINE_PROMPTS, ctrl, num_inference_steps=steps, guidance_scale=7.5, latent=lat, **kw)
        assert ctrl.cur_step == steps and torch.equal(x_t, lat)
        return [image], pipe
    if family == "masactrl":
        steps = 4
        trajectory = [torch.randn(1, 4, 8, 8, generator=g).to(device) for _ in range(steps + 1)]   # stands for the inversion's latents
        editor = api.attention_control.MutualSelfAttentionControl(1, 10, total_steps=steps)
        api.register.regiter_attention_editor_diffusers(pipe, editor)
        kw = dict(height=64, width=64, num_inference_steps=steps, guidance_scale=7.5, latents=torch.cat([lat, lat]),
                  ref_intermediate_latents=trajectory)
        if nti:
            kw["uncond_embeddings_list"] = null_text_embeddings(steps, 22, device)
        elif not xl:
            kw.update(unconditioning=null_text_embeddings(steps, 22, device), neg_prompt="blurry")
        image, x_t = cls(pipe, steps)(PIPELINE_PROMPTS, **kw)
        assert editor.cur_step == steps and torch.equal(x_t, torch.cat([lat, lat]))
        return [image], pipe
    if family == "pnp":
        steps = 5
        kw = {"uncond_embeddings_list": null_text_embeddings(steps, 32, device)} if nti else {}
        image = cls(pipe, steps)(PIPELINE_PROMPTS, height=64, width=64, num_inference_steps=steps, guidance_scale=7.5, latents=lat,
                                 pnp_attn_t=0.5, pnp_f_t=0.8, **kw)
        return [image], pipe
    steps = 3
    edit_dir = (torch.randn(1, 77, 32, generator=g) * 0.05).to(device)
    kw = {"uncond_embeddings_list": null_text_embeddings(steps, 42, device)} if nti else {}
    rec, edit = cls(pipe, steps)(PIPELINE_PROMPTS, height=64, width=64, num_inference_steps=steps, guidance_scale=7.5, latents=lat.clone(),
                                 guidance_amount=0.1, edit_dir=edit_dir, **kw)
    return [rec, edit], pipe


def mirror_api(family):
    """This package's modules under the attribute names the reference's `model` package uses."""
    from types import SimpleNamespace
    import image_editing_framework_b200 as pkg
    mod = {"p2p": pkg.p2p, "masactrl": pkg.masactrl, "pnp": pkg.pnp, "pix2pix-zero": pkg.pix2pix_zero}[family]
    return SimpleNamespace(sd_utils=mod, attention_control=mod, register=mod)


# --------------------------------------------------------------------------------- full-geometry scenarios (64x64 latents, real head dims)
# Slim-channel stand-ins whose ATTENTION geometry is the real one: 64x64 latents (N = 4096 / 1024 / 256 / 64 tokens) and the real
# head dims — SD-1.5's 40 / 80 / 160 / 160, and 64 everywhere like SD-2.1 / SDXL — with 2-8 heads instead of 8-20 so that the reference's
# materialised fp32 probabilities fit a CPU run. Large layers therefore take the tcgen05 kernels, small ones the mma.sync ones, exactly
# as at full width.
FULLGEO_CONFIGS = {
    "sd15_slim": dict(sample_size=64, block_out_channels=(80, 160, 320, 320), num_heads=(2, 2, 2, 2), cross_attention_dim=64,
                      norm_num_groups=8, name="sd15_slim"),
    "d64_slim": dict(sample_size=64, block_out_channels=(128, 256, 512, 512), num_heads=(2, 4, 8, 8), cross_attention_dim=64,
                     norm_num_groups=8, use_linear_projection=True, name="d64_slim"),
}
FULLGEO_CASES = [("sd15_slim", "p2p_replace"), ("sd15_slim", "p2p_refine"), ("sd15_slim", "p2p_store"), ("sd15_slim", "masactrl"),
                 ("sd15_slim", "masactrl_union"), ("sd15_slim", "pnp"), ("d64_slim", "p2p_refine"), ("d64_slim", "masactrl"), ("d64_slim", "pnp")]
FULLGEO_STEPS, FULLGEO_ROWS, FULLGEO_GUIDANCE = 3, 16, 7.5
FULLGEO_PROMPTS = {"p2p_replace": ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"],
                   "p2p_refine": ["a bowl of soup", "a bowl of pea soup"],
                   "p2p_store": ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"],
                   "masactrl": ["a photo of a sitting cat", "a photo of a running cat"],
                   "masactrl_union": ["a photo of a sitting cat", "a photo of a running cat"],
                   "pnp": ["a photo of a wooden horse", "a photo of a bronze horse"]}


class RowRecorder:
    """Like Recorder, but keeps FULLGEO_ROWS evenly spaced token rows of every attention output (a 64x64 layer's full output is 5 MB)."""

    def __init__(self, unet, keep):
        self.keep, self.step, self.records, self.tokens = set(keep), 0, {}, []
        for m in unet.modules():
            if type(m).__name__ == "Attention":
                m.forward = self._wrap(m.forward)

    def _wrap(self, inner):
        def fwd(*a, **kw):
            out = inner(*a, **kw)
            if self.step in self.keep:
                n = out.shape[1]
                rows = torch.linspace(0, n - 1, min(n, FULLGEO_ROWS)).long().to(out.device)
                self.records.setdefault(self.step, []).append(out.detach().index_select(1, rows).float().cpu())
            return out
        return fwd


def run_fullgeo(cfg_name, kind, api, device, fused_step=True):
    """One scenario written against `api` = the reference's modules of the family (oracle.reference_loader, CPU fp32: the golden
    generator) or this package's mirror (the GPU test). Loop shape: p2p/model/sd_utils.py:67-79 (diffusion_step), masactrl/model/
    sd_utils.py:107-113, pnp/model/sd_utils.py:99-107. Returns (controller or None, {step: [rows of each layer output]}, latents per step)."""
    pipe = make_pipeline(UNetConfig(**FULLGEO_CONFIGS[cfg_name]), seed=21, device=device)
    steps, prompts = FULLGEO_STEPS, FULLGEO_PROMPTS[kind]
    pipe.scheduler.set_timesteps(steps)
    hw = 64
    ctrl = None
    family = kind.split("_")[0]
    if family == "p2p":
        common = dict(prompts=prompts, tokenizer=pipe.tokenizer, num_steps=steps, cross_replace_steps=0.8, self_replace_steps=0.6, device=device)
        ctrl = {"p2p_replace": lambda: api.attention_control.AttentionReplace(**common),
                "p2p_refine": lambda: api.attention_control.AttentionRefine(**common),
                "p2p_store": lambda: api.attention_base.AttentionStore(False)}[kind]()
        driver = api.sd_utils.P2P(pipe, steps)
        api.register.register_attention_control(pipe, ctrl)
    elif family == "masactrl":
        # (the reference's Union class is driven through its minimally repaired subclass, see make_goldens.py::repaired_union)
        cls = {"masactrl": "MutualSelfAttentionControl", "masactrl_union": "MutualSelfAttentionControlUnion"}[kind]
        ctrl = getattr(api.attention_control, cls + "Repaired", getattr(api.attention_control, cls))(1, 10, total_steps=steps)
        api.register.regiter_attention_editor_diffusers(pipe, ctrl)
    else:
        ts = pipe.scheduler.timesteps
        api.register.register_attention_control_efficient(pipe, ts[:int(steps * 0.5)])
        api.register.register_conv_control_efficient(pipe, ts[:int(steps * 0.8)])
    rec = RowRecorder(pipe.unet, (steps - 1,))
    context = editing.encode_prompts(pipe, prompts)
    # source and target rows start from DIFFERENT latents (the scripts start both from the inverted image): on a random-init UNet
    # identical rows stay nearly identical for a few steps, K_src ~ K_tgt, and a broken edit would pass unnoticed. With distinct rows
    # the controlled results are far from the uncontrolled ones (make_goldens.py records how far), so the gates below discriminate.
    latents = torch.cat([latent(31, (1, 4, hw, hw), device), latent(32, (1, 4, hw, hw), device)])
    fused = FusedDDIM(pipe.scheduler) if fused_step else None
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(pipe.scheduler.timesteps):
            rec.step = i
            if family == "p2p":
                latents = driver.diffusion_step(pipe, ctrl, latents, context, t if not fused_step else int(t), FULLGEO_GUIDANCE, False)
            else:
                if family == "pnp":
                    api.register.register_time(pipe, t.item())
                noise = pipe.unet(torch.cat([latents] * 2), t if not fused_step else int(t), encoder_hidden_states=context).sample
                if fused_step:
                    latents = fused.step(noise, int(t), latents, FULLGEO_GUIDANCE)
                else:
                    nu, nc = noise.chunk(2)
                    latents = pipe.scheduler.step(nu + FULLGEO_GUIDANCE * (nc - nu), t, latents).prev_sample
            per_step.append(latents.float().cpu())
    return ctrl, rec.records, per_step
