"""Generate the committed golden fixtures by running the REFERENCE's own Python (read-only, /root/reference)
on the stand-in pipeline, CPU fp32, fixed seeds.  Run in the build container only:

    python tests/golden/make_goldens.py

The reference ships no tests or golden vectors, so these files are what pins the oracle
(oracle/controlled_attention.py) and the product's host-side integer code (seq_aligner / ptp_utils).
Everything a test needs to rebuild the inputs (config, seeds, prompts) is stored next to the outputs.
Compositions the reference cannot execute as shipped (AttentionStore + edit, LocalBlend — SURVEY.md facts 0.5-0.7)
are generated from the reference classes re-composed the upstream Prompt-to-Prompt way and are labelled "recomposed".
"""
import dataclasses
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402
from image_editing_framework_b200.standin import make_pipeline, tiny_config, WordPieceTokenizer, DDIMScheduler  # noqa: E402
from image_editing_framework_b200.standin.unet import UNetConfig  # noqa: E402

CPU = torch.device("cpu")
LAT = 8  # golden latents are 8x8 (tokens 64/16/4/1) to keep the committed fixtures small

PROMPT_CASES = [
    # (source, target, kind)
    ("a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench", "replace"),
    ("a squirrel eating a burger", "a lion eating a burger", "replace"),            # 3-token word -> 1-token word: ratio branch
    ("a cat on a table", "a hippopotamus on a table", "replace"),                   # 1 token -> 3 tokens
    ("a painting of a house", "a watercolor painting of a large house by the lake", "refine"),
    ("a bowl of soup", "a bowl of pea soup", "refine"),
    ("children playing near the river at sunset", "children near the frozen river", "refine"),  # deletions and insertions
    ("a", "a", "replace"),
]


def _save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ------------------------------------------------------------------------------------------------ host-side integer code
def gen_aligner():
    ref = load_reference("p2p")
    tok = WordPieceTokenizer()
    cases = []
    for src, tgt, kind in PROMPT_CASES:
        c = dict(source=src, target=tgt, kind=kind, src_ids=tok.encode(src), tgt_ids=tok.encode(tgt))
        if kind == "replace":
            c["replacement_mapper"] = ref.seq_aligner.get_replacement_mapper([src, tgt], tok)
        c["refinement_mapper"], c["refinement_alphas"] = ref.seq_aligner.get_refinement_mapper([src, tgt], tok)
        words = tgt.split(" ")
        c["word_inds"] = {w: ref.seq_aligner.get_word_inds(tgt, w, tok).tolist() for w in set(words)}
        c["word_inds_by_pos"] = [ref.seq_aligner.get_word_inds(tgt, i, tok).tolist() for i in range(len(words))]
        c["equalizer"] = ref.seq_aligner.get_equalizer(tok, tgt, (words[-1],), (2.0,))
        c["alpha_float"] = ref.ptp_utils.get_time_words_attention_alpha([src, tgt], 10, 0.8, tok)
        c["alpha_dict"] = ref.ptp_utils.get_time_words_attention_alpha([src, tgt], 10, {"default_": 1.0, words[-1]: (0.2, 0.6)}, tok)
        cases.append(c)
    multi = ["a photo of a cat", "a photo of a dog", "a photo of a fox"]
    extra = dict(prompts=multi, replacement=ref.seq_aligner.get_replacement_mapper(multi, tok),
                 refinement=ref.seq_aligner.get_refinement_mapper(multi, tok),
                 alpha=ref.ptp_utils.get_time_words_attention_alpha(multi, 50, (0.1, 0.7), tok))
    _save("aligner.pt", dict(cases=cases, multi=extra))


# ------------------------------------------------------------------------------------------------ per-layer recording
class Recorder:
    """Wraps the forward of every patched Attention so that each call's output is recorded in call order."""

    def __init__(self, unet, steps_to_keep):
        self.keep, self.step, self.records = set(steps_to_keep), 0, {}
        self.mods = [m for m in unet.modules() if type(m).__name__ == "Attention"]
        for m in self.mods:
            inner = m.forward
            m.forward = self._wrap(inner)

    def _wrap(self, inner):
        def fwd(*a, **kw):
            out = inner(*a, **kw)
            if self.step in self.keep:
                self.records.setdefault(self.step, []).append(out.detach().clone())
            return out
        return fwd


def _latent(seed, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def gen_p2p():
    ref = load_reference("p2p")
    out = {}
    steps, keep = 5, (1, 4)
    for name, kind in (("replace", "replace"), ("refine", "refine"), ("reweight", "reweight"), ("store", "store"), ("empty", "empty")):
        pipe = make_pipeline(tiny_config(), seed=0)
        src, tgt = ("a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench") if kind != "refine" else \
                   ("a bowl of soup", "a bowl of pea soup")
        prompts = [src, tgt]
        common = dict(prompts=prompts, tokenizer=pipe.tokenizer, num_steps=steps, cross_replace_steps=0.8, self_replace_steps=0.6, device=CPU)
        if kind == "replace":
            ctrl = ref.attention_control.AttentionReplace(**common)
        elif kind == "refine":
            ctrl = ref.attention_control.AttentionRefine(**common)
        elif kind == "reweight":
            eq = ref.seq_aligner.get_equalizer(pipe.tokenizer, tgt, ("dog",), (3.0,))
            prev = ref.attention_control.AttentionReplace(**common)
            ctrl = ref.attention_control.AttentionReweight(equalizer=eq, controller=prev, **common)
        elif kind == "store":
            ctrl = ref.attention_base.AttentionStore(False)
        else:
            ctrl = ref.attention_base.EmptyControl(False)
        editor = ref.sd_utils.P2P(pipe, steps)
        ref.register.register_attention_control(pipe, ctrl)
        rec = Recorder(pipe.unet, keep)
        context = _context(pipe, prompts)
        latent = _latent(1, (1, 4, LAT, LAT))
        latents = latent.expand(2, 4, LAT, LAT)
        per_step = []
        with torch.no_grad():
            for i, t in enumerate(pipe.scheduler.timesteps):
                rec.step = i
                latents = editor.diffusion_step(pipe, ctrl, latents, context, t, 7.5, False)
                per_step.append(latents.clone())
        if kind in ("store", "empty"):
            rec.records = {}
        g = dict(prompts=prompts, steps=steps, keep=keep, latent_seed=1, pipe_seed=0, guidance=7.5, layer_outputs=rec.records, latent_hw=LAT,
                 latents_per_step=per_step, num_att_layers=ctrl.num_att_layers, cur_step=ctrl.cur_step)
        if kind == "store":
            avg = ctrl.get_average_attention()
            g["average_attention"] = {k: [t.clone() for t in v] for k, v in avg.items()}
        out[name] = g
    _save("p2p.pt", out)


def _context(pipe, prompts):
    tok = pipe.tokenizer(prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt")
    cond = pipe.text_encoder(tok.input_ids)[0]
    un = pipe.tokenizer([""] * len(prompts), padding="max_length", max_length=77, return_tensors="pt")
    return torch.cat([pipe.text_encoder(un.input_ids)[0], cond])


def gen_p2p_localblend():
    """RECOMPOSED oracle: AttentionReplace + AttentionStore + LocalBlend the upstream way (the reference classes crash as
    shipped). 64x64 latents so that the stored maps LocalBlend reads are 16x16."""
    ref = load_reference("p2p")
    AB, AC = ref.attention_base, ref.attention_control

    class ReplaceWithStore(AC.AttentionReplace):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            self.step_store = AB.AttentionStore.get_empty_store()
            self.attention_store = {}

        def forward(self, attn, is_cross, place_in_unet):
            AB.AttentionStore.forward(self, attn, is_cross, place_in_unet)
            return super().forward(attn, is_cross, place_in_unet)

        between_steps = AB.AttentionStore.between_steps
        get_empty_store = staticmethod(AB.AttentionStore.get_empty_store)

    cfg = UNetConfig(sample_size=64, block_out_channels=(16, 32, 32, 32), num_heads=(2, 2, 2, 2), cross_attention_dim=32, norm_num_groups=8, name="tiny64")
    pipe = make_pipeline(cfg, seed=3)
    prompts = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]
    steps = 3
    from oracle import controlled_attention as orc

    class RecordingBlend(ref.ptp_utils.LocalBlend):
        """The reference's LocalBlend; every call also records the per-prompt masks, taken from the oracle's restatement on the same
        inputs after checking that its blended latents equal the reference's bit for bit."""
        masks = []

        def __call__(self, x_t, attention_store):
            out = super().__call__(x_t, attention_store)
            maps = attention_store["down_cross"][2:4] + attention_store["up_cross"][:3]
            again, mask = orc.local_blend(x_t, maps, self.alpha_layers.reshape(len(prompts), -1), self.threshold, return_mask=True)
            assert torch.equal(again, out)
            self.masks.append(mask.clone())
            return out

    # threshold 0.65 and distinct source / target latents: about half of the latent pixels are blended (with the default 0.3 and
    # identical rows the mask of this random-init UNet is 98 % ones and the blend is a near no-op, i.e. nothing is tested)
    threshold = 0.65
    lb = RecordingBlend(pipe.tokenizer, prompts, [["cat"], ["dog"]], threshold=threshold, device=CPU)
    ctrl = ReplaceWithStore(prompts, pipe.tokenizer, steps, 0.8, 0.6, lb, device=CPU)
    editor = ref.sd_utils.P2P(pipe, steps)
    ref.register.register_attention_control(pipe, ctrl)
    context = _context(pipe, prompts)
    latents = torch.cat([_latent(5, (1, 4, 64, 64)), _latent(6, (1, 4, 64, 64))])
    per_step = []
    with torch.no_grad():
        for t in pipe.scheduler.timesteps:
            latents = editor.diffusion_step(pipe, ctrl, latents, context, t, 7.5, False)
            per_step.append(latents.clone())
    maps = ctrl.attention_store["down_cross"][2:4] + ctrl.attention_store["up_cross"][:3]
    print("   LocalBlend mask coverage per step:", [round(m.float().mean().item(), 3) for m in lb.masks])
    _save("p2p_localblend.pt", dict(recomposed=True, prompts=prompts, steps=steps, latent_seeds=(5, 6), threshold=threshold, pipe_seed=3, guidance=7.5,
                                    config=dataclasses.asdict(cfg), blend_words=[["cat"], ["dog"]], latents_per_step=per_step,
                                    store_16=[m.clone() for m in maps], masks_per_step=[m.to(torch.uint8) for m in lb.masks]))
    # stand-alone LocalBlend known-answer on synthetic maps
    g = torch.Generator().manual_seed(8)
    syn = [torch.rand(16, 256, 77, generator=g) ** 4 for _ in range(5)]
    x = torch.randn(2, 4, 64, 64, generator=g)
    store = {"down_cross": [None, None, syn[0], syn[1]], "up_cross": syn[2:]}
    _save("local_blend.pt", dict(seed=8, x_out=lb(x, store), alpha_layers=lb.alpha_layers.clone()))


def gen_masactrl():
    ref = load_reference("masactrl")
    pipe = make_pipeline(tiny_config(), seed=1)
    steps, keep = 4, (0, 2)
    prompts = ["a photo of a sitting cat", "a photo of a running cat"]
    editor = ref.sd_utils.MasaCtrl(pipe, steps)
    ctrl = ref.attention_control.MutualSelfAttentionControl(1, 10, total_steps=steps)
    ref.register.regiter_attention_editor_diffusers(pipe, ctrl)
    rec = Recorder(pipe.unet, keep)
    context = _context(pipe, prompts)
    init = _latent(2, (1, 4, LAT, LAT))
    latents = torch.cat([init, init])
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(pipe.scheduler.timesteps):
            rec.step = i
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            nu, nc = noise.chunk(2, dim=0)
            latents = pipe.scheduler.step(nu + 7.5 * (nc - nu), t, latents, return_dict=True)["prev_sample"]  # sd_utils.py:107-113
            per_step.append(latents.clone())
    _save("masactrl.pt", dict(prompts=prompts, steps=steps, keep=keep, latent_seed=2, pipe_seed=1, guidance=7.5, start_step=1, start_layer=10,
                              layer_outputs=rec.records, latents_per_step=per_step, num_att_layers=ctrl.num_att_layers, latent_hw=LAT))


def _masa_loop(pipe, ctrl, prompts, latent_seed, hw):
    """-> (latents per step, {last step: 16 token rows of every attention layer's output})"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenarios
    context = _context(pipe, prompts)
    # distinct source / target rows: with identical rows the masked variants differ from plain mutual control by less than bf16 noise
    latents = torch.cat([_latent(latent_seed, (1, 4, hw, hw)), _latent(latent_seed + 1, (1, 4, hw, hw))])
    rec = scenarios.RowRecorder(pipe.unet, (len(pipe.scheduler.timesteps) - 1,))
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(pipe.scheduler.timesteps):
            rec.step = i
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            nu, nc = noise.chunk(2, dim=0)
            latents = pipe.scheduler.step(nu + 7.5 * (nc - nu), t, latents, return_dict=True)["prev_sample"]
            per_step.append(latents.clone())
    return per_step, rec.records


def gen_masactrl_masks():
    """MutualSelfAttentionControlMask / MaskAuto (masactrl/model/attention_control.py:110-326) through the reference's own register
    closure, on a 32x32-latent stand-in (so 16x16 cross-attention layers exist for MaskAuto)."""
    ref = load_reference("masactrl")
    cfg = UNetConfig(sample_size=32, block_out_channels=(16, 32, 32, 32), num_heads=(2, 2, 2, 2), cross_attention_dim=32, norm_num_groups=8, name="tiny32")
    steps, hw = 3, 32
    prompts = ["a photo of a sitting cat", "a photo of a running cat"]
    mask_s = torch.zeros(64, 64)
    mask_s[10:44, 16:50] = 1
    mask_t = torch.zeros(64, 64)
    mask_t[20:60, 8:40] = 1
    out = dict(prompts=prompts, steps=steps, latent_seed=9, pipe_seed=4, guidance=7.5, start_step=1, start_layer=10, latent_hw=hw,
               config=dataclasses.asdict(cfg), mask_s=mask_s, mask_t=mask_t, thres=0.3, ref_token_idx=[5], cur_token_idx=[5])
    for name in ("mask", "mask_auto"):
        pipe = make_pipeline(cfg, seed=4)
        ref.sd_utils.MasaCtrl(pipe, steps)  # sets the timesteps the reference way
        if name == "mask":
            ctrl = ref.attention_control.MutualSelfAttentionControlMask(1, 10, total_steps=steps, mask_s=mask_s, mask_t=mask_t)
        else:
            ctrl = ref.attention_control.MutualSelfAttentionControlMaskAuto(1, 10, total_steps=steps, thres=0.3, ref_token_idx=[5], cur_token_idx=[5])
            # (0.3: a fifth of the positions is foreground and < 10 % of them lie within 0.03 of the threshold; at 0.1 a quarter does,
            #  and the min-max normalised maps of this random-init UNet carry about that much bf16 noise)
        ref.register.regiter_attention_editor_diffusers(pipe, ctrl)
        out[name], out[name + "_layers"] = _masa_loop(pipe, ctrl, prompts, 9, hw)
    # plain mutual control on the same inputs: shows how far the masks move the result (a test that passes with the masks
    # ignored would be worthless)
    pipe = make_pipeline(cfg, seed=4)
    ref.sd_utils.MasaCtrl(pipe, steps)
    ctrl = ref.attention_control.MutualSelfAttentionControl(1, 10, total_steps=steps)
    ref.register.regiter_attention_editor_diffusers(pipe, ctrl)
    out["mutual"], mutual_layers = _masa_loop(pipe, ctrl, prompts, 9, hw)
    import scenarios
    for name in ("mask", "mask_auto"):
        dist = [(a - b).abs().max().item() for a, b in zip(out[name + "_layers"][steps - 1], mutual_layers[steps - 1])]
        out[name + "_layer_dist_to_mutual"] = dist
        print(f"   {name} vs plain mutual control: latents {scenarios.psnr(out[name][-1], out['mutual'][-1]):.1f} dB, "
              f"max layer-output distance {max(dist):.3f}")
    _save("masactrl_masks.pt", out)


def gen_pnp():
    ref = load_reference("pnp")
    pipe = make_pipeline(tiny_config(), seed=2)
    steps, keep = 4, (0, 3)
    prompts = ["a photo of a wooden horse", "a photo of a bronze horse"]
    pipe.scheduler.set_timesteps(steps)
    ts = pipe.scheduler.timesteps
    qk_t, f_t = int(steps * 0.5), int(steps * 0.8)
    ref.register.register_attention_control_efficient(pipe, ts[:qk_t])
    ref.register.register_conv_control_efficient(pipe, ts[:f_t])
    rec = Recorder(pipe.unet, keep)
    context = _context(pipe, prompts)
    init = _latent(3, (1, 4, LAT, LAT))
    latents = torch.cat([init, init])
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(ts):
            rec.step = i
            ref.register.register_time(pipe, t.item())
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            nu, nc = noise.chunk(2)
            latents = pipe.scheduler.step(nu + 7.5 * (nc - nu), t, latents).prev_sample      # pnp/model/sd_utils.py:99-107
            per_step.append(latents.clone())
    _save("pnp.pt", dict(prompts=prompts, steps=steps, keep=keep, latent_seed=3, pipe_seed=2, guidance=7.5, pnp_attn_t=0.5, pnp_f_t=0.8,
                         layer_outputs=rec.records, latents_per_step=per_step, latent_hw=LAT))


XL_TINY = dict(sample_size=8, block_out_channels=(32, 64, 64), transformer_layers=(0, 2, 3), num_heads=(2, 2, 2), cross_attention_dim=32,
               norm_num_groups=8, use_linear_projection=True, name="tiny_xl")


def gen_pnp_xl():
    """SDXL topology (3 blocks, several transformer layers per attention, linear projections): the *_xl hook variants."""
    ref = load_reference("pnp")
    pipe = make_pipeline(UNetConfig(**XL_TINY), seed=5)
    steps, keep = 4, (0, 3)
    prompts = ["a photo of a wooden horse", "a photo of a bronze horse"]
    pipe.scheduler.set_timesteps(steps)
    ts = pipe.scheduler.timesteps
    ref.register.register_attention_control_efficient_xl(pipe, ts[:int(steps * 0.5)])
    ref.register.register_conv_control_efficient_xl(pipe, ts[:int(steps * 0.8)])
    rec = Recorder(pipe.unet, keep)
    context = _context(pipe, prompts)
    init = _latent(7, (1, 4, LAT, LAT))
    latents = torch.cat([init, init])
    per_step = []
    with torch.no_grad():
        for i, t in enumerate(ts):
            rec.step = i
            ref.register.register_time_xl(pipe, t.item())
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            nu, nc = noise.chunk(2)
            latents = pipe.scheduler.step(nu + 7.5 * (nc - nu), t, latents).prev_sample
            per_step.append(latents.clone())
    n_attn = sum(1 for m in pipe.unet.modules() if type(m).__name__ == "Attention")
    _save("pnp_xl.pt", dict(prompts=prompts, steps=steps, keep=keep, latent_seed=7, pipe_seed=5, guidance=7.5, pnp_attn_t=0.5, pnp_f_t=0.8,
                            config=XL_TINY, layer_outputs=rec.records, latents_per_step=per_step, latent_hw=LAT, n_attention=n_attn))


def gen_pix2pix_zero():
    ref = load_reference("pix2pix-zero")
    pipe = make_pipeline(tiny_config(), seed=4)
    unet, _ = ref.attention_control.prep_unet(pipe.unet)
    rec = Recorder(unet, (0,))
    prompts = ["a photo of a cat"]
    context = _context(pipe, prompts)
    x = _latent(4, (1, 4, LAT, LAT))
    with torch.no_grad():
        out = unet(torch.cat([x] * 2), torch.tensor(981), encoder_hidden_states=context).sample
    probs = {name: m.attn_probs.clone() for name, m in unet.named_modules() if type(m).__name__ == "Attention" and "attn2" in name}
    _save("pix2pix_zero.pt", dict(prompts=prompts, latent_seed=4, pipe_seed=4, t=981, unet_out=out, layer_outputs=rec.records[0], cross_probs=probs, latent_hw=LAT))


def gen_pix2pix_zero_loop():
    """The two loops of pix2pix-zero/model/sd_utils.py:86-182 (map collection, then guidance gradient -> SGD step -> recomputed
    noise -> DDIM step) restated around the REFERENCE's own MyAttnProcessor and the diffusers-style scheduler: the reference
    driver itself needs diffusers' encode_prompt / prepare_latents, which are not installable here ("recomposed")."""
    ref = load_reference("pix2pix-zero")
    pipe = make_pipeline(tiny_config(), seed=4)
    unet, _ = ref.attention_control.prep_unet(pipe.unet)
    steps, guidance, lr = 3, 7.5, 0.1
    pipe.scheduler.set_timesteps(steps)
    emb_src = _context(pipe, ["a photo of a cat"])
    emb_edit = _context(pipe, ["a photo of a dog"])
    cross = [(n, m) for n, m in unet.named_modules() if type(m).__name__ == "Attention" and "attn2" in n]
    lat0 = _latent(14, (1, 4, LAT, LAT))
    latents = lat0.clone()
    d_ref = {}
    with torch.no_grad():
        for t in pipe.scheduler.timesteps:
            noise = unet(torch.cat([latents] * 2), t, encoder_hidden_states=emb_src).sample
            d_ref[t.item()] = {n: m.attn_probs.detach().cpu() for n, m in cross}
            nu, nc = noise.chunk(2)
            latents = pipe.scheduler.step(nu + guidance * (nc - nu), t, latents)["prev_sample"]
    rec = latents.clone()
    latents = lat0.clone()
    per_step = []
    for t in pipe.scheduler.timesteps:
        x_in = torch.cat([latents] * 2).detach().clone()
        x_in.requires_grad = True
        opt = torch.optim.SGD([x_in], lr=lr)
        unet(x_in, t, encoder_hidden_states=emb_edit.detach())
        loss = 0.0
        for n, m in cross:
            loss += ((m.attn_probs - d_ref[t.item()][n].detach()) ** 2).sum((1, 2)).mean(0)
        loss.backward(retain_graph=False)
        opt.step()
        with torch.no_grad():
            noise = unet(x_in.detach(), t, encoder_hidden_states=emb_edit).sample
        latents = x_in.detach().chunk(2)[0]
        nu, nc = noise.chunk(2)
        latents = pipe.scheduler.step(nu + guidance * (nc - nu), t, latents)["prev_sample"]
        per_step.append(latents.clone())
    # without guidance (lr = 0) the edit loop is plain sampling with the edit prompt: shows how much the guidance moves the result
    latents = lat0.clone()
    with torch.no_grad():
        for t in pipe.scheduler.timesteps:
            noise = unet(torch.cat([latents] * 2), t, encoder_hidden_states=emb_edit).sample
            nu, nc = noise.chunk(2)
            latents = pipe.scheduler.step(nu + guidance * (nc - nu), t, latents)["prev_sample"]
    _save("pix2pix_zero_loop.pt", dict(recomposed=True, steps=steps, guidance=guidance, guidance_amount=lr, latent_seed=14, pipe_seed=4, latent_hw=LAT,
                                       prompts=["a photo of a cat", "a photo of a dog"], rec=rec, edit_per_step=per_step, edit_no_guidance=latents))


def gen_oracle_pins():
    """Small known-answer vectors straight from the reference's controller METHODS on random tensors: they pin every
    function of oracle/controlled_attention.py without needing /root/reference at test time."""
    p2p, masa, pnp = load_reference("p2p"), load_reference("masactrl"), load_reference("pnp")
    from image_editing_framework_b200.standin import Attention
    tok = WordPieceTokenizer()
    g = torch.Generator().manual_seed(10)
    H, N, M, d, steps = 2, 32, 77, 8, 10
    out = dict(H=H, N=N, M=M, d=d, steps=steps, seed=10)
    prompts = ["a squirrel eating a burger", "a lion eating a burger", "a hippopotamus eating a burger"]
    rprompts = ["a bowl of soup", "a bowl of pea soup", "a large bowl of hot soup"]
    cross = torch.randn(6 * H, N, M, generator=g).softmax(-1)
    selfp = torch.randn(6 * H, N, N, generator=g).softmax(-1)
    big = torch.randn(6 * H, 257, 257, generator=torch.Generator().manual_seed(12)).softmax(-1)  # > 16^2 tokens: never replaced
    out.update(cross=cross, selfp=selfp, big_seed=12, prompts=prompts, rprompts=rprompts)

    def run(ctrl, probs, is_cross, step):
        ctrl.num_att_layers, ctrl.cur_step, ctrl.cur_att_layer = 100, step, 0
        return ctrl(probs.clone(), is_cross, "down")

    kw = dict(tokenizer=tok, num_steps=steps, cross_replace_steps={"default_": 0.8, "burger": (0.0, 0.3)}, self_replace_steps=0.4, device=CPU)
    rep = p2p.attention_control.AttentionReplace(prompts=prompts, **kw)
    kw["cross_replace_steps"] = 0.8
    ref_ = p2p.attention_control.AttentionRefine(prompts=rprompts, **kw)
    eq = p2p.seq_aligner.get_equalizer(tok, prompts[1], ("lion",), (2.0, -1.0))
    rew = p2p.attention_control.AttentionReweight(prompts=prompts, equalizer=eq, controller=rep, **kw)
    rew0 = p2p.attention_control.AttentionReweight(prompts=prompts, equalizer=eq, controller=None, **kw)
    for name, c in (("replace", rep), ("refine", ref_), ("reweight_chain", rew), ("reweight", rew0)):
        out[name] = {f"cross_step{s}": run(c, cross, True, s) for s in (0, 2, 9)}
        out[name].update({f"self_step{s}": run(c, selfp, False, s) for s in (0, 5)})
        out[name]["self_big_unchanged"] = bool(torch.equal(run(c, big, False, 0), big))
    out["tables"] = dict(replace_mapper=rep.mapper, replace_alpha=rep.cross_replace_alpha, refine_mapper=ref_.mapper, refine_alphas=ref_.alphas,
                         refine_alpha=ref_.cross_replace_alpha, equalizer=eq, num_self_replace=rep.num_self_replace)
    # MasaCtrl mutual self-attention on '(b h) n d' tensors
    q, k, v = (torch.randn(4 * H, N, d, generator=g) for _ in range(3))
    ed = masa.attention_control.MutualSelfAttentionControl(0, 0, total_steps=4)
    ed.num_att_layers = 100
    sim = torch.einsum("bid,bjd->bij", q, k) * d ** -0.5
    out["masactrl"] = dict(q=q, k=k, v=v, out=ed.forward(q, k, v, sim, sim.softmax(-1), False, "up", H, scale=d ** -0.5))
    out["masactrl_plain"] = masa.attention_base.AttentionBase.forward(ed, q, k, v, sim, sim.softmax(-1), False, "up", H, scale=d ** -0.5)
    # PnP closure on a stand-in Attention
    torch.manual_seed(11)
    attn = Attention(16, None, H, d).eval()

    class Holder:
        pass
    hold = Holder()
    hold.unet = Holder()
    blk = Holder()
    blk.attn1 = attn
    att = Holder()
    att.transformer_blocks = [blk]
    up = Holder()
    up.attentions = [att, att, att]
    hold.unet.up_blocks = [None, up, up, up]
    pnp.register.register_attention_control_efficient(hold, torch.tensor([981, 961]))
    x = torch.randn(4, N, 16, generator=g)
    attn.t = 981
    with torch.no_grad():
        inj = attn.forward(x)
        attn.t = 1
        plain = attn.forward(x)
    out["pnp"] = dict(x=x, state=attn.state_dict(), injected=inj, plain=plain)
    _save("oracle_pins.pt", out)


def gen_ddim():
    ref = load_reference("p2p")
    from types import SimpleNamespace
    sch = DDIMScheduler()
    sch.set_timesteps(50)
    model = SimpleNamespace(scheduler=sch)
    inv = ref.ddim.ddim_inversion()
    g = torch.Generator().manual_seed(6)
    eps, x = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 4, 64, 64, generator=g)
    rev = {int(t): inv.ddim_reverse(model, eps, torch.tensor(int(t)), x) for t in (1, 21, 501, 981)}
    eu, ec = torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 4, 64, 64, generator=g)
    xx = torch.randn(2, 4, 64, 64, generator=g)
    fwd = {int(t): sch.step(eu + 7.5 * (ec - eu), int(t), xx)["prev_sample"] for t in (981, 501, 21, 1)}
    _save("ddim.pt", dict(seed=6, reverse=rev, forward=fwd, guidance=7.5, timesteps=sch.timesteps.clone(), alphas_cumprod=sch.alphas_cumprod.clone()))


def gen_pipelines():
    """uint8 images of the 16 pipeline-level classes (`*/model/sd_utils.py`) on the scenarios of tests/scenarios.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenarios
    out = {}
    for family, name in scenarios.PIPELINE_CASES:
        images, _ = scenarios.run_pipeline_case(family, name, load_reference(family), CPU)
        out[name] = [torch.from_numpy(im.copy()) for im in images]
    _save("pipelines.pt", out)


def repaired_union(ref):
    """The reference's MutualSelfAttentionControlUnion (masactrl/model/attention_control.py:71-107) with ONE repair. As shipped, its
    source branch calls `super().forward` — the *mutual* forward — on a single batch row; that forward splits its input into CFG
    halves again (`q.chunk(2)` now cuts the HEADS in two) and attn_batch then computes `b = (H/2) // H = 0` rows and fails. What the
    comment above those lines says ("source image branch") and what upstream MasaCtrl does is plain attention for the source rows, so
    the repaired class runs the source rows through AttentionBase.forward (out = attn v). The target branch — queries of the target
    row against the token-wise concatenation [K_src; K_tgt], [V_src; V_tgt] — is the reference's code, untouched."""
    AC = ref.attention_control
    base_forward = ref.attention_base.AttentionBase.forward

    class MutualSelfAttentionControlUnionRepaired(AC.MutualSelfAttentionControlUnion):
        def forward(self, q, k, v, sim, attn, is_cross, place_in_unet, num_heads, **kwargs):
            if is_cross or self.cur_step not in self.step_idx or self.cur_att_layer // 2 not in self.layer_idx:
                return base_forward(self, q, k, v, sim, attn, is_cross, place_in_unet, num_heads, **kwargs)
            (qu_s, qu_t, qc_s, qc_t), (ku_s, ku_t, kc_s, kc_t) = q.chunk(4), k.chunk(4)
            (vu_s, vu_t, vc_s, vc_t), (au_s, au_t, ac_s, ac_t) = v.chunk(4), attn.chunk(4)
            src_u = base_forward(self, qu_s, ku_s, vu_s, sim, au_s, is_cross, place_in_unet, num_heads, **kwargs)     # <- the repair
            src_c = base_forward(self, qc_s, kc_s, vc_s, sim, ac_s, is_cross, place_in_unet, num_heads, **kwargs)     # <- the repair
            tgt_u = self.attn_batch(qu_t, torch.cat([ku_s, ku_t]), torch.cat([vu_s, vu_t]), sim[:num_heads], au_t, is_cross, place_in_unet,
                                    num_heads, **kwargs)
            tgt_c = self.attn_batch(qc_t, torch.cat([kc_s, kc_t]), torch.cat([vc_s, vc_t]), sim[:num_heads], ac_t, is_cross, place_in_unet,
                                    num_heads, **kwargs)
            return torch.cat([src_u, tgt_u, src_c, tgt_c], dim=0)

    return MutualSelfAttentionControlUnionRepaired


def gen_fullgeo():
    """Whole edits at the BASELINE attention geometry (64x64 latents; head dims 40 / 80 / 160 and 64) through the reference's own
    register closures + controllers + P2P.diffusion_step, fp32 on the CPU (tests/scenarios.py::run_fullgeo with api = the reference's
    modules). One file per (config, scenario); per-layer outputs are kept on 16 token rows at the last step."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import time
    import scenarios
    only = os.environ.get("FULLGEO_ONLY")
    for cfg_name, kind in scenarios.FULLGEO_CASES:
        if only and only not in f"{cfg_name}:{kind}":
            continue
        t0 = time.time()
        family = {"p2p": "p2p", "masactrl": "masactrl", "pnp": "pnp"}[kind.split("_")[0]]
        api = load_reference(family)
        if kind == "masactrl_union":
            api.attention_control.MutualSelfAttentionControlUnionRepaired = repaired_union(api)
        ctrl, records, per_step = scenarios.run_fullgeo(cfg_name, kind, api, CPU, fused_step=False)
        g = dict(config=scenarios.FULLGEO_CONFIGS[cfg_name], kind=kind, steps=scenarios.FULLGEO_STEPS, rows=scenarios.FULLGEO_ROWS,
                 layer_outputs=records, latents_per_step=per_step)
        if kind == "masactrl_union":
            g["repaired"] = "source rows through AttentionBase.forward instead of the mutual forward (see make_goldens.py::repaired_union)"
        if kind == "p2p_store":
            avg = ctrl.get_average_attention()
            # the 16x16 cross maps (what LocalBlend / MaskAuto read) in full, the larger stored maps on a strip of query rows
            g["average_attention"] = {k: [t[:, ::max(1, t.shape[1] // 8)].clone() for t in v] for k, v in avg.items()}
        if ctrl is not None:
            g["cur_step"] = ctrl.cur_step
        if kind != "p2p_store":
            # how far the UNCONTROLLED run (same prompts, latents, UNet; reference closures with an EmptyControl) is from this one:
            # a test that would also pass with the edit missing is worthless, so the fixture carries the distance it must beat
            p2p_ref = load_reference("p2p")
            plain_api = SimpleNamespace(attention_base=SimpleNamespace(AttentionStore=lambda lr: p2p_ref.attention_base.EmptyControl(lr)),
                                        attention_control=None, sd_utils=p2p_ref.sd_utils, register=p2p_ref.register)
            saved = scenarios.FULLGEO_PROMPTS["p2p_store"]
            scenarios.FULLGEO_PROMPTS["p2p_store"] = scenarios.FULLGEO_PROMPTS[kind]
            try:
                _, rec0, per0 = scenarios.run_fullgeo(cfg_name, "p2p_store", plain_api, CPU, fused_step=False)
            finally:
                scenarios.FULLGEO_PROMPTS["p2p_store"] = saved
            last = scenarios.FULLGEO_STEPS - 1
            g["uncontrolled_layer_dist"] = [(a - b).abs().max().item() for a, b in zip(rec0[last], records[last])]
            g["uncontrolled_latents_psnr"] = [scenarios.psnr(a, b) for a, b in zip(per0, per_step)]
            print(f"   uncontrolled run: max layer distance {max(g['uncontrolled_layer_dist']):.3f}, latents {g['uncontrolled_latents_psnr'][-1]:.1f} dB")
        _save(f"fullgeo_{cfg_name}_{kind}.pt", g)
        print(f"   ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["aligner", "oracle_pins", "ddim", "p2p", "masactrl", "pnp", "pnp_xl", "pix2pix_zero", "p2p_localblend", "masactrl_masks", "pix2pix_zero_loop", "pipelines"]
    for w in which:
        globals()["gen_" + w]()
