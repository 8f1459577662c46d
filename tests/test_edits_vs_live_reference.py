"""Whole edits against the LIVE reference on random configurations (only where /root/reference is mounted; skipped elsewhere):
the reference's own register closures + controllers + P2P.diffusion_step on the CPU stand-in versus this repository's register /
controller / driver code on the same stand-in with oracle-backed ops (fp32). Complements the committed goldens, which pin a fixed set
of configurations, with 3-prompt batches, tuple / dict step windows and random prompts."""
import random

import pytest
import torch

from oracle import reference_loader
from oracle import cpu_ops as cpu_backend
from image_editing_framework_b200 import p2p, masactrl, editing
from image_editing_framework_b200.ddim import FusedDDIM
from image_editing_framework_b200.standin import make_pipeline, tiny_config

pytestmark = pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not mounted")
WORDS = ["a", "photo", "of", "cat", "dog", "squirrel", "hippopotamus", "burger", "eating", "sitting", "on", "the", "bench", "large", "house"]
CPU = torch.device("cpu")


def _context(pipe, prompts):
    tok = pipe.tokenizer(prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt")
    un = pipe.tokenizer([""] * len(prompts), padding="max_length", max_length=77, return_tensors="pt")
    return torch.cat([pipe.text_encoder(un.input_ids)[0], pipe.text_encoder(tok.input_ids)[0]])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_p2p_edit_matches_live_reference(monkeypatch, seed):
    ref = reference_loader.load_reference("p2p")
    rng = random.Random(seed)
    n_prompts = rng.choice([2, 3])
    src = " ".join(rng.choice(WORDS) for _ in range(rng.randint(3, 6)))
    kind = rng.choice(["replace", "refine", "reweight"])
    if kind == "refine":
        prompts = [src] + [src + " " + rng.choice(WORDS) for _ in range(n_prompts - 1)]
    else:
        prompts = [src] + [" ".join(rng.choice(WORDS) if rng.random() < 0.4 else w for w in src.split(" ")) for _ in range(n_prompts - 1)]
    steps = rng.choice([3, 4])
    cross = rng.choice([0.8, (0.0, 0.5), {"default_": 0.7, prompts[1].split(" ")[0]: (0.0, 0.3)}])
    self_ = rng.choice([0.6, (0.25, 1.0)])
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(1, 4, 8, 8, generator=g)

    def build(mod_ac, mod_aligner, tok):
        kw = dict(prompts=prompts, tokenizer=tok, num_steps=steps, cross_replace_steps=cross, self_replace_steps=self_, device=CPU)
        if kind == "replace":
            return mod_ac.AttentionReplace(**kw)
        if kind == "refine":
            return mod_ac.AttentionRefine(**kw)
        eq = mod_aligner.get_equalizer(tok, prompts[1], (prompts[1].split(" ")[-1],), (2.5,))
        eq = eq.expand(n_prompts - 1, -1).contiguous()
        return mod_ac.AttentionReweight(equalizer=eq, controller=mod_ac.AttentionReplace(**kw), **kw)

    # reference
    pipe = make_pipeline(tiny_config(), seed=seed)
    ctrl = build(ref.attention_control, ref.seq_aligner, pipe.tokenizer)
    editor = ref.sd_utils.P2P(pipe, steps)
    ref.register.register_attention_control(pipe, ctrl)
    context = _context(pipe, prompts)
    latents = lat.expand(n_prompts, 4, 8, 8)
    with torch.no_grad():
        for t in pipe.scheduler.timesteps:
            latents = editor.diffusion_step(pipe, ctrl, latents, context, t, 7.5, False)
    want = latents.clone()
    # this repository (host logic on oracle-backed ops)
    cpu_backend.install(monkeypatch)
    pipe = make_pipeline(tiny_config(), seed=seed)
    mine = build(p2p, p2p.seq_aligner, pipe.tokenizer)
    got = editing.p2p_edit(pipe, prompts, mine, lat, steps, 7.5)
    assert mine.cur_step == ctrl.cur_step and mine.num_att_layers == ctrl.num_att_layers
    assert torch.allclose(got, want, atol=3e-4), (kind, prompts, cross, self_, (got - want).abs().max().item())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_masactrl_edit_matches_live_reference(monkeypatch, seed):
    ref = reference_loader.load_reference("masactrl")
    rng = random.Random(100 + seed)
    steps = rng.choice([3, 4, 5])
    start_step, start_layer = rng.randrange(steps), rng.choice([8, 10, 12])
    layer_idx = rng.choice([None, [9, 11, 13, 15]])
    prompts = [" ".join(rng.choice(WORDS) for _ in range(4)) for _ in range(2)]
    lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(seed))
    pipe = make_pipeline(tiny_config(), seed=seed)
    ref.sd_utils.MasaCtrl(pipe, steps)
    ed = ref.attention_control.MutualSelfAttentionControl(start_step, start_layer, layer_idx=layer_idx, total_steps=steps)
    ref.register.regiter_attention_editor_diffusers(pipe, ed)
    context = _context(pipe, prompts)
    latents = torch.cat([lat, lat])
    with torch.no_grad():
        for t in pipe.scheduler.timesteps:
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            nu, nc = noise.chunk(2, dim=0)
            latents = pipe.scheduler.step(nu + 7.5 * (nc - nu), t, latents, return_dict=True)["prev_sample"]
    want = latents.clone()
    cpu_backend.install(monkeypatch)
    pipe = make_pipeline(tiny_config(), seed=seed)
    mine = masactrl.MutualSelfAttentionControl(start_step, start_layer, layer_idx=layer_idx, total_steps=steps)
    masactrl.regiter_attention_editor_diffusers(pipe, mine)
    got = editing.masactrl_edit(pipe, prompts, torch.cat([lat, lat]), steps, 7.5)
    assert mine.cur_step == ed.cur_step
    assert torch.allclose(got, want, atol=3e-4), (start_step, start_layer, layer_idx, (got - want).abs().max().item())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_pnp_edit_matches_live_reference(monkeypatch, seed):
    """Plug-and-Play with random step count and injection thresholds (incl. 0 and 1: nothing / everything injected)."""
    ref = reference_loader.load_reference("pnp")
    rng = random.Random(200 + seed)
    steps = rng.choice([3, 4, 6])
    attn_t, f_t = rng.choice([0.0, 0.34, 0.5, 1.0]), rng.choice([0.0, 0.5, 0.8, 1.0])
    prompts = [" ".join(rng.choice(WORDS) for _ in range(4)) for _ in range(2)]
    lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(seed))
    pipe = make_pipeline(tiny_config(), seed=seed)
    pipe.scheduler.set_timesteps(steps)
    ts = pipe.scheduler.timesteps
    ref.register.register_attention_control_efficient(pipe, ts[:int(steps * attn_t)])
    ref.register.register_conv_control_efficient(pipe, ts[:int(steps * f_t)])
    context = _context(pipe, prompts)
    latents = torch.cat([lat, lat])
    with torch.no_grad():
        for t in ts:
            ref.register.register_time(pipe, t.item())
            noise = pipe.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context).sample
            nu, nc = noise.chunk(2)
            latents = pipe.scheduler.step(nu + 7.5 * (nc - nu), t, latents).prev_sample
    want = latents.clone()
    cpu_backend.install(monkeypatch)
    pipe = make_pipeline(tiny_config(), seed=seed)
    got = editing.pnp_edit(pipe, prompts, torch.cat([lat, lat]), steps, 7.5, pnp_attn_t=attn_t, pnp_f_t=f_t)
    assert torch.allclose(got, want, atol=3e-4), (steps, attn_t, f_t, (got - want).abs().max().item())
