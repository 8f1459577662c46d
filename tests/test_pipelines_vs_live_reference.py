"""Pipeline-level drivers (the classes of the reference's `*/model/sd_utils.py`) against the LIVE reference on the CPU stand-in
(only where /root/reference is mounted; skipped elsewhere): the reference's class + its own register closures and scheduler versus
this repository's class of the same name on oracle-backed ops. Images must agree to one uint8 step, start latents exactly."""
import numpy as np
import pytest
import torch

from oracle import reference_loader
from oracle import cpu_ops as cpu_backend
from image_editing_framework_b200 import p2p
from image_editing_framework_b200.standin import make_pipeline, tiny_config
from image_editing_framework_b200.standin.unet import UNetConfig

pytestmark = pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not mounted")
CPU = torch.device("cpu")
PROMPTS = ["a photo of a cat sitting on the bench", "a photo of a dog sitting on the bench"]


class _XL:
    """StableDiffusionXLPipeline members the XL drivers touch, over the stand-in pipeline (added_cond_kwargs are recorded, the
    stand-in UNet has no add-embedding)."""

    def __init__(self, seed, config=None):
        self._p = make_pipeline(config or tiny_config(), seed=seed)
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


def _controller(mod, tok, steps):
    return mod.AttentionReplace(prompts=PROMPTS, tokenizer=tok, num_steps=steps, cross_replace_steps=0.8, self_replace_steps=0.5, device=CPU)


def _null_text(steps, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(1, 77, tiny_config().cross_attention_dim, generator=g) * 0.1 for _ in range(steps)]


@pytest.mark.parametrize("name", ["P2P", "P2P_NTI", "P2P_XL", "P2P_XL_NTI"])
def test_p2p_pipeline_classes_match_live_reference(monkeypatch, name):
    ref = reference_loader.load_reference("p2p")
    steps, xl, nti = 4, "XL" in name, "NTI" in name
    lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(11))
    kw = {"uncond_embeddings_list": _null_text(steps, 12)} if nti else {}

    def run(cls, ac_mod, pipe):
        ctrl = _controller(ac_mod, pipe.tokenizer, steps)
        editor = cls(pipe, steps)
        # both sides hard-code 512^2 / 1024^2: keep their loop but start from an 8x8 latent so the CPU run stays small
        editor.init_latent = lambda latent, model, h, w, gen, bs: (latent, latent.expand(bs, 4, 8, 8))
        image, x_t = editor.text2image_ldm_stable(pipe, PROMPTS, ctrl, num_inference_steps=steps, guidance_scale=7.5, latent=lat, **kw)
        return image, x_t, ctrl

    want_img, want_xt, ref_ctrl = run(getattr(ref.sd_utils, name), ref.attention_control, _XL(7) if xl else make_pipeline(tiny_config(), seed=7))
    cpu_backend.install(monkeypatch)
    mine = _XL(7) if xl else make_pipeline(tiny_config(), seed=7)
    got_img, got_xt, ctrl = run(getattr(p2p, name), p2p, mine)
    assert got_img.dtype == np.uint8 and got_img.shape == want_img.shape
    assert torch.equal(got_xt, want_xt) and ctrl.cur_step == ref_ctrl.cur_step == steps
    assert np.abs(got_img.astype(np.int16) - want_img.astype(np.int16)).max() <= 1
    if xl:
        assert len(mine.added) == steps and all(a["time_ids"].shape == (2 * len(PROMPTS), 6) for a in mine.added)


def test_p2p_pipeline_refuses_low_resource(monkeypatch):
    cpu_backend.install(monkeypatch)
    pipe = make_pipeline(tiny_config(), seed=1)
    with pytest.raises(NotImplementedError, match="low_resource"):
        p2p.P2P(pipe, 2).text2image_ldm_stable(pipe, PROMPTS, None, num_inference_steps=2, latent=torch.zeros(1, 4, 64, 64), low_resource=True)


@pytest.mark.parametrize("name", ["MasaCtrl", "MasaCtrl_NTI", "MasaCtrl_XL", "MasaCtrl_XL_NTI"])
@pytest.mark.parametrize("guidance", [7.5, 1.0])
def test_masactrl_pipeline_classes_match_live_reference(monkeypatch, name, guidance):
    from image_editing_framework_b200 import masactrl
    ref = reference_loader.load_reference("masactrl")
    steps, xl, nti = 4, "XL" in name, "NTI" in name
    if (nti or xl) and guidance <= 1.0:
        pytest.skip("without guidance the reference's NTI samplers pair a doubled context with an un-doubled batch and its XL "
                    "encode_prompt_xl raises UnboundLocalError: nothing to compare with")
    g = torch.Generator().manual_seed(21)
    lat = torch.randn(1, 4, 8, 8, generator=g)
    trajectory = [torch.randn(1, 4, 8, 8, generator=g) for _ in range(steps + 1)]      # stands for the DDIM inversion's latents
    null = _null_text(steps, 22)

    def run(cls, mod, pipe):
        editor = mod.attention_control.MutualSelfAttentionControl(1, 10, total_steps=steps) if hasattr(mod, "attention_control") else \
            mod.MutualSelfAttentionControl(1, 10, total_steps=steps)
        (mod.register if hasattr(mod, "register") and hasattr(mod.register, "regiter_attention_editor_diffusers") else mod) \
            .regiter_attention_editor_diffusers(pipe, editor)
        kw = dict(height=64, width=64, num_inference_steps=steps, guidance_scale=guidance, latents=torch.cat([lat, lat]),
                  ref_intermediate_latents=trajectory)
        if nti:
            kw["uncond_embeddings_list"] = null
        elif not xl:
            kw["unconditioning"] = null if guidance > 1.0 else None
            kw["neg_prompt"] = "blurry"
        image, x_t = cls(pipe, steps)(PROMPTS, **kw)
        return image, x_t, editor

    mk = (lambda: _XL(9)) if xl else (lambda: make_pipeline(tiny_config(), seed=9))
    want_img, want_xt, ref_ed = run(getattr(ref.sd_utils, name), ref, mk())
    cpu_backend.install(monkeypatch)
    got_img, got_xt, ed = run(getattr(masactrl, name), masactrl, mk())
    assert torch.equal(got_xt, want_xt) and ed.cur_step == ref_ed.cur_step == steps
    assert got_img.shape == want_img.shape and np.abs(got_img.astype(np.int16) - want_img.astype(np.int16)).max() <= 1


@pytest.mark.parametrize("name", ["PnP", "PnP_NTI", "PnP_XL", "PnP_XL_NTI"])
def test_pnp_pipeline_classes_match_live_reference(monkeypatch, name):
    from image_editing_framework_b200 import pnp
    ref = reference_loader.load_reference("pnp")
    steps, xl, nti = 5, "XL" in name, "NTI" in name
    lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(31))
    kw = {"uncond_embeddings_list": _null_text(steps, 32)} if nti else {}

    def run(cls, pipe):
        return cls(pipe, steps)(PROMPTS, height=64, width=64, num_inference_steps=steps, guidance_scale=7.5, latents=lat,
                                pnp_attn_t=0.5, pnp_f_t=0.8, **kw), pipe

    # the *_xl hook tables address SDXL's block topology (3 blocks, no attention in the first): a tiny UNet of that shape
    xl_cfg = UNetConfig(sample_size=8, block_out_channels=(32, 64, 64), transformer_layers=(0, 2, 3), num_heads=(2, 2, 2),
                        cross_attention_dim=32, norm_num_groups=8, use_linear_projection=True, name="tiny_xl")
    mk = (lambda: _XL(13, xl_cfg)) if xl else (lambda: make_pipeline(tiny_config(), seed=13))
    want, ref_pipe = run(getattr(ref.sd_utils, name), mk())
    cpu_backend.install(monkeypatch)
    got, pipe = run(getattr(pnp, name), mk())
    assert got.dtype == np.uint8 and got.shape == want.shape == (2, 64, 64, 3)
    assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
