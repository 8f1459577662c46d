"""End-to-end parity on the B200: the golden scenarios (outputs of the reference's own code, CPU fp32) replayed through the
registered closures + controllers + CUDA kernels. Gates from BASELINE.json north_star: per-layer attention outputs within
2e-2 max-abs in bf16 relative to the reference fp32; final latents >= 40 dB PSNR."""
import pytest
import torch

import scenarios
from scenarios import golden, psnr
from image_editing_framework_b200 import _cabi

pytestmark = pytest.mark.gpu
LAYER_TOL = 2e-2
PSNR_DB = 40.0


@pytest.fixture(autouse=True)
def _exact_fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _compare(records, per_step, g):
    worst = 0.0
    for step, outs in g["layer_outputs"].items():
        got = records[step]
        assert len(got) == len(outs)
        for i, (a, b) in enumerate(zip(got, outs)):
            err = (a - b).abs().max().item()
            worst = max(worst, err)
            assert err < LAYER_TOL, f"step {step} layer {i}: max abs err {err}"
    db = psnr(per_step[-1], g["latents_per_step"][-1])
    assert db >= PSNR_DB, f"final latents PSNR {db:.1f} dB"
    return worst, db


@pytest.mark.parametrize("kind", ["replace", "refine", "reweight", "store", "empty"])
def test_p2p_edit_matches_reference(cuda, kind):
    g = golden("p2p.pt")[kind]
    before = _cabi.launch_count()
    ctrl, records, per_step = scenarios.run_p2p(kind, g, cuda)
    assert _cabi.launch_count() - before >= g["steps"] * 33, "the attention calls did not go through libief_b200"
    assert ctrl.cur_step == g["cur_step"]
    _compare(records, per_step, g)
    if kind == "store":
        avg = ctrl.get_average_attention()
        for key, maps in g["average_attention"].items():
            for a, b in zip(avg[key], maps):
                assert (a.cpu() - b).abs().max().item() < LAYER_TOL  # probabilities in [0,1] from bf16 q,k: same 2e-2 gate


def test_masactrl_edit_matches_reference(cuda):
    g = golden("masactrl.pt")
    _, records, per_step = scenarios.run_masactrl(g, cuda)
    _compare(records, per_step, g)


@pytest.mark.parametrize("which", ["mask", "mask_auto"])
def test_masactrl_masked_variants_match_reference(cuda, which):
    """MutualSelfAttentionControlMask / MaskAuto on distinct source / target rows: every attention layer's output of the last step
    within 2e-2 of the reference's, where the same run with the masks ignored (plain mutual control) is 0.36 / 0.47 away (recorded in
    the fixture), and the final latents within the 40 dB gate."""
    g = golden("masactrl_masks.pt")
    ctrl, per_step = scenarios.run_masactrl_masks(g, cuda, which)
    want = g[which + "_layers"][g["steps"] - 1]
    assert len(ctrl.layer_rows) == len(want)
    # MaskAuto derives its spatial mask by thresholding min-max normalised cross-attention maps: a query position whose map value sits
    # near the threshold can land on the other side under bf16 (the normalisation amplifies the rounding of a flat map), and its
    # whole output row then is the background instead of the foreground result, or vice versa. At most 2 of the 16 recorded
    # positions of a layer, and at most 5 % of all recorded positions, may do so; the fixed-mask variant admits none.
    flips_allowed = 2 if which == "mask_auto" else 0
    off = total = 0
    for i, (a, b) in enumerate(zip(ctrl.layer_rows, want)):
        per_position = (a - b).abs().amax(dim=(0, 2))                 # [16 token rows]
        bad = int((per_position >= LAYER_TOL).sum())
        off, total = off + bad, total + per_position.numel()
        assert bad <= flips_allowed, f"{which} layer {i}: {bad} of {per_position.numel()} positions off, worst {per_position.max().item():.4f}"
    assert off <= 0.05 * total, f"{which}: {off} of {total} recorded positions off"
    assert max(g[which + "_layer_dist_to_mutual"]) > 10 * LAYER_TOL, "fixture: the masks must matter"
    db = psnr(per_step[-1], g[which][-1])
    assert db >= PSNR_DB, f"{which}: final latents PSNR {db:.1f} dB"


def test_pnp_edit_matches_reference(cuda):
    g = golden("pnp.pt")
    records, per_step = scenarios.run_pnp(g, cuda)
    _compare(records, per_step, g)


def test_pnp_xl_edit_matches_reference(cuda):
    g = golden("pnp_xl.pt")
    records, per_step = scenarios.run_pnp(g, cuda, xl=True)
    _compare(records, per_step, g)


def test_pix2pix_zero_processor_matches_reference(cuda):
    g = golden("pix2pix_zero.pt")
    out, records, probs, _ = scenarios.run_pix2pix_zero(g, cuda)
    for i, (a, b) in enumerate(zip(records, g["layer_outputs"])):
        assert (a - b).abs().max().item() < LAYER_TOL, f"layer {i}"
    for name, p in probs.items():
        assert (p - g["cross_probs"][name]).abs().max().item() < 5e-3, name
    assert psnr(out, g["unet_out"]) >= PSNR_DB


def test_pix2pix_zero_autograd_pass_stays_differentiable(cuda):
    from image_editing_framework_b200 import pix2pix_zero, editing
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    pipe = make_pipeline(tiny_config(), seed=4, device=cuda)
    unet, _ = pix2pix_zero.prep_unet(pipe.unet)
    ctx = editing.encode_prompts(pipe, ["a photo of a cat"])[1:]
    x = torch.randn(1, 4, 8, 8, device=cuda, requires_grad=True)
    unet(x, 981, encoder_hidden_states=ctx)
    loss = sum((m.attn_probs ** 2).sum() for n, m in unet.named_modules() if type(m).__name__ == "Attention" and "attn2" in n)
    loss.backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0


def test_p2p_localblend_recomposed_oracle(cuda):
    """AttentionReplace + AttentionStore + LocalBlend (recomposed oracle, see make_goldens.py) with a mask that covers about half of
    the latent: the per-step masks equal the reference's except for isolated pixels whose normalised map sits on the threshold, and
    on every pixel whose mask history agrees the final latents are within the 40 dB gate."""
    g = golden("p2p_localblend.pt")
    ctrl, per_step, maps = scenarios.run_p2p_localblend(g, cuda)
    for a, b in zip(maps, g["store_16"]):
        assert (a - b).abs().max().item() < LAYER_TOL * g["steps"]  # SUM over steps of maps that each carry the 2e-2 gate
    agree = torch.ones(64, 64, dtype=torch.bool)
    for step, (got, want) in enumerate(zip(ctrl.blend_masks, g["masks_per_step"])):
        want = want.float()
        same = (got == want).all(0)
        assert (~same).float().mean().item() <= 0.01, f"step {step}: {(~same).sum().item()} of 4096 mask pixels differ"
        agree &= same
    assert 0.2 < g["masks_per_step"][-1].float().mean().item() < 0.8, "fixture: the mask must matter"
    a, b = per_step[-1][:, :, agree], g["latents_per_step"][-1][:, :, agree]
    assert psnr(a, b) >= PSNR_DB
    assert psnr(per_step[-1], g["latents_per_step"][-1]) >= 30.0   # including the (<= 1 %) pixels whose mask flipped under bf16


def test_fused_ddim_inversion_loop_matches_reference_formula(cuda):
    """ddim_inversion.ddim_reverse (inversion/ddim.py:9-18) through the fused kernel, bit-exact in fp32."""
    from types import SimpleNamespace
    from image_editing_framework_b200.ddim import ddim_inversion
    from image_editing_framework_b200.standin import DDIMScheduler
    g = golden("ddim.pt")
    sch = DDIMScheduler()
    sch.set_timesteps(50)
    model = SimpleNamespace(scheduler=sch)
    gen = torch.Generator().manual_seed(g["seed"])
    eps, x = torch.randn(1, 4, 64, 64, generator=gen), torch.randn(1, 4, 64, 64, generator=gen)
    inv = ddim_inversion()
    for t, want in g["reverse"].items():
        got = inv.ddim_reverse(model, eps.to(cuda), torch.tensor(t), x.to(cuda)).cpu()
        assert torch.equal(got, want), f"t={t}: {(got - want).abs().max().item()}"
    from image_editing_framework_b200.ddim import FusedDDIM
    eu, ec = torch.randn(2, 4, 64, 64, generator=gen), torch.randn(2, 4, 64, 64, generator=gen)
    xx = torch.randn(2, 4, 64, 64, generator=gen)
    fused = FusedDDIM(sch)
    for t, want in g["forward"].items():
        got = fused.step(torch.cat([eu, ec]).to(cuda), t, xx.to(cuda), g["guidance"]).cpu()
        assert torch.equal(got, want), f"t={t}: {(got - want).abs().max().item()}"


def test_cuda_graph_replay_equals_eager(cuda):
    """The fused attention launches are captured like torch ops (tensor maps / row tables travel by value)."""
    import io
    from contextlib import redirect_stdout
    from image_editing_framework_b200 import masactrl, editing
    from image_editing_framework_b200.graphs import GraphedCall
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    from image_editing_framework_b200.standin.unet import UNetConfig
    cfg = UNetConfig(sample_size=32, block_out_channels=(64, 128, 128, 128), num_heads=(2, 2, 2, 2), cross_attention_dim=32, norm_num_groups=8, name="g")
    pipe = make_pipeline(cfg, seed=0, device=cuda, dtype=torch.bfloat16)
    ctx = editing.encode_prompts(pipe, ["a cat", "a dog"])
    with redirect_stdout(io.StringIO()):
        ed = masactrl.MutualSelfAttentionControl(0, 10, total_steps=4)
    masactrl.regiter_attention_editor_diffusers(pipe, ed)

    def fwd(x, t, c):
        ed.cur_step, ed.cur_att_layer = 1, 0
        return pipe.unet(x, t, encoder_hidden_states=c).sample

    x = torch.randn(4, 4, 32, 32, device=cuda, dtype=torch.bfloat16)   # 32x32 latents: 1024-token layers run the tcgen05 kernel
    t = torch.tensor(501, device=cuda)
    with torch.no_grad():
        want = fwd(x, t, ctx).clone()
    g = GraphedCall(fwd, [x, t, ctx], launch_counter=_cabi.launch_count)
    assert g.captured_launches == 32
    x2 = torch.randn_like(x)
    with torch.no_grad():
        want2 = fwd(x2, t, ctx).clone()
    assert torch.equal(g(x, t, ctx), want)
    assert torch.equal(g(x2, t, ctx), want2)


@pytest.mark.parametrize("method", ["p2p_replace_localblend", "p2p_refine", "masactrl", "masactrl_mask_auto", "pnp"])
def test_graph_replay_matches_eager(cuda, method):
    """editing.*(graphs=True): phase-keyed CUDA-graph replay of the UNet forwards gives the eager result, with the controller
    counters, the map store and LocalBlend behaving as in eager mode."""
    from image_editing_framework_b200 import p2p, masactrl, editing
    from image_editing_framework_b200.standin import make_pipeline
    from image_editing_framework_b200.standin.unet import UNetConfig
    lb_case = method == "p2p_replace_localblend"   # LocalBlend reads the 16x16 maps of the 64x64-latent topology
    cfg = UNetConfig(**(golden("p2p_localblend.pt") if lb_case else golden("masactrl_masks.pt"))["config"])
    hw = 64 if lb_case else 32
    prompts = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]
    steps = 10
    results, stats = [], {}
    for graphs in (False, True):
        pipe = make_pipeline(cfg, seed=7, device=cuda)
        lat1 = scenarios.latent(11, (1, 4, hw, hw), cuda)
        lat2 = torch.cat([lat1, lat1])
        common = dict(prompts=prompts, tokenizer=pipe.tokenizer, num_steps=steps, cross_replace_steps=0.8, self_replace_steps=0.4, device=cuda)
        if method.startswith("p2p"):
            if method == "p2p_refine":
                ctrl = p2p.AttentionRefine(**common)
            else:
                ctrl = p2p.AttentionReplace(local_blend=p2p.LocalBlend(pipe.tokenizer, prompts, [["cat"], ["dog"]], device=cuda), **common)
            out = editing.p2p_edit(pipe, prompts, ctrl, lat1, steps, 7.5, graphs=graphs, stats=stats)
            assert ctrl.cur_step == steps and ctrl.cur_att_layer == 0
            if method == "p2p_replace_localblend":
                results.append(ctrl.get_average_attention()["up_cross"][0].float().cpu())
        elif method.startswith("masactrl"):
            ed = (masactrl.MutualSelfAttentionControl(3, 10, total_steps=steps) if method == "masactrl" else
                  masactrl.MutualSelfAttentionControlMaskAuto(3, 10, total_steps=steps, ref_token_idx=[5], cur_token_idx=[5]))
            masactrl.regiter_attention_editor_diffusers(pipe, ed)
            out = editing.masactrl_edit(pipe, prompts, lat2, steps, 7.5, graphs=graphs, editor=ed, stats=stats)
            assert ed.cur_step == steps and ed.cur_att_layer == 0
        else:
            out = editing.pnp_edit(pipe, prompts, lat2, steps, 7.5, graphs=graphs, stats=stats)
        results.append(out.float().cpu())
    n = len(results) // 2
    for a, b in zip(results[:n], results[n:]):
        assert psnr(b, a) >= 50.0, f"graph replay deviates from eager: PSNR {psnr(b, a):.1f} dB"
    assert stats["replays"] >= steps - 6 and stats["captures"] >= 1 and stats["replayed_launches"] > 0, stats


def test_persistent_graph_runner_across_edits(cuda):
    """One GraphedUNet reused over several images with controller.reset() in between: the map-store buffers are reused in place
    (captured pointers stay valid), the third edit is pure replay, and every edit equals the eager result."""
    from image_editing_framework_b200 import p2p, editing
    from image_editing_framework_b200.graphs import GraphedUNet
    from image_editing_framework_b200.standin import make_pipeline
    from image_editing_framework_b200.standin.unet import UNetConfig
    cfg = UNetConfig(**golden("p2p_localblend.pt")["config"])
    prompts = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]
    steps = 6
    pipe = make_pipeline(cfg, seed=7, device=cuda)
    common = dict(prompts=prompts, tokenizer=pipe.tokenizer, num_steps=steps, cross_replace_steps=0.8, self_replace_steps=0.5, device=cuda)
    lats = [scenarios.latent(20 + i, (1, 4, 64, 64), cuda) for i in range(3)]
    eager = []
    for lat in lats:
        ctrl = p2p.AttentionReplace(local_blend=p2p.LocalBlend(pipe.tokenizer, prompts, [["cat"], ["dog"]], device=cuda), **common)
        eager.append(editing.p2p_edit(pipe, prompts, ctrl, lat, steps, 7.5).float().cpu())
        p2p.unregister_attention_control(pipe, ctrl)
    ctrl = p2p.AttentionReplace(local_blend=p2p.LocalBlend(pipe.tokenizer, prompts, [["cat"], ["dog"]], device=cuda), **common)
    runner = GraphedUNet(pipe.unet, ctrl)
    captures = []
    for lat, want in zip(lats, eager):
        ctrl.reset()
        out = editing.p2p_edit(pipe, prompts, ctrl, lat, steps, 7.5, graphs=runner).float().cpu()
        p2p.unregister_attention_control(pipe, ctrl)
        captures.append(runner.captures)
        assert psnr(out, want) >= 50.0, f"PSNR {psnr(out, want):.1f} dB"
        assert ctrl.cur_step == steps
    assert captures[2] == captures[1], f"third edit still captured: {captures}"
    assert runner.replays >= 2 * steps
    runner.close()


@pytest.mark.parametrize("graphs", [False, True])
def test_pix2pix_zero_loops_match_reference(cuda, graphs):
    g = golden("pix2pix_zero_loop.pt")
    before = _cabi.launch_count()
    rec, edit = scenarios.run_pix2pix_zero_loop(g, cuda, graphs=graphs)
    # (launches replayed from a graph are not counted by the library: with graphs only the eager and the captured forward are)
    assert _cabi.launch_count() - before >= (64 if graphs else g["steps"] * 32), "the no_grad passes did not go through libief_b200"
    assert psnr(rec, g["rec"]) >= PSNR_DB, f"reconstruction PSNR {psnr(rec, g['rec']):.1f} dB"
    # (on this random-init stand-in the guidance moves the latents by less than bf16 noise; that it is applied correctly is
    # checked in fp32 by tests/test_host_logic.py::test_pix2pix_zero_loops_reproduce_reference on the same golden)
    db = psnr(edit, g["edit_per_step"][-1])
    assert db >= PSNR_DB, f"edit PSNR {db:.1f} dB"


def test_fused_qkv_projection_matches_separate_linears(cuda, monkeypatch):
    """hooks.project_qkv: one GEMM against the concatenated to_q/to_k/to_v (self) or to_k/to_v (cross) weights, q/k/v as strided
    views consumed in place by the kernels; the cache follows in-place weight updates."""
    from image_editing_framework_b200 import hooks, ops
    from image_editing_framework_b200.standin import Attention
    torch.manual_seed(0)
    attn = Attention(query_dim=320, cross_attention_dim=768, heads=8, dim_head=40).to(cuda).to(torch.bfloat16)
    self_attn = Attention(query_dim=320, heads=8, dim_head=40).to(cuda).to(torch.bfloat16)
    x = torch.randn(2, 1024, 320, device=cuda, dtype=torch.bfloat16)
    ctx = torch.randn(2, 77, 768, device=cuda, dtype=torch.bfloat16)
    with torch.no_grad():
        for module, context in ((self_attn, None), (attn, ctx)):
            q, k, v = hooks.project_qkv(module, x, context)
            assert not k.is_contiguous() and k.data_ptr() != v.data_ptr()      # views of one GEMM output
            monkeypatch.setenv("IEF_FUSED_QKV", "0")
            q0, k0, v0 = hooks.project_qkv(module, x, context)
            monkeypatch.delenv("IEF_FUSED_QKV")
            for a, b in ((q, q0), (k, k0), (v, v0)):
                assert (a.float() - b.float()).abs().max().item() <= 2 ** -6 * b.float().abs().max().item()
            fn = ops.attention if context is None else ops.cross_attention_edit
            o, o0 = fn(q, k, v, 8, 40 ** -0.5), fn(q0, k0, v0, 8, 40 ** -0.5)
            assert (o.float() - o0.float()).abs().max().item() < 2e-2
        # in-place weight update invalidates the cached concatenation
        k_before = hooks.project_qkv(self_attn, x, None)[1].clone()
        self_attn.to_k.weight.mul_(2.0)
        k_after = hooks.project_qkv(self_attn, x, None)[1]
        assert (k_after.float() - 2 * k_before.float()).abs().max().item() <= 2 ** -6 * k_after.float().abs().max().item()


def test_graph_runner_falls_back_to_eager_for_unreplayable_controllers(cuda):
    """masactrl.AttentionStore keeps per-step python lists of fresh tensors: graph_key() is None and every forward runs eagerly."""
    from image_editing_framework_b200 import masactrl, editing
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    prompts = ["a photo of a sitting cat", "a photo of a running cat"]
    outs, stats = [], {}
    for graphs in (False, True):
        pipe = make_pipeline(tiny_config(), seed=3, device=cuda)
        ed = masactrl.AttentionStore(res=[16], min_step=0, max_step=10)
        masactrl.regiter_attention_editor_diffusers(pipe, ed)
        lat = scenarios.latent(5, (1, 4, 16, 16), cuda)
        outs.append(editing.masactrl_edit(pipe, prompts, torch.cat([lat, lat]), 4, 7.5, graphs=graphs, editor=ed, stats=stats).float().cpu())
        assert ed.cur_step == 4 and ed.valid_steps == 4 and len(ed.self_attns) > 0
    assert stats["replays"] == 0 and stats["eager_calls"] == 4
    assert psnr(outs[1], outs[0]) >= 60.0


@pytest.mark.parametrize("family,name", scenarios.PIPELINE_CASES)
def test_pipeline_classes_match_reference_images(cuda, family, name):
    """The pipeline-level classes with the reference's names (`*/model/sd_utils.py`: text conditioning -> controlled denoising loop
    -> VAE decode) on the GPU through the CUDA kernels, against the uint8 images the reference's classes produced (CPU fp32)."""
    want = golden("pipelines.pt")[name]
    before = _cabi.launch_count()
    got, _ = scenarios.run_pipeline_case(family, name, scenarios.mirror_api(family), cuda)
    # fewest for PnP_XL: 5 steps x (the 6 hooked self-attention layers + the step update)
    assert _cabi.launch_count() - before >= 30, "the denoising loop did not go through libief_b200"
    for a, b in zip(got, want):
        db = psnr(torch.from_numpy(a).float() / 255, b.float() / 255)
        assert db >= PSNR_DB, f"{name}: image PSNR {db:.1f} dB"


def test_null_text_inversion_runs_on_the_fused_step(cuda):
    """*/inversion/nti.py on the GPU: DDIM inversion (fused reverse step) and the null-text search, whose latent advances through the
    fused guided step; replaying the found embeddings with P2P_NTI must reconstruct the inverted latent better than the plain "" prompt."""
    from image_editing_framework_b200 import nti, p2p
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    steps = 6
    pipe = make_pipeline(tiny_config(), seed=6, device=cuda)
    pipe.scheduler.set_timesteps(steps)
    x0 = scenarios.latent(51, (1, 4, 8, 8), cuda)
    prompt = scenarios.PIPELINE_PROMPTS[:1]
    inv = nti.NTI()
    before = _cabi.launch_count()
    trajectory, context = inv.ddim_inversion_loop(pipe, x0, prompt)
    found = inv.null_optimization(pipe, trajectory, context, 10, 1e-7, 7.5)
    assert _cabi.launch_count() - before >= 2 * steps
    assert len(found) == steps and all(e.shape == (1, 77, 32) and torch.isfinite(e).all() for e in found)

    def reconstruct(**kw):
        editor = p2p.P2P_NTI(pipe, steps)
        editor.init_latent = lambda latent, model, h, w, gen, bs: (latent, latent.expand(bs, 4, 8, 8))
        editor.latent2image = lambda vae, latents: latents          # compare in latent space
        return editor.text2image_ldm_stable(pipe, prompt, None, num_inference_steps=steps, guidance_scale=7.5, latent=trajectory[-1], **kw)[0]

    err_null = (reconstruct(uncond_embeddings_list=found) - x0).pow(2).mean().item()
    err_plain = (reconstruct() - x0).pow(2).mean().item()
    assert err_null < err_plain, (err_null, err_plain)


@pytest.mark.parametrize("family", ["p2p", "masactrl", "pnp", "pix2pix-zero"])
def test_pipeline_classes_graph_replay_matches_eager(cuda, family):
    """graphs=True on the reference-named classes: same images as the eager call (to one uint8 step), forwards actually replayed, and
    a second edit on the same instance + controller reuses the captured graphs."""
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    import image_editing_framework_b200 as pkg
    steps = 6
    pipe = make_pipeline(tiny_config(), seed=3, device=cuda)
    lat = scenarios.latent(8, (1, 4, 8, 8), cuda)
    prompts = scenarios.PIPELINE_PROMPTS

    def edit(graphs, state={}):
        key = (family, graphs)
        if family == "p2p":
            ctrl, editor = state.get(key) or (pkg.p2p.AttentionReplace(prompts, pipe.tokenizer, steps, 0.8, 0.5, device=cuda), pkg.p2p.P2P(pipe, steps, graphs=graphs))
            state[key] = (ctrl, editor)
            ctrl.reset()
            editor.init_latent = lambda latent, model, h, w, gen, bs: (latent, latent.expand(bs, 4, 8, 8))
            try:
                return editor.text2image_ldm_stable(pipe, prompts, ctrl, num_inference_steps=steps, latent=lat)[0], editor
            finally:
                pkg.p2p.unregister_attention_control(pipe, ctrl)
        if family == "masactrl":
            ed, editor = state.get(key) or (pkg.masactrl.MutualSelfAttentionControl(2, 10, total_steps=steps), pkg.masactrl.MasaCtrl(pipe, steps, graphs=graphs))
            state[key] = (ed, editor)
            ed.reset()
            pkg.masactrl.regiter_attention_editor_diffusers(pipe, ed)
            try:
                return editor(prompts, height=64, width=64, num_inference_steps=steps, latents=torch.cat([lat, lat]))[0], editor
            finally:
                pkg.masactrl.unregister_attention_control(pipe, ed)
        if family == "pnp":
            editor = state.setdefault(key, pkg.pnp.PnP(pipe, steps, graphs=graphs))
            return editor(prompts, height=64, width=64, num_inference_steps=steps, latents=lat, pnp_attn_t=0.5, pnp_f_t=0.8), editor
        editor = state.setdefault(key, pkg.pix2pix_zero.P2P_Zero(pipe, steps, graphs=graphs))
        try:
            return editor(prompts, height=64, width=64, num_inference_steps=steps, latents=lat.clone())[1], editor
        finally:
            pkg.pix2pix_zero.restore_original_processors(pipe.unet, editor.original_processors)

    want, _ = edit(False)
    first, editor = edit(True)
    second, editor = edit(True)
    for got in (first, second):
        if family == "pix2pix-zero":   # its guidance pass amplifies the rounding differences between captured and eager forwards
            assert psnr(torch.from_numpy(got).float(), torch.from_numpy(want).float()) >= PSNR_DB
        else:
            assert abs(got.astype("int16") - want.astype("int16")).max() <= 1
    runner = editor._runner
    assert runner.replays > steps and runner.captures <= 4, (runner.replays, runner.captures, runner.eager_calls)


def test_ddim_inversion_graph_replay_matches_eager(cuda):
    from image_editing_framework_b200.ddim import ddim_inversion
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    pipe = make_pipeline(tiny_config(), seed=4, device=cuda)
    pipe.scheduler.set_timesteps(8)
    x0 = scenarios.latent(9, (1, 4, 8, 8), cuda)
    want, _ = ddim_inversion().ddim_inversion_loop(pipe, x0, scenarios.PIPELINE_PROMPTS[:1])
    inv = ddim_inversion()
    inv.graphs = True
    got, _ = inv.ddim_inversion_loop(pipe, x0, scenarios.PIPELINE_PROMPTS[:1])
    runner = pipe.unet._ief_inversion_runners[id(None)]
    assert runner.replays >= 6 and runner.captures == 1
    assert len(got) == len(want) == 9 and all(torch.allclose(a, b, atol=1e-4, rtol=1e-4) for a, b in zip(got, want))
    # with a do-nothing editor registered the inversion's attention runs on the fused kernels and is replayed all the same
    import image_editing_framework_b200 as pkg
    plain = pkg.masactrl.AttentionBase()
    pkg.masactrl.regiter_attention_editor_diffusers(pipe, plain)
    try:
        hooked, _ = inv.ddim_inversion_loop(pipe, x0, scenarios.PIPELINE_PROMPTS[:1])
    finally:
        pkg.masactrl.unregister_attention_control(pipe, plain)
    runner = pipe.unet._ief_inversion_runners[id(plain)]
    assert runner.replays >= 6 and runner.captures == 1 and plain.cur_step == 8
    assert psnr(hooked[-1].float(), want[-1].float()) >= PSNR_DB


def test_xl_pipeline_class_graph_replay_carries_added_cond_kwargs(cuda):
    """graphs=True on an SDXL class: the pooled embedding / size ids travel as copied graph inputs, results match the eager call."""
    import image_editing_framework_b200 as pkg
    from image_editing_framework_b200.standin import tiny_config
    steps = 6
    pipe = scenarios.XLPipelineDouble(5, tiny_config(), cuda)
    lat = scenarios.latent(10, (1, 4, 8, 8), cuda)
    ed = pkg.masactrl.MutualSelfAttentionControl(2, 10, total_steps=steps)

    def edit(editor):
        ed.reset()
        pkg.masactrl.regiter_attention_editor_diffusers(pipe, ed)
        try:
            return editor(scenarios.PIPELINE_PROMPTS, height=64, width=64, num_inference_steps=steps, latents=torch.cat([lat, lat]))[0]
        finally:
            pkg.masactrl.unregister_attention_control(pipe, ed)

    want = edit(pkg.masactrl.MasaCtrl_XL(pipe, steps))
    graphed = pkg.masactrl.MasaCtrl_XL(pipe, steps, graphs=True)
    for _ in range(2):
        got = edit(graphed)
        assert abs(got.astype("int16") - want.astype("int16")).max() <= 1
    assert graphed._runner.replays > steps


# ------------------------------------------------------------------------------------------------ BASELINE attention geometry
class _ImplLog:
    """Records which kernel family served every ops.attention / ops.cross_attention_edit call (token count, head_dim, impl name)."""

    def __init__(self):
        from image_editing_framework_b200 import ops
        self.ops, self.calls = ops, []
        self._attn, self._cross = ops.attention, ops.cross_attention_edit

    def __enter__(self):
        def attention(q, k, v, heads, scale, **kw):
            out = self._attn(q, k, v, heads, scale, **kw)
            self.calls.append(("self", q.shape[1], q.shape[2] // heads if q.dim() == 3 else q.shape[3], _cabi.last_attn_impl()))
            return out

        def cross(q, k, v, heads, scale, **kw):
            out = self._cross(q, k, v, heads, scale, **kw)
            self.calls.append(("cross", q.shape[1], q.shape[2] // heads if q.dim() == 3 else q.shape[3], _cabi.last_cross_impl()))
            return out
        self.ops.attention, self.ops.cross_attention_edit = attention, cross
        return self

    def __exit__(self, *a):
        self.ops.attention, self.ops.cross_attention_edit = self._attn, self._cross


@pytest.mark.parametrize("cfg_name,kind", scenarios.FULLGEO_CASES)
def test_edits_at_baseline_attention_geometry_match_reference(cuda, cfg_name, kind):
    """64x64 latents with the real head dims (SD-1.5's 40 / 80 / 160, and 64): whole edits through this package's register closures,
    controllers and kernels against runs of the REFERENCE's own closures + controllers + diffusion_step on the same stand-in, fp32 on
    the CPU (tests/golden/fullgeo_*.pt, make_goldens.py::gen_fullgeo). Gates: every attention layer output of the last step within
    2e-2 max-abs (16 token rows per layer), final latents >= 40 dB, and the large layers served by the tcgen05 kernels."""
    import image_editing_framework_b200 as pkg
    g = golden(f"fullgeo_{cfg_name}_{kind}.pt")
    api = {"p2p": pkg.p2p, "masactrl": pkg.masactrl, "pnp": pkg.pnp}[kind.split("_")[0]]
    with _ImplLog() as log:
        ctrl, records, per_step = scenarios.run_fullgeo(cfg_name, kind, api, cuda)
    worst = 0.0
    for step, outs in g["layer_outputs"].items():
        got = records[step]
        assert len(got) == len(outs) == 32
        for i, (a, b) in enumerate(zip(got, outs)):
            err = (a - b).abs().max().item()
            worst = max(worst, err)
            assert err < LAYER_TOL, f"step {step} layer {i} ({tuple(b.shape)}): max abs err {err}"
    for i, (a, b) in enumerate(zip(per_step, g["latents_per_step"])):
        db = psnr(a, b)
        assert db >= PSNR_DB, f"latents after step {i}: PSNR {db:.1f} dB"
    if "cur_step" in g:
        assert ctrl.cur_step == g["cur_step"]
    if "uncontrolled_layer_dist" in g:
        # the gates discriminate: the same run WITHOUT the control is far outside them (fixture property), and we are far closer to
        # the controlled reference than the uncontrolled run is
        assert max(g["uncontrolled_layer_dist"]) > 5 * LAYER_TOL
        if g["uncontrolled_latents_psnr"][-1] < 60.0:    # (above that the bf16 noise floor of ~73 dB leaves no room for a margin:
            assert psnr(per_step[-1], g["latents_per_step"][-1]) >= g["uncontrolled_latents_psnr"][-1] + 6.0   # the per-layer gate decides)
    # the kernels that matter at this geometry really ran: tcgen05 for every self-attention layer with >= 1024 tokens (with the
    # probability sweep behind it where maps are stored), and for the plain cross-attention rows of the large layers
    big_self = [c for c in log.calls if c[0] == "self" and c[1] >= 1024]
    assert big_self and all(c[3] in ("tcgen05", "tcgen05+probs") for c in big_self), sorted(set(big_self))
    if kind.startswith("p2p"):   # edited / stored cross-attention rows of the large layers: the tensor-pipe edit kernel
        big_cross = [c for c in log.calls if c[0] == "cross" and c[1] >= 1024]
        want = {"p2p_replace": "tcgen05-edit", "p2p_refine": "tcgen05-edit", "p2p_store": "tcgen05-edit"}[kind]
        assert big_cross and any(c[3] == want for c in big_cross) and all(c[3] in ("tcgen05", "tcgen05-edit") for c in big_cross), sorted(set(big_cross))
    if kind == "masactrl":       # (PnP hooks eight self-attention layers only: its cross-attention stays the UNet's own)
        big_cross = [c for c in log.calls if c[0] == "cross" and c[1] >= 1024]
        assert big_cross and all(c[3] == "tcgen05" for c in big_cross), sorted(set(big_cross))
    if kind == "p2p_store":
        avg = ctrl.get_average_attention()
        for key, maps in g["average_attention"].items():
            assert len(avg[key]) == len(maps)
            for a, b in zip(avg[key], maps):
                a = a.float().cpu()
                assert (a[:, ::max(1, a.shape[1] // 8)] - b).abs().max().item() < LAYER_TOL, key
    print(f"{cfg_name}/{kind}: worst layer err {worst:.4f}, final PSNR {psnr(per_step[-1], g['latents_per_step'][-1]):.1f} dB")


# ------------------------------------------------------------------------------------------------ PIE-Bench-shaped sweep (configs[4])
def _load_sweep():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ief_sweep", os.path.join(root, "tools", "sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_sharded_sweep_reproduces_the_serial_sweep_bit_for_bit(cuda):
    """tools/sweep.py: a 6-image subset x 4 methods, once as one rank and once as the two shards of a 2-rank sweep (image i -> rank
    i mod 2, separate pipeline replicas, kept controllers re-pointed with retarget(), CUDA-graph replay on): the uint8 result
    images must agree bit for bit (CRC per image and method)."""
    sweep = _load_sweep()
    methods, n = list(sweep.METHODS), 6
    make = lambda: sweep.Worker(cuda, 6, True, "tiny", torch.float32, deterministic=True)   # (cuDNN backward-data: Pix2Pix-zero's guidance pass)
    serial = sweep.sweep(make(), n, methods, rank=0, world=1)
    sharded = {}
    for rank in range(2):
        sharded.update(sweep.sweep(make(), n, methods, rank=rank, world=2))
    assert sorted(sharded) == list(range(n))
    for i in range(n):
        assert sharded[i]["crc"] == serial[i]["crc"], (i, sharded[i]["crc"], serial[i]["crc"])
    eager = sweep.sweep(sweep.Worker(cuda, 6, False, "tiny", torch.float32, deterministic=True), 3, methods, rank=0, world=1)
    torch.backends.cudnn.deterministic = False
    for i in range(3):     # replayed graphs launch the same kernels as the eager loop
        assert {m: eager[i]["crc"][m] for m in ("p2p", "masactrl", "pnp")} == {m: serial[i]["crc"][m] for m in ("p2p", "masactrl", "pnp")}, i


def test_two_gpu_sweep_equals_one_gpu_sweep(cuda):
    """The same through torchrun on two real GPUs (skipped on a one-GPU box): crc_of_crcs of the gathered records."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [os.path.join(root, "tools", "sweep.py"), "--images", "6", "--ddim-steps", "6", "--config", "tiny", "--deterministic"]
    one = subprocess.run([sys.executable] + base, capture_output=True, text=True, timeout=600, check=True)
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29631"] + base, capture_output=True, text=True, timeout=600, check=True)
    a = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    b = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    assert a["n_gpus"] == 1 and b["n_gpus"] == 2 and a["images"] == b["images"] == 6
    assert a["crc_of_crcs"] == b["crc_of_crcs"]


def test_custom_reference_style_controller_on_the_gpu(cuda):
    """A controller that overrides replace_cross_attention (reference interface, materialised probabilities) registered on the GPU:
    the kernel emits the maps, the user's torch code edits them, P'V follows — against the same code driven by the CPU oracle ops."""
    import image_editing_framework_b200 as pkg
    from image_editing_framework_b200 import p2p, editing
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    from oracle import cpu_ops

    class HalfReplace(p2p.AttentionReplace):
        def replace_cross_attention(self, attn_base, att_replace):
            return 0.5 * super().replace_cross_attention(attn_base, att_replace) + 0.5 * att_replace

    prompts = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]
    x = torch.randn(4, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    outs = {}
    for dev in (cuda, torch.device("cpu")):
        pipe = make_pipeline(tiny_config(), seed=6, device=dev)
        ctrl = HalfReplace(prompts, pipe.tokenizer, 4, 0.8, 0.6, device=dev)
        assert ctrl._needs_probabilities()
        ctx = editing.encode_prompts(pipe, prompts)
        p2p.register_attention_control(pipe, ctrl)
        before = _cabi.launch_count()
        with torch.no_grad():
            if dev.type == "cpu":
                with cpu_ops.patched():
                    outs[dev.type] = pipe.unet(x.to(dev), 981, encoder_hidden_states=ctx).sample
            else:
                outs[dev.type] = pipe.unet(x.to(dev), 981, encoder_hidden_states=ctx).sample.float().cpu()
                assert _cabi.launch_count() - before >= 32
        p2p.unregister_attention_control(pipe, ctrl)
    assert psnr(outs["cuda"], outs["cpu"]) >= PSNR_DB
