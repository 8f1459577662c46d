"""CPU suite: goldens produced by the reference's own code pin (a) the bit-exact host-side integer logic, (b) the oracle,
(c) the controller / register / driver host logic (driven here through oracle-backed fake ops, see oracle/cpu_ops.py)."""
import os
import re

import numpy as np
import pytest
import torch

from image_editing_framework_b200 import p2p, _cabi, editing
from image_editing_framework_b200.p2p import seq_aligner, ptp_utils
from image_editing_framework_b200.standin import WordPieceTokenizer, make_pipeline, tiny_config, sd15_config, sd21_config, sdxl_config, attention_geometry
from oracle import controlled_attention as orc
from oracle import reference_loader

from oracle import cpu_ops as cpu_backend
import scenarios
from scenarios import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ief_b200.h")).read()
    declared = set(re.findall(r"\b(ief_[a-z0-9_]+)\s*\(", header))
    declared -= {"ief_tensor4", "ief_attn_params"}
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = _cabi.lib()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.ief_abi_version() == _cabi.IEF_ABI_VERSION


def test_struct_sizes_match_header_layout():
    import ctypes as C
    # 4 tensor4 (32 B each) + 6 int32 + float + int32 (=160) + 5 pointers + ptr + int32(+pad) + 2 pointers
    assert C.sizeof(_cabi.Tensor4) == 32
    assert C.sizeof(_cabi.AttnParams) == 128 + 32 + 5 * 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8
    assert C.sizeof(_cabi.CrossParams) == 128 + 32 + 8 + 8 + 8 + 7 * 8 + 8 + 8 + 8


def test_integration_stub_matches_the_binding():
    """The ctypes stub INTEGRATION.md shows a maintainer must lay its structs out exactly as the tested binding does."""
    import ctypes as C
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = re.search(r"```python\n(# p2p/model/_ief\.py.*?)```", doc, re.S).group(1)
    stub = stub.replace('C.CDLL("libief_b200.so")', f'C.CDLL({_cabi.LIB_PATH!r})')
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)            # loads the library and checks the ABI number it quotes
    for name, ours in (("Tensor4", _cabi.Tensor4), ("AttnParams", _cabi.AttnParams)):
        theirs = ns[name]
        assert C.sizeof(theirs) == C.sizeof(ours), name
        assert [(f[0], getattr(theirs, f[0]).offset) for f in theirs._fields_] == \
               [(f[0], getattr(ours, f[0]).offset) for f in ours._fields_], name


def test_ops_refuse_cpu_tensors():
    from image_editing_framework_b200 import ops
    q = torch.zeros(1, 8, 16, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.attention(q, q, q, 2, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.cfg_ddim_step(torch.zeros(4), None, torch.zeros(4), 0.0, 0.5, 0.6)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "image_editing_framework_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src


# ------------------------------------------------------------------------------------------------ token aligner (bit-exact)
@pytest.fixture(scope="module")
def aligner_golden():
    return golden("aligner.pt")


def test_tokenizer_is_the_one_the_goldens_used(aligner_golden):
    tok = WordPieceTokenizer()
    for c in aligner_golden["cases"]:
        assert tok.encode(c["source"]) == c["src_ids"] and tok.encode(c["target"]) == c["tgt_ids"]
    ids = tok.encode("a hippopotamus")
    assert len(ids) == 2 + 1 + 3 and "".join(tok.decode([i]) for i in ids[2:-1]) == "hippopotamus"


def test_mappers_bit_exact(aligner_golden):
    tok = WordPieceTokenizer()
    for c in aligner_golden["cases"]:
        pr = [c["source"], c["target"]]
        m, a = seq_aligner.get_refinement_mapper(pr, tok)
        assert m.dtype == torch.int64 and torch.equal(m, c["refinement_mapper"]), c["target"]
        assert torch.equal(a, c["refinement_alphas"])
        if c["kind"] == "replace":
            r = seq_aligner.get_replacement_mapper(pr, tok)
            assert r.dtype == torch.float32 and torch.equal(r, c["replacement_mapper"]), c["target"]
        for w, inds in c["word_inds"].items():
            assert seq_aligner.get_word_inds(c["target"], w, tok).tolist() == inds
        for i, inds in enumerate(c["word_inds_by_pos"]):
            assert seq_aligner.get_word_inds(c["target"], i, tok).tolist() == inds
        words = c["target"].split(" ")
        assert torch.equal(seq_aligner.get_equalizer(tok, c["target"], (words[-1],), (2.0,)), c["equalizer"])
        assert torch.equal(ptp_utils.get_time_words_attention_alpha(pr, 10, 0.8, tok), c["alpha_float"])
        assert torch.equal(ptp_utils.get_time_words_attention_alpha(pr, 10, {"default_": 1.0, words[-1]: (0.2, 0.6)}, tok), c["alpha_dict"])
    mg = aligner_golden["multi"]
    assert torch.equal(seq_aligner.get_replacement_mapper(mg["prompts"], tok), mg["replacement"])
    m, a = seq_aligner.get_refinement_mapper(mg["prompts"], tok)
    assert torch.equal(m, mg["refinement"][0]) and torch.equal(a, mg["refinement"][1])
    assert torch.equal(ptp_utils.get_time_words_attention_alpha(mg["prompts"], 50, (0.1, 0.7), tok), mg["alpha"])


def test_refinement_mapper_has_gaps_and_minus_one(aligner_golden):
    c = [c for c in aligner_golden["cases"] if c["target"].startswith("a watercolor")][0]
    assert (c["refinement_mapper"] == -1).any() and (c["refinement_alphas"] == 0).any()


def test_replacement_rejects_different_word_counts():
    with pytest.raises(ValueError):
        seq_aligner.get_replacement_mapper(["a cat", "a big cat"], WordPieceTokenizer())


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree only exists in the build container")
def test_aligner_matches_live_reference_on_random_prompts():
    ref = reference_loader.load_reference("p2p")
    tok = WordPieceTokenizer()
    rng = np.random.default_rng(0)
    vocab = ["a", "the", "cat", "dog", "hippopotamus", "on", "under", "table", "watercolor", "painting", "of", "blue", "squirrel", "eats"]
    for _ in range(40):
        n = int(rng.integers(1, 9))
        x = " ".join(rng.choice(vocab, n))
        y_same = x.split(" ")
        for j in rng.choice(n, size=min(2, n), replace=False):
            y_same[j] = str(rng.choice(vocab))
        y_same = " ".join(y_same)
        y_any = " ".join(rng.choice(vocab, int(rng.integers(1, 12))))
        assert torch.equal(seq_aligner.get_replacement_mapper([x, y_same], tok), ref.seq_aligner.get_replacement_mapper([x, y_same], tok))
        a, b = seq_aligner.get_refinement_mapper([x, y_any], tok), ref.seq_aligner.get_refinement_mapper([x, y_any], tok)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


# ------------------------------------------------------------------------------------------------ oracle pinned by reference outputs
@pytest.fixture(scope="module")
def pins():
    return golden("oracle_pins.pt")


@pytest.mark.parametrize("name,mode", [("replace", "replace"), ("refine", "refine"), ("reweight_chain", "replace"), ("reweight", "none")])
def test_oracle_p2p_edit_matches_reference(pins, name, mode):
    t, H = pins["tables"], pins["H"]
    kw = dict(mode=mode, num_self_replace=t["num_self_replace"])
    if name in ("replace", "reweight_chain"):
        kw.update(mapper=t["replace_mapper"], alpha_table=t["replace_alpha"])
    elif name == "refine":
        kw.update(mapper=t["refine_mapper"], refine_alphas=t["refine_alphas"], alpha_table=t["refine_alpha"])
    else:
        kw.update(alpha_table=t["replace_alpha"])
    if name.startswith("reweight"):
        kw["equalizer"] = t["equalizer"]
    for key, want in pins[name].items():
        if key == "self_big_unchanged":
            big = torch.randn(6 * H, 257, 257, generator=torch.Generator().manual_seed(pins["big_seed"])).softmax(-1)
            assert want and torch.equal(orc.p2p_edit_probs(big, H, 3, False, 0, **kw), big)
            continue
        is_cross = key.startswith("cross")
        step = int(key.split("step")[1])
        got = orc.p2p_edit_probs(pins["cross"] if is_cross else pins["selfp"], H, 3, is_cross, step, **kw)
        assert torch.allclose(got, want, atol=1e-6, rtol=1e-5), f"{name}/{key}: {(got - want).abs().max().item()}"


def test_oracle_masactrl_and_indexed_form_match_reference(pins):
    m, H, d = pins["masactrl"], pins["H"], pins["d"]
    q, k, v = (orc.batch_to_head(t, H) for t in (m["q"], m["k"], m["v"]))   # '(b h) n d' -> [B, N, H*d]
    got = orc.masactrl_mutual(q, k, v, H, d ** -0.5)
    assert torch.allclose(got, m["out"], atol=1e-6, rtol=1e-5)
    idx = orc.indexed_attention(q, k, v, H, d ** -0.5, k_src=[0, 0, 2, 2], v_src=[0, 0, 2, 2])
    assert torch.allclose(idx, m["out"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(orc.plain_attention(q, k, v, H, d ** -0.5), pins["masactrl_plain"], atol=1e-6, rtol=1e-5)


def test_oracle_pnp_injection_matches_reference(pins):
    from image_editing_framework_b200.standin import Attention
    p, H, d = pins["pnp"], pins["H"], pins["d"]
    attn = Attention(16, None, H, d)
    attn.load_state_dict(p["state"])
    x = p["x"]
    with torch.no_grad():
        q, k, v = attn.to_q(x), attn.to_k(x), attn.to_v(x)
        qi, ki = orc.pnp_inject_qk(q, k)
        inj = attn.to_out[0](orc.plain_attention(qi, ki, v, H, attn.scale))
        plain = attn.to_out[0](orc.plain_attention(q, k, v, H, attn.scale))
        idx = attn.to_out[0](orc.indexed_attention(q, k, v, H, attn.scale, q_src=[0, 2, 2, 2], k_src=[0, 2, 2, 2]))
    assert torch.allclose(inj, p["injected"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(plain, p["plain"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(idx, p["injected"], atol=1e-6, rtol=1e-5)


def test_oracle_ddim_matches_reference():
    g = golden("ddim.pt")
    gen = torch.Generator().manual_seed(g["seed"])
    eps, x = torch.randn(1, 4, 64, 64, generator=gen), torch.randn(1, 4, 64, 64, generator=gen)
    ac = g["alphas_cumprod"]
    for t, want in g["reverse"].items():
        cur = min(999, t - 20)
        a_cur = ac[cur] if cur >= 0 else ac[0]
        assert torch.equal(orc.ddim_step(eps, x, a_cur, ac[t]), want)
    eu, ec = torch.randn(2, 4, 64, 64, generator=gen), torch.randn(2, 4, 64, 64, generator=gen)
    xx = torch.randn(2, 4, 64, 64, generator=gen)
    for t, want in g["forward"].items():
        prev = t - 20
        a_prev = ac[prev] if prev >= 0 else ac[0]
        assert torch.equal(orc.ddim_step(orc.cfg_combine(eu, ec, g["guidance"]), xx, ac[t], a_prev), want)
    assert g["timesteps"].tolist() == list(range(981, 0, -20))


def test_oracle_local_blend_matches_reference():
    g = golden("local_blend.pt")
    gen = torch.Generator().manual_seed(g["seed"])
    syn = [torch.rand(16, 256, 77, generator=gen) ** 4 for _ in range(5)]
    x = torch.randn(2, 4, 64, 64, generator=gen)
    got = orc.local_blend(x, syn, g["alpha_layers"], 0.3)
    assert torch.equal(got, g["x_out"])


# ------------------------------------------------------------------------------------------------ stand-in geometry
def test_standin_attention_geometry_matches_survey():
    sd = attention_geometry(sd15_config())
    assert len(sd) == 16 and [(n, d) for _, n, _, d in sd] == [(4096, 40)] * 2 + [(1024, 80)] * 2 + [(256, 160)] * 2 + [(64, 160)] + \
        [(256, 160)] * 3 + [(1024, 80)] * 3 + [(4096, 40)] * 3
    sd21 = attention_geometry(sd21_config())
    assert len(sd21) == 16 and {d for *_, d in sd21} == {64} and sd21[0][1] == 9216
    xl = attention_geometry(sdxl_config())
    assert len(xl) == 70 and {d for *_, d in xl} == {64} and sum(1 for _, n, _, _ in xl if n == 4096) == 10


def test_tiny_unet_has_32_attention_layers_in_reference_discovery_order():
    pipe = make_pipeline(tiny_config(), seed=0)
    ctrl = p2p.EmptyControl(False)
    p2p.register_attention_control(pipe, ctrl)
    assert ctrl.num_att_layers == 32
    p2p.unregister_attention_control(pipe, ctrl)
    assert ctrl.num_att_layers == 0 and all(not hasattr(m, "_ief") for m in pipe.unet.modules())


# ------------------------------------------------------------------------------------------------ host logic end to end (fake ops)
LAYER_TOL = 2e-4  # fp32 everywhere: only reduction-order noise


def _compare(records, per_step, g, tol=LAYER_TOL):
    for step, outs in g["layer_outputs"].items():
        got = records[step]
        assert len(got) == len(outs)
        for i, (a, b) in enumerate(zip(got, outs)):
            assert a.shape == b.shape and (a - b).abs().max().item() < tol, f"step {step} layer {i}: {(a - b).abs().max().item()}"
    for i, (a, b) in enumerate(zip(per_step, g["latents_per_step"])):
        assert (a - b).abs().max().item() < tol * 5, f"latents after step {i}: {(a - b).abs().max().item()}"


@pytest.mark.parametrize("kind", ["replace", "refine", "reweight", "store", "empty"])
def test_p2p_host_logic_reproduces_reference(monkeypatch, kind):
    cpu_backend.install(monkeypatch)
    g = golden("p2p.pt")[kind]
    ctrl, records, per_step = scenarios.run_p2p(kind, g, torch.device("cpu"))
    assert ctrl.num_att_layers == g["num_att_layers"] and ctrl.cur_step == g["cur_step"] and ctrl.cur_att_layer == 0
    _compare(records, per_step, g)
    if kind == "store":
        avg = ctrl.get_average_attention()
        for key, maps in g["average_attention"].items():
            assert len(avg[key]) == len(maps)
            for a, b in zip(avg[key], maps):
                assert a.shape == b.shape and torch.allclose(a, b, atol=1e-5)


def test_masactrl_host_logic_reproduces_reference(monkeypatch):
    cpu_backend.install(monkeypatch)
    g = golden("masactrl.pt")
    ctrl, records, per_step = scenarios.run_masactrl(g, torch.device("cpu"))
    assert ctrl.num_att_layers == g["num_att_layers"] == 32
    _compare(records, per_step, g)


@pytest.mark.parametrize("which", ["mask", "mask_auto"])
def test_masactrl_masked_variants_reproduce_reference(monkeypatch, which):
    """MutualSelfAttentionControlMask / MaskAuto: key-bias tables, fg/bg passes, spatial blend, 16x16 cross-map aggregation."""
    cpu_backend.install(monkeypatch)
    g = golden("masactrl_masks.pt")
    ctrl, per_step = scenarios.run_masactrl_masks(g, torch.device("cpu"), which)
    for a, b in zip(per_step, g[which]):
        assert torch.allclose(a, b, atol=2e-4), (a - b).abs().max()
    # the masks matter: plain mutual control on the same inputs ends somewhere else
    assert (g["mutual"][-1] - g[which][-1]).abs().max() > 50 * (per_step[-1] - g[which][-1]).abs().max()
    assert ctrl.cur_step == g["steps"]
    for a, b in zip(ctrl.layer_rows, g[which + "_layers"][g["steps"] - 1]):
        assert torch.allclose(a, b, atol=2e-4), (a - b).abs().max()


def test_pnp_host_logic_reproduces_reference(monkeypatch):
    cpu_backend.install(monkeypatch)
    g = golden("pnp.pt")
    records, per_step = scenarios.run_pnp(g, torch.device("cpu"))
    _compare(records, per_step, g)


def test_pnp_xl_host_logic_reproduces_reference(monkeypatch):
    """SDXL topology: 3 blocks, 2-3 transformer layers per attention module, linear projections; the *_xl hook tables."""
    cpu_backend.install(monkeypatch)
    g = golden("pnp_xl.pt")
    records, per_step = scenarios.run_pnp(g, torch.device("cpu"), xl=True)
    assert g["n_attention"] == 2 * (2 * 2 + 2 * 3 + 3 + 3 * 3 + 3 * 2)
    _compare(records, per_step, g)
    # the un-hook functions restore the original forwards
    from image_editing_framework_b200 import pnp
    from image_editing_framework_b200.standin.unet import UNetConfig
    pipe = make_pipeline(UNetConfig(**g["config"]), seed=0)
    before = [m.forward for m in pipe.unet.modules() if type(m).__name__ == "Attention"]
    pnp.register_attention_control_efficient_xl(pipe, [1])
    pnp.register_conv_control_efficient_xl(pipe, [1])
    pnp.unregister_attention_control_efficient_xl(pipe)
    pnp.unregister_conv_control_efficient_xl(pipe)
    after = [m.forward for m in pipe.unet.modules() if type(m).__name__ == "Attention"]
    assert all(a == b for a, b in zip(before, after))


def test_pix2pix_zero_processor_reproduces_reference(monkeypatch):
    cpu_backend.install(monkeypatch)
    g = golden("pix2pix_zero.pt")
    out, records, probs, (unet, originals) = scenarios.run_pix2pix_zero(g, torch.device("cpu"))
    assert (out - g["unet_out"]).abs().max().item() < LAYER_TOL
    for a, b in zip(records, g["layer_outputs"]):
        assert (a - b).abs().max().item() < LAYER_TOL
    assert set(probs) == set(g["cross_probs"]) and len(probs) == 16
    for name, p in probs.items():
        assert torch.allclose(p, g["cross_probs"][name], atol=1e-5)
    from image_editing_framework_b200 import pix2pix_zero
    pix2pix_zero.restore_original_processors(unet, originals)
    assert all(type(m.get_processor()).__name__ == "AttnProcessor" for m in unet.modules() if type(m).__name__ == "Attention")


def test_p2p_localblend_recomposed_oracle(monkeypatch):
    """RECOMPOSED oracle (the reference's LocalBlend path crashes as shipped)."""
    cpu_backend.install(monkeypatch)
    g = golden("p2p_localblend.pt")
    ctrl, per_step, maps = scenarios.run_p2p_localblend(g, torch.device("cpu"))
    for a, b in zip(maps, g["store_16"]):
        assert torch.allclose(a, b, atol=1e-5)
    for i, (a, b) in enumerate(zip(per_step, g["latents_per_step"])):
        assert (a - b).abs().max().item() < 1e-3, f"step {i}"


@pytest.mark.parametrize("name", ["replace", "refine", "reweight_chain", "reweight"])
def test_controller_call_on_materialised_probabilities_matches_reference(pins, name):
    """controller(attn, is_cross, place) — the reference's own entry point (p2p/model/attention_base.py:16-28) — served by the
    mirrored classes on a materialised probability tensor: outputs of the reference's controllers on the same tensors are the pins."""
    tok = WordPieceTokenizer()
    steps = pins["steps"]
    kw = dict(tokenizer=tok, num_steps=steps, cross_replace_steps={"default_": 0.8, "burger": (0.0, 0.3)}, self_replace_steps=0.4, device="cpu")
    rep = p2p.AttentionReplace(prompts=pins["prompts"], **kw)
    kw["cross_replace_steps"] = 0.8
    eq = p2p.seq_aligner.get_equalizer(tok, pins["prompts"][1], ("lion",), (2.0, -1.0))
    ctrl = {"replace": lambda: rep, "refine": lambda: p2p.AttentionRefine(prompts=pins["rprompts"], **kw),
            "reweight_chain": lambda: p2p.AttentionReweight(prompts=pins["prompts"], equalizer=eq, controller=rep, **kw),
            "reweight": lambda: p2p.AttentionReweight(prompts=pins["prompts"], equalizer=eq, controller=None, **kw)}[name]()
    assert not ctrl._needs_probabilities()          # the stock classes stay on the fused path when registered
    for key, want in pins[name].items():
        if key == "self_big_unchanged":
            continue
        is_cross, step = key.startswith("cross"), int(key.split("step")[1])
        ctrl.num_att_layers, ctrl.cur_step, ctrl.cur_att_layer = 100, step, 0
        got = ctrl((pins["cross"] if is_cross else pins["selfp"]).clone(), is_cross, "down")
        assert torch.allclose(got, want, atol=1e-6, rtol=1e-5), f"{name}/{key}: {(got - want).abs().max().item()}"
        assert ctrl.cur_att_layer == 1


def test_custom_controller_written_against_the_reference_runs_through_the_registered_closure(monkeypatch):
    """A user subclass that only knows the reference interface — forward(attn, is_cross, place) on probabilities, or an overridden
    replace_cross_attention — takes the compatibility route: the kernel emits the map, the user's code edits it, P'V follows."""
    cpu_backend.install(monkeypatch)
    pipe = make_pipeline(tiny_config(), seed=0)
    seen = []

    class Halve77(p2p.AttentionControl):                       # reference-style: implements forward() only
        def forward(self, attn, is_cross, place_in_unet):
            seen.append((is_cross, tuple(attn.shape)))
            return attn * 0.5 if is_cross else attn

    ctrl = Halve77(False)
    p2p.register_attention_control(pipe, ctrl)
    x = torch.randn(4, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    ctx = editing.encode_prompts(pipe, ["a cat", "a dog"])
    with torch.no_grad():
        got = pipe.unet(x, 981, encoder_hidden_states=ctx).sample
    assert ctrl._needs_probabilities() and ctrl.cur_step == 1 and ctrl.cur_att_layer == 0 and len(seen) == ctrl.num_att_layers == 32
    assert all(shape[0] == 2 * 2 for _, shape in seen)          # the conditional half only: 2 rows x 2 heads
    p2p.unregister_attention_control(pipe, ctrl)

    class Plain(p2p.AttentionControl):
        def forward(self, attn, is_cross, place_in_unet):
            return attn

    plain = Plain(False)
    p2p.register_attention_control(pipe, plain)
    with torch.no_grad():
        same = pipe.unet(x, 981, encoder_hidden_states=ctx).sample
    p2p.unregister_attention_control(pipe, plain)
    empty = p2p.EmptyControl(False)
    p2p.register_attention_control(pipe, empty)
    with torch.no_grad():
        want = pipe.unet(x, 981, encoder_hidden_states=ctx).sample
    p2p.unregister_attention_control(pipe, empty)
    assert torch.allclose(same, want, atol=1e-5) and not torch.allclose(got, want, atol=1e-3)

    class HalfReplace(p2p.AttentionReplace):                  # overrides the edit itself
        def replace_cross_attention(self, attn_base, att_replace):
            return 0.5 * super().replace_cross_attention(attn_base, att_replace) + 0.5 * att_replace

    prompts = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]
    custom = HalfReplace(prompts, pipe.tokenizer, 4, 0.8, 0.6, device="cpu")
    stock = p2p.AttentionReplace(prompts, pipe.tokenizer, 4, 0.8, 0.6, device="cpu")
    assert custom._needs_probabilities() and not stock._needs_probabilities()
    ctx = editing.encode_prompts(pipe, prompts)
    outs = []
    for c in (custom, stock):
        p2p.register_attention_control(pipe, c)
        with torch.no_grad():
            outs.append(pipe.unet(x, 981, encoder_hidden_states=ctx).sample)
        p2p.unregister_attention_control(pipe, c)
    assert not torch.allclose(outs[0], outs[1], atol=1e-4)      # the override took effect
    assert torch.equal(outs[0][:3], outs[1][:3]) or torch.allclose(outs[0][:3], outs[1][:3], atol=1e-5)   # only the target row is edited


def test_attention_mask_is_rejected(monkeypatch):
    cpu_backend.install(monkeypatch)
    pipe = make_pipeline(tiny_config(), seed=0)
    ctrl = p2p.EmptyControl(False)
    p2p.register_attention_control(pipe, ctrl)
    attn = pipe.unet.mid_block.attentions[0].transformer_blocks[0].attn1
    with pytest.raises(NotImplementedError):
        attn.forward(torch.zeros(1, 4, attn.to_q.in_features), attention_mask=torch.zeros(1, 4))


def test_graph_replay_protocol_tracks_eager_counters(monkeypatch):
    """graph_advance() must leave a controller exactly where a complete eager forward leaves it, and graph_key() must change
    whenever the host state that shapes the launches changes (self-replace window, first stored step, MasaCtrl step list)."""
    import copy
    from image_editing_framework_b200 import p2p, masactrl, editing
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    cpu_backend.install(monkeypatch)
    dev = torch.device("cpu")
    pipe = make_pipeline(tiny_config(), seed=1, device=dev)
    prompts = ["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"]
    steps = 5
    pipe.scheduler.set_timesteps(steps)
    ctx = editing.encode_prompts(pipe, prompts)
    x = torch.randn(4, 4, 8, 8)
    lb = p2p.LocalBlend(pipe.tokenizer, prompts, [["cat"], ["dog"]], device=dev)
    ctrl = p2p.AttentionReplace(prompts, pipe.tokenizer, steps, 0.8, 0.6, lb, device=dev)
    p2p.register_attention_control(pipe, ctrl)
    keys = []
    with torch.no_grad():
        for t in pipe.scheduler.timesteps.tolist():
            keys.append(ctrl.graph_key())
            shadow = copy.copy(ctrl)
            shadow.graph_advance()
            pipe.unet(x, t, encoder_hidden_states=ctx)
            assert (ctrl.cur_step, ctrl.cur_att_layer, ctrl._slot) == (shadow.cur_step, shadow.cur_att_layer, shadow._slot)
    # store: first step allocates, later accumulate; self-replace window = steps [0, 3)
    # (the trailing 0 is the edit-table epoch: it only moves when retarget() had to replace a device table instead of rewriting it)
    assert keys == [(("store", False), True, 0), (("store", True), True, 0), (("store", True), True, 0), (("store", True), False, 0),
                    (("store", True), False, 0)]
    p2p.unregister_attention_control(pipe, ctrl)
    ed = masactrl.MutualSelfAttentionControl(2, 10, total_steps=steps)
    masactrl.regiter_attention_editor_diffusers(pipe, ed)
    keys = []
    with torch.no_grad():
        for t in pipe.scheduler.timesteps.tolist():
            keys.append(ed.graph_key())
            shadow = copy.copy(ed)
            shadow.graph_advance()
            pipe.unet(x, t, encoder_hidden_states=ctx)
            assert (ed.cur_step, ed.cur_att_layer) == (shadow.cur_step, shadow.cur_att_layer)
    assert keys == [(False,), (False,), (True,), (True,), (True,)]
    assert masactrl.AttentionStore().graph_key() is None


def test_pix2pix_zero_loops_reproduce_reference(monkeypatch):
    """editing.pix2pix_zero_edit: map collection with the cache kept on the device, guidance gradient through the torch pass, SGD
    step on the latents, recomputed noise, DDIM step — against the loops restated around the reference's own processor."""
    cpu_backend.install(monkeypatch)
    g = golden("pix2pix_zero_loop.pt")
    rec, edit = scenarios.run_pix2pix_zero_loop(g, torch.device("cpu"))
    assert torch.allclose(rec, g["rec"], atol=2e-4), (rec - g["rec"]).abs().max()
    assert torch.allclose(edit, g["edit_per_step"][-1], atol=5e-4), (edit - g["edit_per_step"][-1]).abs().max()
    # the guidance matters: without it the edit loop lands elsewhere
    moved = (g["edit_no_guidance"] - g["edit_per_step"][-1]).abs().max()
    assert moved > 20 * (edit - g["edit_per_step"][-1]).abs().max(), moved


def test_cross_edit_sparse_mapper_form():
    """ops.CrossEdit.sparsify: per target token the (<= 8) contributing source tokens in ascending order, -1 padded; a mapper
    with a denser column yields (None, None) and the kernel multiplies the dense form."""
    from image_editing_framework_b200 import ops
    m = torch.eye(77).repeat(2, 1, 1)
    m[0, 3, 3], m[0, 3, 4], m[0, 3, 5] = 0.0, 0.5, 0.5
    m[1, 10:14] = torch.rand(4, 77, generator=torch.Generator().manual_seed(0)).softmax(-1)
    idx, w = ops.CrossEdit.sparsify(m)
    assert idx.shape == (2, 77, 8) and idx.dtype == torch.int32 and w.dtype == torch.float32
    dense = torch.zeros_like(m)
    for s_ in range(2):
        for n in range(77):
            nz = idx[s_, n][idx[s_, n] >= 0]
            assert torch.equal(nz, nz.sort().values)
            dense[s_, nz.long(), n] = w[s_, n, :len(nz)]
    assert torch.equal(dense, m)
    assert idx[0, 3].tolist() == [-1] * 8 and idx[0, 4].tolist()[:2] == [3, 4]
    m[1, 10:30] = 1.0 / 77
    assert ops.CrossEdit.sparsify(m) == (None, None)


def test_lean_store_keeps_localblend_result(monkeypatch):
    """controller.lean_store = True (opt-in): only the 16x16 cross maps LocalBlend reads are stored, the other entries are None
    placeholders at the reference's list positions; the edit result is unchanged."""
    cpu_backend.install(monkeypatch)
    g = golden("p2p_localblend.pt")
    full = scenarios.run_p2p_localblend(g, torch.device("cpu"))
    lean = scenarios.run_p2p_localblend(g, torch.device("cpu"), lean_store=True)
    assert torch.equal(full[1][-1], lean[1][-1])
    store = lean[0].attention_store
    assert all(m is None for m in store["down_self"] + store["up_self"] + store["mid_self"])
    kept = [m for key in ("down_cross", "up_cross", "mid_cross") for m in store[key] if m is not None]
    assert len(kept) == 5 and all(m.shape[1] == 256 for m in kept)
    assert [m is None for m in store["down_cross"]] == [True, True, False, False]


# ------------------------------------------------------------------------------------------------ SDXL inversion, image2latent
class _XLPipeline:
    """The handful of StableDiffusionXLPipeline members */inversion/ddim.py:60-109 touches, over the stand-in UNet."""

    def __init__(self):
        self._p = make_pipeline(tiny_config(), seed=3)
        self.unet, self.scheduler, self.vae = self._p.unet, self._p.scheduler, self._p.vae
        self.seen = []
        inner = self.unet.forward

        def recording_forward(sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None, **kw):
            self.seen.append((int(timestep), cross_attention_kwargs, added_cond_kwargs))
            return inner(sample, timestep, encoder_hidden_states)
        self.unet.forward = recording_forward

    @property
    def _execution_device(self):
        return self.unet.device

    def encode_prompt(self, prompt, device, **kw):
        assert kw["do_classifier_free_guidance"] is True and kw["prompt_2"] is None and kw["num_images_per_prompt"] == 1
        pe, ne = self._p.encode_prompt(prompt, device)
        return pe, ne, pe.mean(1), ne.mean(1)

    def _get_add_time_ids(self, original_size, crops, target_size, dtype):
        return torch.tensor([list(original_size) + list(crops) + list(target_size)], dtype=dtype)


def test_ddim_inversion_xl_and_image2latent(monkeypatch):
    cpu_backend.install(monkeypatch)
    from image_editing_framework_b200.ddim import ddim_inversion, ddim_inversion_xl
    model = _XLPipeline()
    model.scheduler.set_timesteps(6)
    x0 = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        traj, ctx = ddim_inversion_xl().ddim_inversion_loop(model, x0, ["a cat on a table"], height=64, width=64)
    steps = model.scheduler.timesteps.tolist()
    assert len(traj) == 7 and len(ctx) == 4 and [s[0] for s in model.seen] == steps[::-1]
    for _, cak, extra in model.seen:
        assert cak is None and torch.equal(extra["text_embeds"], ctx[2]) and extra["time_ids"].tolist() == [[64, 64, 0, 0, 64, 64]]
    # every hop is the reference's ddim_reverse closed form (inversion/ddim.py:9-18) on the UNet's prediction for the conditional prompt
    ac, stride = model.scheduler.alphas_cumprod, 1000 // 6
    with torch.no_grad():
        for i, t in enumerate(reversed(steps)):
            eps = model._p.unet(traj[i], t, encoder_hidden_states=ctx[0]).sample
            cur = min(999, t - stride)
            a_cur = ac[cur] if cur >= 0 else model.scheduler.final_alpha_cumprod
            assert torch.allclose(traj[i + 1], orc.ddim_step(eps, traj[i], a_cur, ac[t]), atol=1e-6, rtol=1e-5)
    if reference_loader.reference_available():
        ref = reference_loader.load_reference("p2p").ddim
        live = _XLPipeline()
        live.scheduler.set_timesteps(6)
        want, _ = ref.ddim_inversion_xl().ddim_inversion_loop(live, x0, ["a cat on a table"], height=64, width=64)
        assert all(torch.allclose(a, b, atol=1e-5, rtol=1e-5) for a, b in zip(traj, want))
    image = np.random.default_rng(0).integers(0, 256, (64, 64, 3), dtype=np.uint8)
    z = ddim_inversion().image2latent(model, image, "cpu", torch.float32)
    pix = torch.from_numpy(image).float() / 127.5 - 1
    assert torch.allclose(z, model.vae.encode(pix.permute(2, 0, 1)[None])["latent_dist"].mean * model.vae.config.scaling_factor)
    if reference_loader.reference_available():
        assert torch.equal(z, ref.ddim_inversion().image2latent(model, image, "cpu", torch.float32))


# ------------------------------------------------------------------------------------------------ pipeline-level classes
@pytest.mark.parametrize("family,name", scenarios.PIPELINE_CASES)
def test_pipeline_classes_reproduce_reference_images(monkeypatch, family, name):
    """P2P / MasaCtrl / PnP / P2P_Zero and their _XL / _NTI variants (`*/model/sd_utils.py`) against images the reference's own
    classes produced on the same stand-in (tests/golden/pipelines.pt): host logic of the mirrors on oracle-backed ops."""
    cpu_backend.install(monkeypatch)
    want = golden("pipelines.pt")[name]
    got, _ = scenarios.run_pipeline_case(family, name, scenarios.mirror_api(family), torch.device("cpu"))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.dtype == np.uint8 and np.abs(a.astype(np.int16) - b.numpy().astype(np.int16)).max() <= 1


def test_reference_script_flow_after_import_switch(monkeypatch):
    """The masactrl/edit_real.py flow INTEGRATION.md shows (image -> latent -> DDIM inversion -> register editor -> MasaCtrl sampler),
    and its null-text variant, run unchanged on this package's names."""
    cpu_backend.install(monkeypatch)
    from image_editing_framework_b200.ddim import ddim_inversion
    from image_editing_framework_b200.nti import NTI
    from image_editing_framework_b200.masactrl import (MasaCtrl, MasaCtrl_NTI, MutualSelfAttentionControl, regiter_attention_editor_diffusers,
                                                       unregister_attention_control)
    pipe = make_pipeline(tiny_config(), seed=0)
    steps = 3
    image = np.random.default_rng(0).integers(0, 256, (64, 64, 3), dtype=np.uint8)
    source_prompt, target_prompt = ["a cat"], ["a dog"]
    for invertor, editor in ((ddim_inversion(), MasaCtrl(pipe, steps)), (NTI(), MasaCtrl_NTI(pipe, steps))):
        latent = invertor.image2latent(model=pipe, image=image, device="cpu", dtype=torch.float32)
        latents, context = invertor.ddim_inversion_loop(pipe, latent, source_prompt)
        extra = {}
        if isinstance(invertor, NTI):
            extra["uncond_embeddings_list"] = invertor.null_optimization(pipe, latents, context, 2, 1e-5, 7.5)
        controller = MutualSelfAttentionControl(1, 10, total_steps=steps, model_type="SD")
        regiter_attention_editor_diffusers(editor.model, controller)
        images, x_t = editor(prompt=source_prompt + target_prompt, latents=torch.cat([latents[-1]] * 2), guidance_scale=7.5,
                             num_inference_steps=steps, height=64, width=64, **extra)
        unregister_attention_control(pipe, controller)
        assert images.shape == (2, 64, 64, 3) and images.dtype == np.uint8 and controller.cur_step == steps
        assert not np.array_equal(images[0], images[1])       # two prompts, two images


@pytest.mark.parametrize("field,value", [("prediction_type", "v_prediction"), ("clip_sample", True), ("thresholding", True)])
def test_fused_ddim_refuses_scheduler_configs_it_does_not_implement(field, value):
    """scheduler.step honours prediction_type / clip_sample / thresholding (p2p/model/sd_utils.py:76 calls it); the fused step implements
    the one configuration the reference's scripts build (p2p/edit_real.py:58-69) and must say so for any other."""
    from image_editing_framework_b200.ddim import FusedDDIM
    from image_editing_framework_b200.standin.pipeline import DDIMScheduler
    sched = DDIMScheduler()
    sched.set_timesteps(10)
    FusedDDIM(sched)                                   # the scripts' configuration is accepted
    setattr(sched.config, field, value)
    with pytest.raises(ValueError, match=field):
        FusedDDIM(sched)


@pytest.mark.parametrize("cfg_name,kind", [("sd15_slim", "masactrl_union"), ("sd15_slim", "p2p_refine")])
def test_full_geometry_host_logic_reproduces_reference(monkeypatch, cfg_name, kind):
    """The 64x64-latent, real-head-dim scenarios (tests/golden/fullgeo_*.pt) through this package's closures + controllers with the
    oracle-backed fp32 ops: the row tables / second key block (Union) and the edit tables (refine) at the BASELINE geometry, 2e-4."""
    import image_editing_framework_b200 as pkg
    from image_editing_framework_b200 import masactrl, pnp  # noqa: F401
    cpu_backend.install(monkeypatch)
    g = golden(f"fullgeo_{cfg_name}_{kind}.pt")
    api = {"p2p": pkg.p2p, "masactrl": pkg.masactrl, "pnp": pkg.pnp}[kind.split("_")[0]]
    ctrl, records, per_step = scenarios.run_fullgeo(cfg_name, kind, api, torch.device("cpu"))
    for step, outs in g["layer_outputs"].items():
        for i, (a, b) in enumerate(zip(records[step], outs)):
            assert (a - b).abs().max().item() < 2e-4, f"layer {i}"
    for a, b in zip(per_step, g["latents_per_step"]):
        assert (a - b).abs().max().item() < 1e-3
    assert ctrl.cur_step == g["cur_step"]


def test_bench_reference_arm_prints_one_contract_line():
    """`python bench.py --impl reference` (debug-sized) prints exactly ONE JSON line on stdout with the contract's keys, the same `config`
    object our arm prints, and a step time that is a real wall time (not the extrapolated edit)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "tiny", "--ddim-steps", "4", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "edits/s" and line["vs_baseline"] is None and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["extrapolated"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    assert line["config"] == bench.workload_config(4, "tiny")
    assert line["ms_per_step"] < 1e3 * 60 and line["ms_per_edit_extrapolated"] > 0
