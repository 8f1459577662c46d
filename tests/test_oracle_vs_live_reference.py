"""The CPU oracle against the LIVE reference controllers on random inputs (only where /root/reference is mounted; skipped elsewhere):
random prompts, step, head count, map sizes through AttentionReplace / AttentionRefine / AttentionReweight(.__call__) of the reference
and oracle.controlled_attention.p2p_edit_probs — the function every GPU parity test is measured against."""
import random

import pytest
import torch

from oracle import reference_loader
from oracle import controlled_attention as orc
from image_editing_framework_b200.standin import WordPieceTokenizer

pytestmark = pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not mounted")
WORDS = ["a", "photo", "of", "cat", "dog", "squirrel", "hippopotamus", "burger", "eating", "sitting", "on", "the", "bench", "large", "house"]
CPU = torch.device("cpu")


def test_random_p2p_edits_match_live_reference():
    ref = reference_loader.load_reference("p2p")
    AC = ref.attention_control
    tok = WordPieceTokenizer()
    rng = random.Random(99)
    g = torch.Generator().manual_seed(7)
    for case in range(40):
        n_prompts = rng.choice([2, 3])
        nw = rng.randint(2, 6)
        src = " ".join(rng.choice(WORDS) for _ in range(nw))
        kind = rng.choice(["replace", "refine", "reweight", "reweight_chain"])
        if kind == "refine":
            prompts = [src] + [src + " " + " ".join(rng.choice(WORDS) for _ in range(rng.randint(1, 2))) for _ in range(n_prompts - 1)]
        else:
            prompts = [src] + [" ".join(rng.choice(WORDS) if rng.random() < 0.4 else w for w in src.split(" ")) for _ in range(n_prompts - 1)]
        steps = rng.choice([5, 10])
        cross_steps = rng.choice([0.8, (0.1, 0.6)])
        self_steps = rng.choice([0.4, (0.2, 0.8)])
        kw = dict(prompts=prompts, tokenizer=tok, num_steps=steps, cross_replace_steps=cross_steps, self_replace_steps=self_steps, device=CPU)
        okw = {}
        if kind == "replace":
            ctrl = AC.AttentionReplace(**kw)
            okw = dict(mode="replace", mapper=ctrl.mapper)
        elif kind == "refine":
            ctrl = AC.AttentionRefine(**kw)
            okw = dict(mode="refine", mapper=ctrl.mapper, refine_alphas=ctrl.alphas)
        else:
            eq = ref.seq_aligner.get_equalizer(tok, prompts[1], (prompts[1].split(" ")[-1],), (rng.choice([2.0, 0.5, -1.0]),))
            eq = eq.expand(n_prompts - 1, -1).contiguous()
            inner = AC.AttentionReplace(**kw) if kind == "reweight_chain" else None
            ctrl = AC.AttentionReweight(equalizer=eq, controller=inner, **kw)
            okw = dict(mode="replace" if inner is not None else "none", equalizer=ctrl.equalizer)
            if inner is not None:
                okw["mapper"] = inner.mapper
        H = rng.choice([1, 2, 4])
        for is_cross in (True, False):
            N = rng.choice([16, 64, 256, 300])
            M = 77 if is_cross else N
            probs = torch.randn(2 * n_prompts * H, N, M, generator=g).softmax(-1)
            step = rng.randrange(steps)
            ctrl.num_att_layers, ctrl.cur_step, ctrl.cur_att_layer = 100, step, 0
            want = ctrl(probs.clone(), is_cross, "down")
            got = orc.p2p_edit_probs(probs, H, n_prompts, is_cross, step, alpha_table=ctrl.cross_replace_alpha,
                                     num_self_replace=ctrl.num_self_replace, **okw)
            assert torch.allclose(got, want, atol=1e-6, rtol=1e-5), (case, kind, is_cross, (got - want).abs().max().item())
