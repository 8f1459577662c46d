"""Host-side integer code against the LIVE reference (only where /root/reference is mounted, i.e. in the build container; skipped on
the GPU box): random prompt pairs with multi-token words, insertions, deletions and swaps through seq_aligner / ptp_utils of both
trees — mappers, alphas, word indices, equalizers and the time-words alpha table must be bit-identical."""
import random

import pytest
import torch

from oracle import reference_loader
from image_editing_framework_b200.p2p import seq_aligner, ptp_utils
from image_editing_framework_b200.standin import WordPieceTokenizer

pytestmark = pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not mounted")

WORDS = ["a", "photo", "of", "cat", "dog", "squirrel", "hippopotamus", "burger", "eating", "sitting", "on", "the", "bench", "watercolor",
         "painting", "house", "large", "lake", "by", "frozen", "river", "children", "playing", "near", "sunset", "at", "bronze", "wooden", "horse"]


def _prompt(rng, n):
    return " ".join(rng.choice(WORDS) for _ in range(n))


def test_random_prompts_match_live_reference():
    ref = reference_loader.load_reference("p2p")
    tok = WordPieceTokenizer()
    rng = random.Random(1234)
    for case in range(120):
        n = rng.randint(1, 9)
        src = _prompt(rng, n)
        # refinement: random insertions / deletions / substitutions
        tgt_words = []
        for w in src.split(" "):
            r = rng.random()
            if r < 0.15:
                continue
            tgt_words.append(rng.choice(WORDS) if r < 0.3 else w)
            if rng.random() < 0.2:
                tgt_words.append(rng.choice(WORDS))
        tgt = " ".join(tgt_words) or "a"
        m, a = seq_aligner.get_refinement_mapper([src, tgt], tok)
        rm, ra = ref.seq_aligner.get_refinement_mapper([src, tgt], tok)
        assert torch.equal(m, rm) and torch.equal(a, ra), (src, tgt)
        # replacement: same word count, some words swapped (token counts may differ)
        swp = " ".join(rng.choice(WORDS) if rng.random() < 0.3 else w for w in src.split(" "))
        r1 = seq_aligner.get_replacement_mapper([src, swp], tok)
        r2 = ref.seq_aligner.get_replacement_mapper([src, swp], tok)
        assert torch.equal(r1, r2), (src, swp)
        for pos, w in enumerate(tgt.split(" ")):
            assert seq_aligner.get_word_inds(tgt, w, tok).tolist() == ref.seq_aligner.get_word_inds(tgt, w, tok).tolist()
            assert seq_aligner.get_word_inds(tgt, pos, tok).tolist() == ref.seq_aligner.get_word_inds(tgt, pos, tok).tolist()
        w = rng.choice(tgt.split(" "))
        e1 = seq_aligner.get_equalizer(tok, tgt, (w,), (rng.choice([0.5, 2.0, -1.0]),))
        # same draw for both trees
        val = float(e1.flatten()[seq_aligner.get_word_inds(tgt, w, tok)[0]]) if len(seq_aligner.get_word_inds(tgt, w, tok)) else 1.0
        e2 = ref.seq_aligner.get_equalizer(tok, tgt, (w,), (val,))
        assert torch.equal(e1, e2), (tgt, w)
        steps = rng.choice([3, 10, 50])
        spec = rng.choice([0.8, (0.2, 0.7), {"default_": 0.8, rng.choice(swp.split(" ")): (0.0, 0.4)}])
        t1 = ptp_utils.get_time_words_attention_alpha([src, swp], steps, spec, tok)
        t2 = ref.ptp_utils.get_time_words_attention_alpha([src, swp], steps, spec, tok)
        assert torch.equal(t1, t2), (src, swp, spec)
