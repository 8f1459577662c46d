"""Randomised parity sweeps (tools/fuzz_*.py) with fixed seeds, as part of the GPU suite: random shapes, ragged sizes, row sources,
second key blocks, key bias, map output, row masks / store slots / strided operands, edit flavours and the backward, against fp32 materialised references computed with torch on
the same GPU. Each tool exits non-zero on the first mismatch and prints the failing configuration."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,seed,cases", [("fuzz_attn.py", 101, 60), ("fuzz_cross.py", 102, 60), ("fuzz_elementwise.py", 103, 20),
                                             ("fuzz_attn_rows.py", 104, 60)])
def test_randomised_parity_sweep(cuda, tool, seed, cases):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(seed), str(cases)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "ok" in r.stdout
