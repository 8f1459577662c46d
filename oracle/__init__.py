"""CPU restatement of the reference's controlled-attention arithmetic — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs may import this
package. Nothing under `image_editing_framework_b200/` imports it: the product path has no CPU fallback.

The reference is pure Python (no native code), so the oracle is torch fp32 on the CPU. It is pinned against outputs
of the reference's own code run in the build container (tests/golden/make_goldens.py -> tests/golden/*.pt): the
reference ships no tests, golden vectors or known-answer files of its own.
"""
