"""Import the reference's own Python modules (read-only, from /root/reference) under private names.

TEST INFRASTRUCTURE, usable ONLY in the build container: /root/reference does not exist on the GPU box, so nothing in
`-m gpu` tests, smoke() or bench.py calls this. It is used by tests/golden/make_goldens.py to produce the committed
fixtures and by the CPU tests that cross-check the oracle against the live reference when it is present.

All four method directories use the top-level package name `model` and need `diffusers` only for names
(SURVEY.md section 8c), so each load (a) installs a tiny `diffusers` name shim backed by our stand-in classes,
(b) imports `model.*` with the method directory first on sys.path, (c) renames the loaded modules out of the way.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("IEF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "p2p", "model"))


def _install_diffusers_shim():
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_ief_shim", False):
        return  # a real diffusers is importable: use it
    from image_editing_framework_b200.standin import unet as su
    import torch

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, dtype=dtype).to(device)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._ief_shim = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("diffusers")
    mod("diffusers.models")
    mod("diffusers.models.resnet", Upsample2D=su.Upsample2D, Downsample2D=su.Downsample2D)
    mod("diffusers.models.attention_processor", Attention=su.Attention)
    mod("diffusers.utils")
    mod("diffusers.utils.torch_utils", randn_tensor=randn_tensor)


_CACHE = {}
_MODULES = {
    "p2p": ["ptp_utils", "seq_aligner", "attention_base", "attention_control", "register", "sd_utils"],
    "masactrl": ["attention_base", "attention_control", "register", "sd_utils"],
    "pnp": ["register", "sd_utils"],
    "pix2pix-zero": ["attention_control", "sd_utils"],
}


def load_reference(method: str) -> SimpleNamespace:
    """Namespace with the reference's `model.<name>` modules of one method plus `ddim` / `nti` (= inversion.ddim / inversion.nti)."""
    if method in _CACHE:
        return _CACHE[method]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_diffusers_shim()
    root = os.path.join(REFERENCE_ROOT, method)
    stale = [k for k in sys.modules if k in ("model", "inversion") or k.startswith(("model.", "inversion."))]
    for k in stale:
        del sys.modules[k]
    sys.path.insert(0, root)
    sys.dont_write_bytecode, old = True, sys.dont_write_bytecode  # never write __pycache__ into the read-only tree
    try:
        ns = {}
        for name in _MODULES[method]:
            ns[name] = importlib.import_module(f"model.{name}")
        ns["ddim"] = importlib.import_module("inversion.ddim")
        ns["nti"] = importlib.import_module("inversion.nti")
    finally:
        sys.dont_write_bytecode = old
        sys.path.remove(root)
        tag = "ref_" + method.replace("-", "_")
        for k in [k for k in sys.modules if k in ("model", "inversion") or k.startswith(("model.", "inversion."))]:
            sys.modules[f"{tag}.{k}"] = sys.modules.pop(k)
    _CACHE[method] = SimpleNamespace(**ns)
    return _CACHE[method]
