"""ctypes binding of libief_b200.so (the C ABI declared in include/ief_b200.h).

This is the *only* way the Python host side reaches the CUDA kernels; it is also the stub a
maintainer of the reference would add. Nothing here falls back to
torch/CPU arithmetic: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

IEF_ABI_VERSION = 5
IEF_MAX_ROWS = 64

IEF_BF16, IEF_F16, IEF_F32 = 0, 1, 2
IEF_IMPL_AUTO, IEF_IMPL_MMA, IEF_IMPL_TCGEN05 = 0, 1, 2
IEF_EDIT_NONE, IEF_EDIT_REPLACE, IEF_EDIT_REFINE = 0, 1, 2

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# IEF_LIB_PATH: an alternative build of the same library (A/B kernel variants from tools/build_variants.sh); never a different backend
LIB_PATH = os.environ.get("IEF_LIB_PATH") or os.path.join(_PKG_DIR, "libief_b200.so")
BUILD_SCRIPT = os.path.join(_PKG_DIR, "csrc", "build.sh")


class Tensor4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_n", C.c_int64), ("stride_h", C.c_int64)]


class AttnParams(C.Structure):
    _fields_ = [
        ("q", Tensor4), ("k", Tensor4), ("v", Tensor4), ("o", Tensor4),
        ("dtype", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("d", C.c_int32),
        ("scale", C.c_float),
        ("impl", C.c_int32),
        ("q_src", C.POINTER(C.c_int32)), ("k_src", C.POINTER(C.c_int32)), ("v_src", C.POINTER(C.c_int32)),
        ("k_src2", C.POINTER(C.c_int32)), ("v_src2", C.POINTER(C.c_int32)),
        ("probs_out", C.c_void_p),
        ("probs_accum", C.c_int32),
        ("probs_slot", C.POINTER(C.c_int32)),
        ("row_mask", C.POINTER(C.c_uint8)),
        ("key_bias", C.c_void_p),
        ("bias_sel", C.POINTER(C.c_int32)),
        ("n_bias", C.c_int32),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
    ]


class CrossParams(C.Structure):
    _fields_ = [
        ("q", Tensor4), ("k", Tensor4), ("v", Tensor4), ("o", Tensor4),
        ("dtype", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("d", C.c_int32),
        ("scale", C.c_float),
        ("mode", C.c_int32),
        ("base_row", C.POINTER(C.c_int32)), ("edit_slot", C.POINTER(C.c_int32)),
        ("n_slots", C.c_int32),
        ("mapper", C.c_void_p), ("mapper_nz_idx", C.c_void_p), ("mapper_nz_w", C.c_void_p), ("mapper_idx", C.c_void_p), ("refine_alpha", C.c_void_p),
        ("equalizer", C.c_void_p), ("step_alpha", C.c_void_p),
        ("probs_out", C.c_void_p),
        ("probs_accum", C.c_int32),
        ("store_slot", C.POINTER(C.c_int32)),
    ]


class CrossBwdParams(C.Structure):
    _fields_ = [
        ("q", Tensor4), ("k", Tensor4), ("v", Tensor4), ("dout", Tensor4), ("dq", Tensor4),
        ("dtype", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32), ("d", C.c_int32),
        ("scale", C.c_float),
        ("dprobs", C.c_void_p),
        ("ds_out", C.c_void_p),
    ]


class LocalBlendParams(C.Structure):
    _fields_ = [
        ("maps", C.POINTER(C.c_void_p)), ("map_heads", C.POINTER(C.c_int32)),
        ("n_maps", C.c_int32), ("n_prompts", C.c_int32), ("res", C.c_int32), ("n_words", C.c_int32),
        ("word_alpha", C.c_void_p),
        ("threshold", C.c_float),
        ("x_t", C.c_void_p),
        ("C", C.c_int32), ("Hx", C.c_int32), ("Wx", C.c_int32),
        ("workspace", C.c_void_p),
        ("mask_out", C.c_void_p),
    ]


class UmmaProbeParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p),
        ("N", C.c_int32), ("K", C.c_int32),
        ("b_mn_major", C.c_int32), ("a_from_tmem", C.c_int32), ("dtype", C.c_int32),
    ]


# every symbol include/ief_b200.h declares; tests check the .so exports each one
EXPORTS = (
    "ief_attn_fwd", "ief_attn_workspace_bytes", "ief_cross_attn_edit_fwd", "ief_cross_attn_bwd", "ief_store_accumulate", "ief_local_blend", "ief_mask_blend", "ief_cfg_ddim_step",
    "ief_umma_probe", "ief_abi_version", "ief_last_error", "ief_launch_count", "ief_last_attn_impl", "ief_last_cross_impl", "ief_check_device",
)

_lib = None
_lock = threading.Lock()


class IefError(RuntimeError):
    """A libief_b200 call returned a negative status."""

    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with status {code}: {msg}")
        self.code = code


def build_library(force: bool = False) -> str:
    """Compile libief_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH) or _stale():
        subprocess.run(["bash", BUILD_SCRIPT, _PKG_DIR], check=True)
    return LIB_PATH


def _stale() -> bool:
    try:
        t = os.path.getmtime(LIB_PATH)
    except OSError:
        return True
    src_dirs = [os.path.join(_PKG_DIR, "csrc"), os.path.join(_PKG_DIR, "..", "include")]
    for d in src_dirs:
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h", ".sh")) and os.path.getmtime(os.path.join(d, f)) > t:
                return True
    return False


def lib() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `bash {BUILD_SCRIPT}`. There is no CPU/torch fallback for the controlled-attention hot path.")
        L = C.CDLL(LIB_PATH)
        L.ief_abi_version.restype = C.c_int
        L.ief_last_error.restype = C.c_char_p
        L.ief_last_attn_impl.restype = C.c_char_p
        L.ief_last_cross_impl.restype = C.c_char_p
        L.ief_launch_count.restype = C.c_int64
        L.ief_check_device.restype = C.c_int
        L.ief_attn_fwd.argtypes = [C.POINTER(AttnParams), C.c_void_p]
        L.ief_attn_fwd.restype = C.c_int
        L.ief_attn_workspace_bytes.argtypes = [C.POINTER(AttnParams)]
        L.ief_attn_workspace_bytes.restype = C.c_int64
        L.ief_cross_attn_edit_fwd.argtypes = [C.POINTER(CrossParams), C.c_void_p]
        L.ief_cross_attn_edit_fwd.restype = C.c_int
        L.ief_cross_attn_bwd.argtypes = [C.POINTER(CrossBwdParams), C.c_void_p]
        L.ief_cross_attn_bwd.restype = C.c_int
        L.ief_store_accumulate.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_void_p]
        L.ief_store_accumulate.restype = C.c_int
        L.ief_local_blend.argtypes = [C.POINTER(LocalBlendParams), C.c_void_p]
        L.ief_local_blend.restype = C.c_int
        L.ief_cfg_ddim_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.ief_cfg_ddim_step.restype = C.c_int
        L.ief_mask_blend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                     C.POINTER(C.c_uint8), C.c_void_p]
        L.ief_mask_blend.restype = C.c_int
        L.ief_umma_probe.argtypes = [C.POINTER(UmmaProbeParams), C.c_void_p]
        L.ief_umma_probe.restype = C.c_int
        if L.ief_abi_version() != IEF_ABI_VERSION:
            raise ImportError(f"libief_b200.so ABI {L.ief_abi_version()} != binding ABI {IEF_ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(fn: str, code: int) -> None:
    if code != 0:
        raise IefError(fn, code, lib().ief_last_error().decode("utf-8", "replace"))


def i32_array(values):
    if values is None:
        return None
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr


def launch_count() -> int:
    return int(lib().ief_launch_count())


def last_attn_impl() -> str:
    return lib().ief_last_attn_impl().decode()


def last_cross_impl() -> str:
    return lib().ief_last_cross_impl().decode()
