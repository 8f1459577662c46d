"""B200-native controlled-attention hot path behind the reference's controller / register API.

Sub-packages mirror the four method directories of AY-Liu/Image-Editing-Framework:
  p2p/  masactrl/  pnp/  pix2pix_zero/      (+ ddim.py for the fused CFG/DDIM step, standin/ for the absent diffusers)
All arithmetic of the hot path runs in libief_b200.so (hand-written sm_100a CUDA, include/ief_b200.h).
"""
from . import _cabi
from ._cabi import build_library, IefError

__all__ = ["_cabi", "build_library", "IefError"]
