"""radiancecascade_b200 — B200-native radiance-cascade GI hot path behind a C ABI.

The product is radiancecascade_b200/librc_b200.so (include/rc_b200.h, built from
csrc/ for sm_100a).  This package is the thin host-side mirror of the reference's
renderer interface (renderer.py) plus the bundled-scene helpers (scenes.py).
Importing the package does not need a GPU; creating a renderer does, and fails
loudly otherwise — there is no CPU fallback.
"""
from . import _ffi, scenes
from ._ffi import RcError
from .renderer import AppState, Camera, CascadeConfig, DefaultRenderer, ObjScene, Projection, UniformCamera

__all__ = ["AppState", "Camera", "CascadeConfig", "DefaultRenderer", "ObjScene", "Projection", "UniformCamera", "RcError",
           "scenes", "_ffi"]
