"""raytracing-1w on B200 — host-side Python mirror of the reference interface.

The product is the C-ABI library `_build/librt1w.so` (include/rt1w.h): hand-written sm_100a
kernels behind the reference's scene-construction surface.  This package only
  * loads that library with ctypes (and FAILS LOUDLY if it is missing — there is no CPU
    or PyTorch fallback),
  * mirrors the POD structs of include/rt1w.h,
  * gives Python access to the C++ mirror of the reference's scene functions
    (`_build/librt1w_host.so`, main.rs:192-795 / 815-937).

The package directory name contains a hyphen (it is the name the build contract asks for), so
import it with `importlib.import_module("raytracing-1w_b200")`.
"""
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
