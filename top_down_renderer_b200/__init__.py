"""top_down_renderer_b200 — B200-native (sm_100a) localization hot path of KumarRobotics/top_down_renderer.

The product is libtdr_b200.so (hand-written CUDA behind the C ABI of include/tdr.h); this package is the
host-side mirror of the reference's class interfaces plus the ctypes binding.  No CPU fallback exists.
"""
from ._lib import LIB_PATH, TdrError, build, build_hostmath, load  # noqa: F401

__all__ = ["build", "build_hostmath", "load", "LIB_PATH", "TdrError"]
