"""globalillumination_b200 — B200-native (sm_100a) shadow hot path behind the C ABI of include/shadowgi.h.

  csrc/   hand-written CUDA kernels + the C ABI (libshadowgi.so)
  host/   C++17 host side mirroring the reference's SceneLoader / Mesh / OBJ loader / per-technique
          render-pass interface (libshadowgi_host.so)
  capi.py / hostapi.py   ctypes views used by tests/ and bench.py
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
