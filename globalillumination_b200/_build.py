"""In-tree build of the native libraries (sm_100a only).

  libshadowgi.so       CUDA kernels + the C ABI of include/shadowgi.h   (nvcc, -fmad=false: see DESIGN.md §3)
  libshadowgi_host.so  C++17 host side: Configs/*.txt + OBJ loader, Mesh, matrices, render-pass interface (g++)

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
]


def _cxx():
    return "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else (shutil.which("g++") or "g++")


def _nvcc():
    for c in ("/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.access(c, os.X_OK):
            return c
    raise RuntimeError("nvcc not found: libshadowgi.so cannot be built")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force=False, verbose=False):
    out = os.path.join(PKG, "libshadowgi.so")
    srcs = _sources(CSRC, (".cu",))
    deps = srcs + _sources(CSRC, (".cuh",)) + [os.path.join(ROOT, "include", "shadowgi.h")]
    if force or _stale(out, deps):
        cmd = [_nvcc(), "-ccbin", _cxx()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out] + srcs
        subprocess.check_call(cmd)
    return out


def build_host(force=False):
    out = os.path.join(PKG, "libshadowgi_host.so")
    if not os.path.isdir(HOST):
        return None
    srcs = _sources(HOST, (".cpp",))
    if not srcs:
        return None
    deps = srcs + _sources(HOST, (".h",)) + [os.path.join(ROOT, "include", "shadowgi.h")]
    if force or _stale(out, deps):
        cmd = [_cxx(), "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-shared",
               "-I", os.path.join(ROOT, "include"), "-o", out] + srcs + ["-L", PKG, "-lshadowgi", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return out


def build_all(force=False):
    cuda = build_cuda(force)          # the host library links against it
    return cuda, build_host(force)


if __name__ == "__main__":
    print(build_all(force=True))
