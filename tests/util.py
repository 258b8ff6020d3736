"""Shared helpers for the tests: golden scenes, frame set-up (through the ORACLE's matrix code — the checker
side prepares identical inputs for the oracle and for the CUDA path), and comparison helpers."""
import os

import numpy as np

from oracle import oracle_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# input scenes (the reference loader's arrays for its config files): shipped with the package, shared with bench.py
SCENES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "globalillumination_b200", "data")
_cache = {}


def scene(name):
    if name not in _cache:
        z = np.load(os.path.join(SCENES, f"scene_{name}.npz"))
        _cache[name] = {k: z[k] for k in z.files}
    return _cache[name]


def golden(name):
    key = "golden:" + name
    if key not in _cache:
        z = np.load(os.path.join(GOLDEN, name))
        _cache[key] = {k: z[k] for k in z.files}
    return _cache[key]


def frame(sc, W, H, S, light_eye=None, light_at=None):
    return O.frame_matrices(sc["cam_eye"], sc["cam_at"], sc["light_eye"] if light_eye is None else light_eye,
                            sc["light_at"] if light_at is None else light_at, W, H, S, S)


def multi_lights(sc, n_lights, size, W, H, S):
    mvp, mvpb = [], []
    for i in range(n_lights):
        e = O.uniform_light_sample(sc["light_eye"], size, n_lights, i)
        a = O.uniform_light_sample(sc["light_at"], size, n_lights, i)
        fm = frame(sc, W, H, S, e, a)
        mvp.append(fm["light_mvp"]); mvpb.append(fm["light_mvp_b"])
    return np.stack(mvp), np.stack(mvpb)


def params_pair(tech, S, **kw):
    """(oracle params, C-ABI params) with identical fields."""
    from globalillumination_b200 import capi
    po = O.default_params(tech, S, **kw)
    pg = capi.default_params(tech, **{"shadow_map_width": S, "shadow_map_height": S, **kw})
    for f, _ in po._fields_:
        assert getattr(po, f) == getattr(pg, f), f
    return po, pg


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def describe_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    ne = a != b
    if a.dtype.kind == "f":
        ne &= ~(np.isnan(a) & np.isnan(b))
    n = int(ne.sum())
    if n == 0:
        return "identical"
    idx = np.argwhere(ne)[:5]
    return f"{n} of {a.size} differ ({100.0 * n / a.size:.4f}%), first at {idx.tolist()}: " + \
           ", ".join(f"{a[tuple(i)]!r} vs {b[tuple(i)]!r}" for i in idx)
