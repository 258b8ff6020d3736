"""The C-ABI libraries load on a CPU-only box and export every symbol include/*.h declares; compute entry points
are not called here (no GPU), except to check that they fail loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

from globalillumination_b200 import capi, hostapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sg[ih]_[a-z0-9_]+)\s*\(", txt)))


def test_libshadowgi_exports_every_declared_symbol():
    lib = capi.load()
    names = declared("shadowgi.h")
    assert len(names) >= 20 and set(names) == set(capi.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_libshadowgi_host_exports_every_declared_symbol():
    lib = hostapi.load()
    names = declared("shadowgi_host.h")
    assert set(names) == set(hostapi.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_params_struct_layout_matches_header_and_oracle():
    from oracle import oracle_py as O
    assert C.sizeof(capi.SgiParams) == 25 * 4 == C.sizeof(O.Params)
    assert [f for f, _ in capi.SgiParams._fields_] == [f for f, _ in O.Params._fields_]
    p = capi.default_params("pcss")
    assert (p.shadow_intensity, p.kernel_order, p.penumbra_size, p.blocker_search_size, p.kernel_size, p.light_source_radius,
            p.max_search, p.z_near, p.z_far, p.polygon_offset_factor, p.polygon_offset_units, p.sv_infinity) == \
           (0.25, 7, 1, 7, 15, 8, 16, 1, 1000, 4.0, 20.0, 100)
    assert capi.TECH == O.TECH


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SgiError) as e:
        capi.Context(0)
    assert e.value.code == -5                                  # SGI_ERR_NO_DEVICE
    with pytest.raises(hostapi.HostError):
        hostapi.App(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "globalillumination_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="replace").read()
                # mentions in comments are fine; importing / dlopen-ing / including the checker is not
                assert "oracle_py" not in txt and "liboracle" not in txt and "oracle.h" not in txt, os.path.join(d, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), os.path.join(d, f)
