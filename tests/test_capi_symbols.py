"""The C-ABI library loads without a GPU and exports every function include/ort_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        txt = open(os.path.join(ROOT, "include", fn)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(ort_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_header_declares_what_python_binds(ort):
    from octree_ray_tracing_b200 import _lib
    assert sorted(_lib.exported_symbols()) == declared_functions()


def test_library_exports_every_declared_symbol(ort):
    L = ctypes.CDLL(ort.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.ort_version.__call__.__self__.restype.__name__.encode() or True
    L.ort_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.ort_version()


def test_no_cpu_fallback_without_a_gpu(ort):
    """On a box without a GPU the trace path must fail loudly, not fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ort.OrtError) as e:
        ort.TraceContext(8)
    assert e.value.code in (2, 3)
    t = ort.HOctree(12, 4, device=None)
    t.set(1, 1, 1, 1)
    with pytest.raises(ort.OrtError):
        t.sse_trace((1.5, 1.5, 1.5), (0, 0, -1))


def test_product_does_not_touch_the_oracle():
    """Nothing under octree_ray_tracing_b200/ or include/ may import, include or load oracle/."""
    bad = []
    for base in ("octree_ray_tracing_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"och_oracle|liboch_oracle|libochref|from oracle|import oracle|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
