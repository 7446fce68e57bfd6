"""Mints tests/golden/opensimplex_8789.npz from the UNMODIFIED reference class OpenSimplexNoise (opensimplex.h, compiled
into oracle/_ref/libochref.so): sample points, their Evaluate(x, y) values for the demo's seed 8789 and another seed, and
the depth-6 heightmap of the commented get_terrain_heigth line (test_och_h_octree.cpp:568).  Run here (needs
/root/reference); the product's restatement (csrc/ort_opensimplex.h) is held to these bits on any box.
    python tests/golden/make_golden_opensimplex.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as oc  # noqa: E402


def ref_eval(seed, xy):
    R = oc.ref()
    xy = np.ascontiguousarray(xy, np.float64)
    out = np.zeros(len(xy), np.float64)
    R.ochref_opensimplex2(seed, xy.ctypes.data, len(xy), out.ctypes.data)
    return out


def main():
    assert oc.have_ref(), "needs oracle/_ref/libochref.so (the reference compiled in place)"
    rs = np.random.RandomState(8789)
    depth = 6
    dim = 1 << depth
    grid = np.stack(np.meshgrid(np.arange(dim), np.arange(dim), indexing="xy"), -1).reshape(-1, 2)
    pxy = (grid * 4).astype(np.float32) / np.float32(dim)                      # get_terrain_heigth's px, py (floats)
    xy = np.concatenate([rs.uniform(-40, 40, (3000, 2)), rs.uniform(0, 4, (3000, 2)), pxy.astype(np.float64)])
    v = ref_eval(8789, xy)
    v2 = ref_eval(-123456789, xy[:2000])
    h = (ref_eval(8789, pxy.astype(np.float64)) * dim / 16 + dim // 4).astype(np.int64).astype(np.uint16).reshape(dim, dim)   # static_cast<int>: truncation
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "opensimplex_8789.npz"), xy=xy, v8789=v, v_neg=v2, depth=depth, heights=h)
    print("written", len(xy), "points; heights", h.min(), h.max())


if __name__ == "__main__":
    main()
