"""Headless harness: replaces the olcPixelGameEngine frame loop and the demo's terrain set-up
(test_och_h_octree.cpp:504-556, :767-787) with scripted, deterministic equivalents."""
from __future__ import annotations

import os

import numpy as np

from ._lib import check, lib
from .tree import HOctree, _p

# camera poses used by SURVEY.md / BASELINE.md: (pos, yaw, pitch)
POSES = {
    "A": ((1.5, 1.5, 1.5), 0.0, 0.0),          # the demo's start pose (test_och_h_octree.cpp:53-55)
    "B": ((1.5, 1.5, 1.5), 0.7, -0.6),
    "C": ((1.1, 1.1, 1.4), 0.785, -0.3),
}


def heightmap(depth: int, nthreads: int | None = None) -> np.ndarray:
    """get_terrain_heigth over the whole map (test_och_h_octree.cpp:561-566, :587-592)."""
    dim = 1 << depth
    h = np.zeros((dim, dim), np.uint16)
    lib().ort_fixture_heightmap(depth, _p(h), nthreads or os.cpu_count() or 1)
    return h


def grass_bits(depth: int, seed: int = 1) -> np.ndarray:
    """One bit per column standing in for the unseeded std::rand() > RAND_MAX/2 (:780)."""
    dim = 1 << depth
    return np.random.RandomState(seed).randint(0, 2, size=(dim, dim)).astype(np.uint8)


def build_terrain(tree: HOctree, heights=None, grass=None, tunnels: bool = False, nthreads: int | None = None):
    """initialize_h_octree's voxel content (:767-787) through the memoising builder."""
    heights = heightmap(tree.depth, nthreads) if heights is None else np.ascontiguousarray(heights, np.uint16)
    grass = grass_bits(tree.depth) if grass is None else np.ascontiguousarray(grass, np.uint8)
    check(lib().ort_fixture_build_terrain(tree.h, _p(heights), _p(grass), int(tunnels), nthreads or os.cpu_count() or 1))
    return heights, grass


def random_rays(n: int, seed: int = 20261018):
    """Config 3 of BASELINE.json: incoherent rays.  Origins uniform in x,y in [1.05,1.95],
    z in [1.35,1.95] (above the terrain's maximum 1+5/16), directions uniform on the sphere."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    o = np.empty((n, 3), np.float32)
    o[:, 0] = rs.uniform(1.05, 1.95, n)
    o[:, 1] = rs.uniform(1.05, 1.95, n)
    o[:, 2] = rs.uniform(1.35, 1.95, n)
    d = rs.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


# The demo's voxel types and face colours (the reference ships them as voxels.txt:1-30; face order x+ y+ z+ x- y- z-,
# RRGGBB).  Voxel type ids are 1-based in this order (Stone = 1 ... Dirt = 4), as the terrain builder uses them.
DEMO_VOXELS = """
Stone:      44445D 4E4E5B 4E6155 2B352F 33333A 232328
Grass:      5D2917 3D260F 4F2E14 603718 6D2E0D 3F8527
Dark Grass: 5D2917 3D260F 4F2E14 603718 6D2E0D 317D1A
Dirt:       5D2917 3D260F 4F2E14 603718 6D2E0D 56220F
"""


def parse_voxels(text: str, max_voxels: int = 256):
    """(colours uint32[n,6] packed like olc::Pixel::n, names) from the reference's voxels.txt format."""
    import ctypes as C
    raw = text.encode()
    cols = np.zeros((max_voxels, 6), np.uint32)
    names = C.create_string_buffer(16 * max_voxels)
    n = lib().ort_parse_voxels(raw, len(raw), _p(cols), names, max_voxels)
    if n < 0:
        msg = lib().ort_last_error(None)
        raise ValueError(msg.decode() if msg else "malformed voxel file")
    return cols[:n].copy(), [names.raw[16 * i:16 * i + 16].split(b"\0")[0].decode() for i in range(n)]
