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


def heightmap(depth: int, nthreads: int | None = None, noise: str = "simplex") -> np.ndarray:
    """get_terrain_heigth over the whole map (test_och_h_octree.cpp:561-566, :587-592).  noise: "simplex" = the live code's
    och::simplex_n(0.5F) (:566), "opensimplex" = the commented alternative OpenSimplexNoise(8789) (:33, :568)."""
    dim = 1 << depth
    h = np.zeros((dim, dim), np.uint16)
    if noise == "opensimplex":
        lib().ort_fixture_heightmap_opensimplex(depth, OPENSIMPLEX_SEED, _p(h), nthreads or os.cpu_count() or 1)
    elif noise == "simplex":
        lib().ort_fixture_heightmap(depth, _p(h), nthreads or os.cpu_count() or 1)
    else:
        raise ValueError("noise must be 'simplex' or 'opensimplex'")
    return h


OPENSIMPLEX_SEED = 8789          # terrain_noise(8789), test_och_h_octree.cpp:33


def opensimplex2(xy, seed: int = OPENSIMPLEX_SEED) -> np.ndarray:
    """OpenSimplexNoise(seed).Evaluate(x, y) (opensimplex.h:338-386) for an (n, 2) float64 array."""
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    out = np.zeros(len(xy), np.float64)
    lib().ort_opensimplex2(seed, _p(xy), len(xy), _p(out))
    return out


def grass_bits(depth: int, seed: int = 1) -> np.ndarray:
    """One bit per column standing in for the unseeded std::rand() > RAND_MAX/2 (:780)."""
    dim = 1 << depth
    return np.random.RandomState(seed).randint(0, 2, size=(dim, dim)).astype(np.uint8)


def heightmap_gpu(ctx, depth: int) -> np.ndarray:
    """The same map from the CUDA kernel (bit-identical)."""
    dim = 1 << depth
    h = np.zeros((dim, dim), np.uint16)
    check(lib().ort_fixture_heightmap_gpu(ctx.h, depth, _p(h)))
    return h


def carve_bitmap_gpu(ctx, depth: int, heights: np.ndarray) -> np.ndarray:
    """Tunnel bitmap of remove(tree, splatter_noise(-0.5, .., 1/16)) (:735-743, :786) for z <= max height, from the CUDA
    kernel: uint64 words, bit (y*dim + x) of slab z."""
    dim = 1 << depth
    zmax = int(heights.max())
    words = (dim * dim + 63) // 64
    bits = np.zeros((zmax + 1, words), np.uint64)
    check(lib().ort_fixture_carve_gpu(ctx.h, depth, _p(np.ascontiguousarray(heights, np.uint16)), zmax, _p(bits)))
    return bits


def build_terrain(tree: HOctree, heights=None, grass=None, tunnels: bool = False, nthreads: int | None = None, gpu: bool | None = None, noise: str = "simplex"):
    """initialize_h_octree's voxel content (:767-787) through the memoising builder.  gpu: evaluate the noise (heightmap,
    tunnel bitmap) with the CUDA fixture kernels of the tree's context; default: whenever the tree has a context and the
    depth allows (>= 5).  The host threads do it otherwise -- same bits either way."""
    ctx = getattr(tree, "ctx", None)
    if gpu is None:
        gpu = ctx is not None and tree.depth >= 5
    if gpu and ctx is None:
        raise ValueError("build_terrain(gpu=True) needs a tree with a device context")
    if heights is None:
        heights = heightmap_gpu(ctx, tree.depth) if (gpu and noise == "simplex") else heightmap(tree.depth, nthreads, noise)
    else:
        heights = np.ascontiguousarray(heights, np.uint16)
    grass = grass_bits(tree.depth) if grass is None else np.ascontiguousarray(grass, np.uint8)
    carved = carve_bitmap_gpu(ctx, tree.depth, heights) if (gpu and tunnels) else None
    check(lib().ort_fixture_build_terrain_ex(tree.h, _p(heights), _p(grass), int(tunnels), nthreads or os.cpu_count() or 1, _p(carved)))
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


def save_png(path: str, rgba: np.ndarray, W: int, H: int):
    """Write a frame of ort_trace_frame_rgba pixels (olc::Pixel::n packing: r | g<<8 | b<<16 | a<<24) as an 8-bit RGBA
    PNG -- the headless stand-in for looking at the demo window.  Standard library only (zlib)."""
    import struct
    import zlib
    px = np.ascontiguousarray(rgba, np.uint32).reshape(H, W)
    rows = np.empty((H, 1 + 4 * W), np.uint8)
    rows[:, 0] = 0                                             # filter type 0 on every scanline
    rows[:, 1:] = px.view(np.uint8).reshape(H, 4 * W)          # little-endian uint32 -> bytes r, g, b, a

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 6, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(rows.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))
