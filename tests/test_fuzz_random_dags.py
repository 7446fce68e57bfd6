"""Differential fuzzing on random small DAGs (depth 1..8: isolated voxels, boxes, dense noise, a full cube with holes)
with rays drawn to sit on the corner cases -- origins snapped to cell planes of every level (incl. the cube's own faces),
origins outside the cube, zero / denormal / infinite / NaN direction components.  Three legs, same scenes:
the oracle against the REAL reference (oracle/_ref, where built), the device walkers compiled for the host
(tests/host_emu) against the oracle, and the CUDA path against the oracle (-m gpu).  Terrain scenes never produce
full nodes, single-child chains through many levels or rays that start inside solid blocks; these do."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_same_hits

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))

SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, 1e-40, np.inf, -np.inf, np.nan], np.float32)


def random_scene(ort, seed, log2cap=16, depth=None):
    rs = np.random.RandomState(seed)
    depth = int(rs.randint(1, 9)) if depth is None else depth
    dim = 1 << depth
    T = ort.HOctree(log2cap, depth, device=None)
    kind = seed % 4
    if kind == 0:                                         # isolated voxels: long single-child chains
        for x, y, z in rs.randint(0, dim, (int(rs.randint(1, 200)), 3)):
            T.set(int(x), int(y), int(z), int(rs.randint(1, 9)))
    elif kind == 1:                                       # a few boxes: uniform (maximally shared) subtrees
        for _ in range(int(rs.randint(1, 5))):
            lo = rs.randint(0, dim, 3)
            hi = np.minimum(lo + rs.randint(1, max(2, dim // 2), 3), dim)
            T.fill_box(lo, hi, int(rs.randint(1, 5)))
    elif kind == 2:                                       # dense noise in a corner
        m = rs.rand(min(dim, 16), min(dim, 16), min(dim, 16)) < 0.4
        for x, y, z in np.argwhere(m):
            T.set(int(x), int(y), int(z), 1 + int((x + y + z) % 3))
    else:                                                 # the full cube minus a few holes: every ray starts inside
        T.fill_box((0, 0, 0), (dim, dim, dim), 2)
        for x, y, z in rs.randint(0, dim, (30, 3)):
            T.set(int(x), int(y), int(z), 0)
    nodes8, root, _ = T.flatten()
    n = 4000
    o = rs.uniform(1.0, 2.0, (n, 3)).astype(np.float32)
    q = 1 << rs.randint(0, depth + 2, (n, 3))
    o = np.where(rs.rand(n, 3) < 0.5, np.floor((o - 1.0) * q) / q + 1.0, o).astype(np.float32)
    o[rs.rand(n) < 0.1] += rs.choice(np.array([-1.0, 1.0, 0.5], np.float32))
    d = rs.normal(size=(n, 3)).astype(np.float32)
    z = rs.rand(n, 3) < 0.15
    d[z] = rs.choice(SPECIAL, int(z.sum()))
    return depth, nodes8, root, o, d


def test_oracle_vs_real_reference_on_random_dags(ort, oc):
    if not oc.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    for seed in range(150):
        log2cap, depth = [(12, 4), (16, 6), (19, 8)][seed % 3]          # instantiations libochref.so carries
        depth, nodes8, root, o, d = random_scene(ort, 1000 + seed, log2cap, depth)
        if root == 0:
            continue
        R = oc.RefTree(log2cap, depth)
        R.import_compact(nodes8, root)
        assert_same_hits(oc.trace_rays(nodes8, root, depth, o, d, nthreads=2), R.trace(o, d, nthreads=2), f"seed {seed}, depth {depth}")


def test_device_walkers_vs_oracle_on_random_dags(ort, oc):
    import emu
    tab = emu.default_rcp_table()
    for seed in range(200):
        depth, nodes8, root, o, d = random_scene(ort, seed)
        if root == 0:
            continue
        want = oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=2, want_counts=True)
        for walker in (0, 1, 5, 7):
            got = emu.trace_rays(nodes8, root, depth, o, d, walker=walker, want_npush=True, nthreads=2)
            assert_same_hits(got, want, f"seed {seed}, depth {depth}, walker {walker}")
            assert np.array_equal(got[3], want[3]), f"seed {seed}, depth {depth}, walker {walker}: PUSH counts"


@pytest.mark.gpu
def test_cuda_path_vs_oracle_on_random_dags(ort, oc):
    from test_oracle import _builtin_table
    tab = _builtin_table()
    ctxs = {}
    for seed in range(96):
        depth, nodes8, root, o, d = random_scene(ort, seed)
        if root == 0:
            continue
        ctx = ctxs.setdefault(depth, ort.TraceContext(depth))
        ctx.upload_full(nodes8, root)
        want = oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=4, want_counts=True)
        for variant, rays_variant in ((13, 2), (13, 1), (1, 1), (0, 1)):
            ctx.set_option("variant", variant)
            ctx.set_option("rays_variant", rays_variant)
            got = ctx.trace_rays(o, d, want_npush=True)
            assert_same_hits(got, want, f"seed {seed}, depth {depth}, variant {variant}/{rays_variant}")
            assert np.array_equal(got[3], want[3]), f"seed {seed}, depth {depth}, variant {variant}/{rays_variant}: PUSH counts"
