"""CPU tests of the oracle itself: against the golden vectors minted from the real reference
(tests/golden/make_golden.py) and, where oracle/_ref/libochref.so is present, against the real
reference live."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same_hits, same_bits

POSES = "ABC"


@pytest.mark.parametrize("name", ["d6_tunnels", "d8_tunnels"])
def test_oracle_trace_matches_reference_golden(oc, golden, name):
    g = golden(name)
    nodes8, root, depth = g["nodes8"], int(g["root"]), int(g["depth"])
    W, H = int(g["W"]), int(g["H"])
    for p in POSES:
        d = oc.gen_rays(g[f"pose{p}_rot"], float(g[f"pose{p}_fov"]), W, H)
        got = oc.trace_rays(nodes8, root, depth, g[f"pose{p}_pos"], d)
        assert_same_hits(got, (g[f"pose{p}_vox"], g[f"pose{p}_face"], g[f"pose{p}_t"]), f"{name} pose {p}")
    for k in ("rand", "edge"):
        got = oc.trace_rays(nodes8, root, depth, g[f"{k}_o"], g[f"{k}_d"])
        assert_same_hits(got, (g[f"{k}_vox"], g[f"{k}_face"], g[f"{k}_t"]), f"{name} {k}")


def test_oracle_table_mode_equals_hw_mode_on_this_host(oc, golden):
    """The table model of RCPSS must be exact on a host whose instruction it was derived from."""
    tab, bad = oc.rcp_table_from_hw(11)
    if bad:
        pytest.skip(f"this host's RCPSS is not an 11-bit table function ({bad} mismatches)")
    g = golden("d8_tunnels")
    for k in ("rand", "edge"):
        a = oc.trace_rays(g["nodes8"], int(g["root"]), 8, g[f"{k}_o"], g[f"{k}_d"])
        b = oc.trace_rays(g["nodes8"], int(g["root"]), 8, g[f"{k}_o"], g[f"{k}_d"], rcp_tab=tab)
        assert_same_hits(b, a, k)


def test_builtin_rcp_table_reproduces_golden(oc, golden):
    """The table embedded in the product (csrc/ort_rcp_table.h) fed to the oracle's table mode gives the
    reference's golden outputs on ANY host -- this is what makes GPU results host-independent."""
    tab = _builtin_table()
    g = golden("d8_tunnels")
    for k in ("rand", "edge"):
        got = oc.trace_rays(g["nodes8"], int(g["root"]), 8, g[f"{k}_o"], g[f"{k}_d"], rcp_tab=tab)
        assert_same_hits(got, (g[f"{k}_vox"], g[f"{k}_face"], g[f"{k}_t"]), k)


def _builtin_table():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "octree_ray_tracing_b200", "csrc", "ort_rcp_table.h")).read()
    vals = re.findall(r"0x([0-9a-f]{8})u", txt)
    assert len(vals) == 2048
    return np.array([int(v, 16) for v in vals], np.uint32)


def test_oracle_terrain_rebuilds_golden_dag(oc, golden):
    """heightmap + create_volume + surface set()s + tunnels, restated, give the reference's table."""
    g = golden("d6_tunnels")
    D, L = int(g["depth"]), int(g["log2cap"])
    h = oc.heightmap(D)
    assert np.array_equal(h, g["heights"])
    grass = np.unpackbits(g["grass"])[: h.size].reshape(h.shape)
    assert np.array_equal(grass, oc.grass_bits(D))
    T = oc.OracleTree(L, D)
    T.initialize_terrain(h, grass, True)
    assert (T.fillcnt, T.nodecnt) == (int(g["fillcnt"]), int(g["nodecnt"]))
    d = oc.gen_rays(g["poseB_rot"], float(g["poseB_fov"]), int(g["W"]), int(g["H"]))
    assert_same_hits(T.trace(g["poseB_pos"], d), (g["poseB_vox"], g["poseB_face"], g["poseB_t"]), "own table")


def test_digest_of_full_frames(oc):
    """1280x720 frames of the reference's default h_octree<19,8> demo tree: digests from the real reference."""
    from golden.make_golden import digest
    dig = json.load(open(os.path.join(GOLDEN, "digests.json")))["d8_tunnels_1280x720"]
    T = oc.OracleTree(19, 8)
    T.initialize_terrain(oc.heightmap(8), oc.grass_bits(8), True)
    pose = {"A": ((1.5, 1.5, 1.5), 0.0, 0.0), "B": ((1.5, 1.5, 1.5), 0.7, -0.6)}
    for p, (pos, yaw, pitch) in pose.items():
        rot, fov = oc.camera_coeffs(yaw, pitch)
        v, f, t = T.trace(np.array(pos, np.float32), oc.gen_rays(rot, fov, 1280, 720), nthreads=4)
        assert int((v != 0).sum()) == dig[p + "_hits"]
        assert f"{digest(v, f, t):016x}" == dig[p]


# ---- live against the real reference (authoring container, or wherever _ref travelled) -------

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libochref.so")),
                               reason="oracle/_ref/libochref.so not built")


@needs_ref
def test_live_noise_hash_morton(oc):
    rs = np.random.RandomState(0)
    R = oc.ref()
    xy = (rs.rand(100000, 2) * 8).astype(np.float32)
    b = np.zeros(len(xy), np.float32)
    R.ochref_simplex2(0.5, xy.ctypes.data, len(xy), b.ctypes.data)
    assert same_bits(oc.simplex2(0.5, xy), b)
    xyz = (rs.rand(100000, 3) * 64).astype(np.float32)
    b = np.zeros(len(xyz), np.float32)
    R.ochref_simplex3(0.5, xyz.ctypes.data, len(xyz), b.ctypes.data)
    assert same_bits(oc.simplex3(0.5, xyz), b)
    for _ in range(2000):
        c8 = rs.randint(0, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32)
        assert oc.lib().oc_node_hash(c8.ctypes.data) == R.ochref_node_hash(c8.ctypes.data)
        x, y, z = (int(v) for v in rs.randint(0, 65536, 3))
        assert oc.lib().oc_z_encode_16(x, y, z) == R.ochref_z_encode_16(x, y, z)


@needs_ref
def test_live_table_slot_for_slot(oc):
    """Random set()/unset sequences (incl. out-of-range coords and removals of absent voxels) leave the
    restated table byte-identical to the reference's: tags, refcounts, nodes, root, counters."""
    rs = np.random.RandomState(7)
    A, B = oc.OracleTree(16, 6), oc.RefTree(16, 6)
    for rnd in range(4):
        ops = np.concatenate([rs.randint(0, 70, (15000, 3)), rs.randint(0, 4, (15000, 1))], 1).astype(np.uint32)
        A.set_many(ops)
        B.set_many(ops)
        assert (A.root, A.fillcnt, A.nodecnt) == (B.root, B.fillcnt, B.nodecnt)
        assert np.array_equal(A.cashes(), B.cashes())
        live = (B.cashes() != 0) & (B.cashes() != 0xFF)
        assert np.array_equal(A.refcounts()[live], B.refcounts()[live])
        assert np.array_equal(A.nodes()[live], B.nodes()[live])
    pts = rs.randint(0, 64, (5000, 3))
    assert [A.at(*p) for p in pts] == [B.at(*p) for p in pts]


@needs_ref
def test_live_trace_random_trees(oc):
    """Random sparse trees, random + degenerate rays: restated trace == reference sse_trace, bitwise."""
    from golden.make_golden import edge_rays
    rs = np.random.RandomState(11)
    for L, D in ((12, 4), (16, 6)):
        A, B = oc.OracleTree(L, D), oc.RefTree(L, D)
        dim = 1 << D
        ops = np.concatenate([rs.randint(0, dim, (400 * D, 3)), rs.randint(1, 9, (400 * D, 1))], 1).astype(np.uint32)
        A.set_many(ops)
        B.set_many(ops)
        n = 30000
        o = rs.uniform(1.0 + 1e-3, 2.0 - 1e-3, (n, 3)).astype(np.float32)
        d = rs.normal(size=(n, 3)).astype(np.float32)
        assert_same_hits(A.trace(o, d), B.trace(o, d), f"random d{D}")
        o, d = edge_rays(rs, 500)
        assert_same_hits(A.trace(o, d), B.trace(o, d), f"edge d{D}")


def test_oracle_equals_the_real_reference_on_degenerate_rays(oc):
    """The ray classes the kernels special-case (tests/conftest.py degenerate_rays: +-0 / denormal / NaN / inf direction
    components, on-plane origins incl. exactly 1.0f, origins outside the cube, grazing rays) through the REAL
    och::h_octree::sse_trace (oracle/_ref) and the restatement: bitwise equal.  Live where the reference is built."""
    if not oc.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    import octree_ray_tracing_b200 as ort
    from conftest import degenerate_rays
    depth = 8
    T = ort.HOctree(19, depth, device=None)
    ort.harness.build_terrain(T, tunnels=True)
    nodes8, root, _ = T.flatten()
    R = oc.RefTree(19, depth)
    R.import_compact(nodes8, root)
    O, D = degenerate_rays(ort, depth)
    assert np.isnan(D).any() and np.isinf(D).any() and (O == 1.0).any()
    want = R.trace(O, D, nthreads=4)
    got = oc.trace_rays(nodes8, root, depth, O, D, nthreads=4)
    from conftest import assert_same_hits
    assert_same_hits(got, want, "oracle vs real reference, degenerate rays")
