"""The device-side traversal code (csrc/ort_trace.cuh: camera rays, the RCPPS table model, the baseline walk and the
FastWalker / TightWalker / PipeWalker bookkeeping incl. the multi-level POP) compiled for the HOST by tests/host_emu and
held against the reference's golden outputs and the CPU oracle -- the same bar as the GPU parity tests (voxel, face and
hit time bit-exact, per-ray PUSH counts equal), but runnable where there is no GPU.  Every global load of the walk is
bounds-checked against the node array and the reciprocal table (cuda_shim.h; emu.py raises MemoryError on a stray load)
-- the memcheck this pool's GPUs do not offer.  Test infrastructure only: the
product library has no CPU path (tests/test_capi_symbols.py checks that)."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_same_hits, degenerate_rays

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))

WALKERS = (0, 1, 5, 7, 13, 14, 15)     # ort_set_option("variant") ids: traverse(), FastWalker, TightWalker, PipeWalker, LeanWalker tiers, FlatWalker, V4Walker
POSES = {"A": ((1.5, 1.5, 1.5), 0.0, 0.0), "B": ((1.5, 1.5, 1.5), 0.7, -0.6), "C": ((1.1, 1.1, 1.4), 0.785, -0.3)}


@pytest.fixture(scope="module")
def emu():
    import emu as m
    m.lib()
    return m


@pytest.mark.parametrize("name", ["d6_tunnels", "d8_tunnels"])
@pytest.mark.parametrize("walker", WALKERS)
def test_device_walkers_on_the_reference_golden_vectors(emu, golden, name, walker):
    """Expected outputs come from the unmodified reference (tests/golden/make_golden.py): frames of the three poses
    (in-kernel camera rays), random rays and the edge-case rays (axis-parallel, on-plane origins, +-0, tiny / huge
    components -- the ones that leave FastWalker's preconditions or take its one-level POP path)."""
    g = golden(name)
    depth, root = int(g["depth"]), int(g["root"])
    W, H = int(g["W"]), int(g["H"])
    for p in "ABC":
        got = emu.trace_frame(g["nodes8"], root, depth, g[f"pose{p}_pos"], g[f"pose{p}_rot"], float(g[f"pose{p}_fov"]), W, H, walker=walker)
        assert_same_hits(got, (g[f"pose{p}_vox"], g[f"pose{p}_face"], g[f"pose{p}_t"]), f"{name} frame {p}, walker {walker}")
    for k in ("rand", "edge"):
        got = emu.trace_rays(g["nodes8"], root, depth, g[f"{k}_o"], g[f"{k}_d"], walker=walker)
        assert_same_hits(got, (g[f"{k}_vox"], g[f"{k}_face"], g[f"{k}_t"]), f"{name} {k} rays, walker {walker}")


def test_device_walkers_vs_oracle_depth10_frames_and_push_counts(emu, ort, oc):
    """BASELINE config 1 shape (depth-10 terrain, poses A/B/C) at 640x360: every walker returns the oracle's hits and
    performs exactly the oracle's number of child-slot loads per ray -- the multi-level POP and the float position
    bookkeeping change how a round is computed, never which rounds happen."""
    depth = 10
    T = ort.HOctree(22, depth, device=None)
    ort.harness.build_terrain(T, tunnels=False)
    nodes8, root, _ = T.flatten()
    tab = emu.default_rcp_table()
    ncpu = max(1, min(16, os.cpu_count() or 1))
    W, H = 640, 360
    for p, (pos, yaw, pitch) in POSES.items():
        rot, fov = oc.camera_coeffs(yaw, pitch)
        d = oc.gen_rays(rot, fov, W, H)
        wv, wf, wt, wn, _ = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=ncpu, want_counts=True)
        assert (wv != 0).sum() > 1000
        for walker in WALKERS:
            got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, walker=walker, want_npush=True, want_stats=True)
            assert_same_hits(got, (wv, wf, wt), f"depth 10 pose {p}, walker {walker}")
            assert np.array_equal(got[3], wn), f"per-ray PUSH counts differ from the oracle's (pose {p}, walker {walker})"
            st = got[4]
            assert st["rays"] == W * H
            if walker == 1:
                assert int(st["rounds_by_level"].sum()) == int(wn.astype(np.int64).sum())
                assert st["slow_path_rays"] < W * H // 100      # camera rays take the fast path (bar a few axis-parallel ones)


def test_device_walkers_vs_oracle_incoherent_and_degenerate_rays(emu, ort, oc):
    """BASELINE config 3 shape on the depth-8 tunnel scene plus rays built to hit every special case of the fast path:
    zero and denormal direction components (coef = -inf), origins on cell planes, origins outside [1,2)^3, grazing rays."""
    depth = 8
    T = ort.HOctree(19, depth, device=None)
    ort.harness.build_terrain(T, tunnels=True)
    nodes8, root, _ = T.flatten()
    tab = emu.default_rcp_table()
    O, D = degenerate_rays(ort, depth)
    want = oc.trace_rays(nodes8, root, depth, O, D, rcp_tab=tab, nthreads=max(1, min(16, os.cpu_count() or 1)), want_counts=True)
    assert 0.05 < (want[0] != 0).mean() < 0.9
    for walker in WALKERS:
        got = emu.trace_rays(nodes8, root, depth, O, D, walker=walker, want_npush=True)
        assert_same_hits(got, want[:3], f"walker {walker}")
        assert np.array_equal(got[3], want[3]), f"walker {walker}: PUSH counts"


def test_device_walkers_with_the_camera_on_the_cube_boundary(emu, ort, oc):
    """A camera coordinate of exactly 1.0f: rays travelling in the positive direction on that axis start from the mirrored
    coordinate 2.0f, whose masked position bits are 0 (och_h_octree.h:314-320).  The reference walks that as raw bit
    patterns; fast_path_ok must send exactly those rays to the baseline walk (the others keep the fast path)."""
    depth = 8
    T = ort.HOctree(19, depth, device=None)
    ort.harness.build_terrain(T, tunnels=True)
    nodes8, root, _ = T.flatten()
    tab = emu.default_rcp_table()
    W, H = 320, 200
    for pos, yaw, pitch in [((1.0, 1.0, 1.0), 0.785, 0.6), ((1.0, 1.5, 1.75), 0.0, 0.0), ((1.5, 1.0, 1.5), 1.5708, -0.2), ((1.25, 1.5, 1.0), 0.3, 1.2)]:
        rot, fov = oc.camera_coeffs(yaw, pitch)
        d = oc.gen_rays(rot, fov, W, H)
        wv, wf, wt, wn, _ = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=4, want_counts=True)
        for walker in WALKERS:
            got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, walker=walker, want_npush=True, want_stats=True)
            assert_same_hits(got, (wv, wf, wt), f"camera at {pos}, walker {walker}")
            assert np.array_equal(got[3], wn), f"camera at {pos}, walker {walker}: PUSH counts"
            if walker:
                assert got[4]["slow_path_rays"] > 0, "rays travelling +axis from the 1.0f coordinate must leave the fast path"


def test_device_walkers_on_the_pool_octree_layout(emu, ort, oc):
    """och::octree: the same walkers over the pool layout -- raw row indices, root = row 0, MISS reports
    hit_time 0.0F (och_octree.cpp:302) -- against the oracle's restatement of och_octree.cpp:167-320."""
    from golden.make_golden import edge_rays
    rs = np.random.RandomState(8)
    depth, cap = 7, 1 << 16
    A, T = oc.OracleOctree(depth, cap), ort.Octree(depth, cap, device=None)
    ops = []
    for _ in range(40):
        c = rs.randint(8, 120, 3)
        e = rs.randint(2, 10)
        v = int(rs.randint(1, 6))
        ops += [(c[0] + x, c[1] + y, c[2] + z, v, 0) for x in range(e) for y in range(e) for z in range(e)]
    ops = np.array(ops, np.int32)
    A.apply(ops)
    T.apply(ops)
    tab = emu.default_rcp_table()
    n = 200_000
    o = rs.uniform(1.001, 1.999, (n, 3)).astype(np.float32)
    d = rs.normal(size=(n, 3)).astype(np.float32)
    eo, ed = edge_rays(rs, 300)
    O, D = np.concatenate([o, eo]), np.concatenate([d, ed])
    want = A.trace(O, D, rcp_tab=tab, nthreads=4)
    assert (want[0] != 0).sum() > 1000
    pool = np.array(T.nodes())
    for walker in WALKERS:
        got = emu.trace_rays(pool, 0, depth, O, D, walker=walker, miss_t=0.0, pool=True)
        assert_same_hits(got, want, f"pool layout, walker {walker}")
    miss = want[1] == 6
    assert miss.any() and (want[2][miss] == 0.0).all()


def test_device_row_mapping_of_cyclic_strips(emu, golden):
    """ort::frame_row (shift / mask for power-of-two tile heights, division otherwise) against the host-side partition
    (multi_gpu.strip_rows): the strips of all ranks reassemble the frame, for even and ragged tile counts."""
    from octree_ray_tracing_b200 import multi_gpu
    g = golden("d6_tunnels")
    depth, root = int(g["depth"]), int(g["root"])
    W, H = 96, 118                                       # 118 rows: the last tile is partial for most tile heights
    pos, rot, fov = g["poseC_pos"], g["poseC_rot"], float(g["poseC_fov"])
    full = [x.reshape(H, W) for x in emu.trace_frame(g["nodes8"], root, depth, pos, rot, fov, W, H)]
    assert (full[0] != 0).sum() > 500
    for tile_rows, world in ((8, 4), (8, 8), (16, 3), (1, 5), (5, 3), (12, 2), (32, 2), (7, 1)):
        seen = np.zeros(H, np.int32)
        for rank in range(world):
            y0, rows, frame_rows = multi_gpu.strip_rows(rank, world, H, tile_rows)
            seen[frame_rows] += 1
            # the kernels trace whole tiles; rows past the frame's end are clipped by the caller (rows counts only real ones)
            part = emu.trace_frame(g["nodes8"], root, depth, pos, rot, fov, W, H, y0=y0, rows=rows, tile_rows=tile_rows, tile_step=world)
            assert_same_hits(part, [f[frame_rows].ravel() for f in full], f"{tile_rows}-row tiles, rank {rank} of {world}")
        assert (seen == 1).all()


@pytest.mark.parametrize("depth,log2cap", [(1, 8), (2, 8), (16, 18)])
def test_device_walkers_at_extreme_depths(emu, ort, oc, depth, log2cap):
    """Depth 1 (one node) and depth 16 (a 16-entry parent stack, voxel-size 2^-16 steps, the multi-level POP across 15
    levels): the scene of the GPU test of the same name, through the emulated walkers.  tools/sanitize_host.sh runs this
    with ASan on the walkers' local stacks."""
    rs = np.random.RandomState(100 + depth)
    dim = 1 << depth
    T = ort.HOctree(log2cap, depth, device=None)
    n = 6 if depth <= 2 else 4000
    pts = rs.randint(0, dim, (n, 3))
    if depth == 16:
        pts[: n // 2] = 32768 + rs.randint(-40, 40, (n // 2, 3))
    pts[:2] = [[0, 0, 0], [dim - 1, dim - 1, dim - 1]]
    ops = np.concatenate([pts, rs.randint(1, 5, (n, 1))], 1).astype(np.uint32)
    T.set_many(ops)
    nodes8, root, _ = T.flatten()
    m = 30000
    o = rs.uniform(1.001, 1.999, (m, 3)).astype(np.float32)
    target = (1.0 + (pts[rs.randint(0, n, m)] + rs.uniform(0, 1, (m, 3))) / dim).astype(np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    d[:300, 1:] = 0.0
    d[300:600, 2] = 0.0
    o[600:900] = np.float32(1.5)
    o[900:1200] = np.float32(1.0)                             # on the cube's own faces: the baseline walk's domain
    tab = emu.default_rcp_table()
    want = oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=4, want_counts=True)
    assert (want[0] != 0).sum() > m // 10
    for walker in WALKERS:
        got = emu.trace_rays(nodes8, root, depth, o, d, walker=walker, want_npush=True)
        assert_same_hits(got, want, f"depth {depth}, walker {walker}")
        assert np.array_equal(got[3], want[3]), f"depth {depth}, walker {walker}: PUSH counts"


@pytest.mark.parametrize("depth,log2cap", [(12, 18), (16, 18)])
def test_lean_tier_split_with_origins_on_the_finest_grid(emu, ort, oc, depth, log2cap):
    """LeanWalker drops FastWalker's negative-t branch; the tier test (ort::lean_path_ok) must send every ray that can
    see a negative t to FastWalker.  A negative t needs the origin ON a cell plane of the finest level with a product
    o * coef that does not fit 24 bits (13+ significant origin bits x the 12-bit RCPPS result), i.e. depth >= 12: origins
    are snapped to the depth's grid (and, for some, to coarser grids), per axis and for all axes.  Both tiers must
    be populated and every ray must equal the oracle bit for bit, PUSH counts included."""
    rs = np.random.RandomState(7 + depth)
    dim = 1 << depth
    T = ort.HOctree(log2cap, depth, device=None)
    n = 3000
    pts = rs.randint(0, dim, (n, 3))
    pts[: n // 2] = dim // 2 + rs.randint(-60, 60, (n // 2, 3))
    T.set_many(np.concatenate([pts, rs.randint(1, 5, (n, 1))], 1).astype(np.uint32))
    nodes8, root, _ = T.flatten()
    m = 40000
    o = rs.uniform(1.001, 1.999, (m, 3))
    snap = rs.randint(0, 4, (m, 3))                          # per axis: 0 free, 1 finest grid, 2 a grid 3 levels up, 3 finest grid
    for k, q in ((1, dim), (2, dim >> 3), (3, dim)):
        sel = snap == k
        o[sel] = 1.0 + np.round((o[sel] - 1.0) * q) / q
    o = o.astype(np.float32)
    target = (1.0 + (pts[rs.randint(0, n, m)] + rs.uniform(0, 1, (m, 3))) / dim).astype(np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    tab = emu.default_rcp_table()
    want = oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=4, want_counts=True)
    assert (want[0] != 0).sum() > m // 20
    got = emu.trace_rays(nodes8, root, depth, o, d, walker=13, want_npush=True, want_stats=True)
    assert_same_hits(got, want, f"depth {depth}, tiered walk")
    assert np.array_equal(got[3], want[3]), f"depth {depth}: PUSH counts"
    st = got[4]
    assert 0 < st["lean_rays"] < st["rays"] - st["slow_path_rays"], "both tiers must have takers"
    # off-grid origins do not need the FastWalker tier (apart from the odd ray with a degenerate direction component or an origin the shift moved onto the grid)
    o2 = (o + np.float32(2.0 ** -22)).astype(np.float32)
    got2 = emu.trace_rays(nodes8, root, depth, o2, d, walker=13, want_stats=True)
    assert got2[3]["lean_rays"] >= 0.99 * (got2[3]["rays"] - got2[3]["slow_path_rays"])
    assert_same_hits(got2, oc.trace_rays(nodes8, root, depth, o2, d, rcp_tab=tab, nthreads=4), f"depth {depth}, off-grid origins")


def test_device_camera_rays_equal_the_oracle_rays(emu, oc):
    """ort::camera_ray (every operation rounded separately, IEEE sqrt and division) against the oracle's statement of
    tree_camera::update_position: traced through a single solid voxel so that the per-pixel direction decides t."""
    nodes8 = np.zeros((2, 8), np.uint32)
    nodes8[0, :] = 2                         # root: all eight children -> node 2
    nodes8[1, :] = 7                         # last level: all voxels solid, payload 7
    W, H = 161, 97
    for pos, yaw, pitch in [((2.5, 1.5, 1.5), 3.14159, 0.0), ((1.2, -0.5, 1.7), 1.3, 0.2), ((1.5, 1.5, 3.0), 0.3, -1.2)]:
        rot, fov = oc.camera_coeffs(yaw, pitch)
        d = oc.gen_rays(rot, fov, W, H)
        want = oc.trace_rays(nodes8, 1, 2, np.array(pos, np.float32), d, rcp_tab=emu.default_rcp_table())
        got = emu.trace_frame(nodes8, 1, 2, pos, rot, fov, W, H, walker=1)
        assert_same_hits(got, want, f"camera rays from {pos}")


def test_simt_model_accounts_for_every_round(emu, golden):
    """tools/simt_model.py's lockstep model (emu_warp_model): its lane-rounds are the PUSH counts of the frame, a warp runs
    at least as many rounds as its longest lane, and the three kinds of warp-rounds partition the total."""
    g = golden("d8_tunnels")
    depth, root = int(g["depth"]), int(g["root"])
    W, H = int(g["W"]), int(g["H"])
    pos, rot, fov = g["poseB_pos"], g["poseB_rot"], float(g["poseB_fov"])
    m = emu.warp_model(g["nodes8"], root, depth, pos, rot, fov, W, H)
    npush = emu.trace_frame(g["nodes8"], root, depth, pos, rot, fov, W, H, walker=1, want_npush=True)[3].astype(np.int64)
    assert m["lane_rounds"] == int(npush.sum()) == m["active_lanes"] == m["descend_lanes"] + m["advance_lanes"]
    assert m["warp_rounds"] == m["descend_only"] + m["advance_only"] + m["both"]
    tiles = npush.reshape(H, W)[: H // 4 * 4, : W // 8 * 8].reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
    if H % 4 == 0 and W % 8 == 0:
        assert m["warp_rounds"] == int(tiles.sum())          # lockstep: a warp runs as long as its longest lane
    assert m["warps"] == ((W + 7) // 8) * ((H + 3) // 4)
