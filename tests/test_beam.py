"""Beam start of camera frames (csrc/ort_beam.cuh) on the CPU: the device code (LeanWalker::start_at, lean_start, beam_march,
beam_tile_start, the host-side level choice) compiled for the host by tests/host_emu and held against the oracle.

Three claims are tested separately:
  1. re-entry: for ANY tau in (0, hit time] (cube exit time for a MISS, anything beyond it included) the walk re-entered
     at tau returns the oracle's voxel, face and hit time bit for bit;
  2. the bound: the tile start times the march yields never exceed the hit time of any ray of the tile;
  3. frames traced with the beam start equal the oracle, with fewer rounds, and the guard never fires.
Test infrastructure only (the product has no CPU path)."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_same_hits, degenerate_rays

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))

POSES = {"A": ((1.5, 1.5, 1.5), 0.0, 0.0), "B": ((1.5, 1.5, 1.5), 0.7, -0.6), "C": ((1.1, 1.1, 1.4), 0.785, -0.3)}
NCPU = max(1, min(16, os.cpu_count() or 1))


@pytest.fixture(scope="module")
def emu():
    import emu as m
    m.lib()
    return m


@pytest.fixture(scope="module")
def scene(ort):
    depth = 9
    T = ort.HOctree(21, depth, device=None)
    ort.harness.build_terrain(T, tunnels=True)
    nodes8, root, _ = T.flatten()
    return depth, nodes8, root


def test_reentry_at_any_time_up_to_the_hit_equals_the_oracle(emu, ort, oc, scene):
    """Claim 1.  Rays: random, degenerate (zero / denormal / NaN / inf components, on-plane and outside origins -- they
    leave the lean tier and must ignore tau) and camera rays.  tau: random fractions of the hit time, the hit time itself,
    the next float below it, other rays' hit times clipped to it, and for rays that MISS anything up to far beyond the cube."""
    depth, nodes8, root = scene
    tab = emu.default_rcp_table()
    O, D = degenerate_rays(ort, depth)
    rot, fov = oc.camera_coeffs(0.7, -0.6)
    dc = oc.gen_rays(rot, fov, 320, 180)
    O = np.concatenate([O, np.tile(np.array([[1.5, 1.5, 1.5]], np.float32), (len(dc), 1))])
    D = np.concatenate([D, dc])
    want = oc.trace_rays(nodes8, root, depth, O, D, rcp_tab=tab, nthreads=NCPU)
    t_hit = want[2].copy()
    rs = np.random.RandomState(5)
    hit = want[0] != 0
    miss = ~hit                                  # a MISS takes any tau; a hit at time 0 (origin inside a voxel) or NaN takes none: tau = 0
    finite = hit & np.isfinite(t_hit) & (t_hit > 0)
    taus = []
    for frac in (1.0, 0.999999, 0.5, None, None):
        f = rs.rand(len(D)).astype(np.float32) if frac is None else np.float32(frac)
        tau = np.where(finite, t_hit * f, np.where(miss, (rs.rand(len(D)) * 4.0).astype(np.float32), 0)).astype(np.float32)
        taus.append(tau)
    taus.append(np.where(finite, np.nextafter(t_hit, np.float32(0)), np.where(miss, np.float32(1e-30), 0)).astype(np.float32))
    taus.append(np.where(finite, np.minimum(t_hit, np.roll(np.where(finite, t_hit, 1.0), 1)), np.where(miss, np.float32(np.inf), 0)).astype(np.float32))
    beam_rays = 0
    for i, tau in enumerate(taus):
        got = emu.trace_rays(nodes8, root, depth, O, D, walker=13, tau=tau, want_stats=True, nthreads=NCPU)
        assert_same_hits(got, want, f"re-entry, tau set {i}")
        assert got[3]["beam_guard"] == 0, "a tau that is a lower bound of the hit time must never trip the guard"
        beam_rays += got[3]["beam_rays"] + got[3]["beam_misses"]
    assert beam_rays > 500_000, "the lean-tier rays must actually have taken the beam start"


def test_a_tau_beyond_the_hit_time_trips_the_guard_or_is_caught(emu, oc, scene):
    """The guard is no proof of anything, but where it fires it must repair: a tau INSIDE the hit voxel (later than the hit
    time, earlier than the voxel's exit) makes the re-entry land on the voxel without a STEP; the ray is walked again."""
    depth, nodes8, root = scene
    tab = emu.default_rcp_table()
    rot, fov = oc.camera_coeffs(0.7, -0.6)
    d = oc.gen_rays(rot, fov, 160, 90)
    o = np.array([1.5, 1.5, 1.5], np.float32)
    want = oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=NCPU)
    hit = want[0] != 0
    tau = np.where(hit, want[2] * np.float32(1.0 + 2.0 ** -14), 0).astype(np.float32)      # ~ a tenth of a voxel further
    got = emu.trace_rays(nodes8, root, depth, o, d, walker=13, tau=tau, want_stats=True, nthreads=NCPU)
    assert got[3]["beam_guard"] > 0
    fired_all_right = np.array_equal(got[0], want[0])          # rays landing in a solid voxel are repaired; the test scene has no one-voxel walls the others could skip
    assert fired_all_right


@pytest.mark.parametrize("depth,tunnels", [(6, True), (8, True), (10, False)])
def test_march_is_a_lower_bound_and_beam_frames_equal_the_oracle(emu, ort, oc, depth, tunnels):
    """Claims 2 and 3 on terrain scenes: the three bench poses, cameras in the corners of the cube, next to the terrain
    and inside tunnels, several frame sizes (the grid level follows the pixel size), whole frames and cyclic strips."""
    T = ort.HOctree(14 + depth, depth, device=None)
    ort.harness.build_terrain(T, tunnels=tunnels)
    nodes8, root, _ = T.flatten()
    tab = emu.default_rcp_table()
    cams = [POSES["A"], POSES["B"], POSES["C"],
            ((1.02, 1.03, 1.9), 0.785, -0.9), ((1.97, 1.96, 1.95), 3.9, -0.5), ((1.5, 1.5, 1.999), 0.3, -1.5),
            ((1.25, 1.75, 1.0 + 5.0 / 16.0 + 0.02), 1.1, -0.05), ((1.5, 1.5, 1.2), 2.0, 0.4), ((1.0625, 1.5, 1.5), 0.0, 0.0)]
    grids = {}
    seen_levels = set()
    total_ref = total_beam = tile_misses = 0
    for ci, (pos, yaw, pitch) in enumerate(cams):
        rot, fov = oc.camera_coeffs(yaw, pitch)
        for (W, H, y0, rows, tr, ts) in [(640, 360, 0, 360, 1, 1), (1920, 1080, 512, 40, 8, 2), (256, 144, 0, 144, 1, 1)][: 3 if ci < 4 else 1]:
            k = emu.beam_level(pos, rot, fov, W, H, depth)
            if k == 0:
                continue
            seen_levels.add(k)
            if k not in grids:
                grids[k] = emu.beam_grid(nodes8, root, k)
            # the oracle on the same rows
            frame_rows = [y0 + (r // tr) * tr * ts + r % tr for r in range(rows)] if ts > 1 else list(range(y0, y0 + rows))
            d = np.concatenate([oc.gen_rays(rot, fov, W, H, y, y + 1) for y in frame_rows])
            want = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=NCPU, want_counts=True)
            got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, y0=y0, rows=rows, tile_rows=tr, tile_step=ts, walker=13,
                                  want_npush=True, want_stats=True, beam=grids[k], want_tau=True)
            what = f"depth {depth}, camera {ci}, {W}x{H} rows {y0}+{rows} tiles {tr}/{ts}, level {k}"
            assert_same_hits(got, want, what)
            st, tau = got[4], got[5]
            assert st["beam_guard"] == 0, what
            assert st["beam_cert_wrong"] == 0, f"{what}: a tile was ended as a whole although some of its rays are not lean-tier rays"
            tile_misses += st["beam_tile_misses"]
            hit = want[0] != 0
            assert (tau[hit] <= want[2][hit]).all(), f"{what}: a tile start later than a hit time of the tile"
            assert (got[3] <= want[3]).all() or (got[3].astype(np.int64) - want[3]).max() <= depth, what
            total_ref += int(want[3].sum()); total_beam += int(got[3].sum())
    assert seen_levels, "no camera got a beam level"
    assert tile_misses > 10_000, "tiles that see nothing should end as a whole"
    assert total_beam < 0.8 * total_ref, f"the beam start should save rounds ({total_beam} vs {total_ref})"


def test_on_grid_origins_with_a_full_precision_reciprocal_table(emu, ort, oc):
    """With Intel's 12-bit reciprocals coef * o is exact for an origin of few bits, the t of the plane through the origin is +0
    and every ray is a lean-tier ray: tiles may end as a whole.  A table of correctly rounded reciprocals (24 significant
    bits, as ort_set_rcp_table accepts from another host) makes that t a rounding residue of either sign: rays change tier
    one by one, no tile may be certified, and the frames must still equal the oracle run with the same table."""
    depth = 8
    T = ort.HOctree(20, depth, device=None)
    ort.harness.build_terrain(T, tunnels=True)
    nodes8, root, _ = T.flatten()
    n = 1 << 11
    tab = (np.float32(1.0) / (np.float32(1.0) + (np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n))).astype(np.float32).view(np.uint32)
    W, H = 640, 360
    for pos in [(1.5, 1.5, 1.5), (1.0 + 77 / 256.0, 1.5, 1.75), (1.3, 1.6, 1.9)]:
        rot, fov = oc.camera_coeffs(0.7, -0.6)
        k = emu.beam_level(pos, rot, fov, W, H, depth, rcp_tab=tab)
        assert k > 0
        grid = emu.beam_grid(nodes8, root, k)
        d = oc.gen_rays(rot, fov, W, H)
        want = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=NCPU)
        got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, walker=13, rcp_tab=tab, want_stats=True, beam=grid)
        assert_same_hits(got, want, f"origin {pos}, 24-bit table")
        st = got[3]
        on_grid = all(float(c) * 256 == int(float(c) * 256) for c in pos) or any(float(c) * 256 == int(float(c) * 256) for c in pos)
        assert st["beam_cert_wrong"] == 0 and st["beam_guard"] == 0
        if on_grid:
            assert st["beam_tile_misses"] == 0, "an on-grid origin with inexact products must not certify tiles"
            assert st["lean_rays"] < st["rays"], "with rounding residues of both signs some rays leave the lean tier"
        else:
            assert st["beam_tile_misses"] > 0


def test_grid_properties_and_level_choice(emu, ort, oc, scene):
    depth, nodes8, root = scene
    for k in (3, 5, 7):
        g = emu.beam_grid(nodes8, root, k)
        n = 1 << k
        assert g.shape == (n, n, n) and g.max() <= k
        occ = g == 0
        # a cell with skip level j: its whole level-j cell is free of marked cells, and the level j-1 cell is not
        for j in range(1, k + 1):
            sel = np.argwhere(g == j)[:200]
            for z, y, x in sel:
                sh = k - j
                blk = occ[(z >> sh) << sh:((z >> sh) + 1) << sh, (y >> sh) << sh:((y >> sh) + 1) << sh, (x >> sh) << sh:((x >> sh) + 1) << sh]
                assert not blk.any()
                if j > 1:
                    sh += 1
                    blk = occ[(z >> sh) << sh:((z >> sh) + 1) << sh, (y >> sh) << sh:((y >> sh) + 1) << sh, (x >> sh) << sh:((x >> sh) + 1) << sh]
                    assert blk.any()
    # empty DAG region: the sky above the terrain is free at level 1 or 2
    assert (emu.beam_grid(nodes8, root, 5)[-1] != 0).all()
    # level choice: finer pixels -> finer grid; a camera in the corner sees longer rays -> coarser; no level for huge pixels,
    # for a matrix that is no rotation, for an origin outside the cube
    rot, fov = oc.camera_coeffs(0.7, -0.6)
    c = (1.5, 1.5, 1.5)
    assert emu.beam_level(c, rot, fov, 3840, 2160, 12) == 7
    assert emu.beam_level((1.1, 1.1, 1.4), rot, fov, 3840, 2160, 12) == 7
    corner = (1.02, 1.03, 1.97)
    assert emu.beam_level(corner, rot, fov, 1920, 1080, 12) == 6          # longer rays: the beam is wider at their far end
    lv = [emu.beam_level(corner, rot, fov, 240 << i, 135 << i, 12) for i in range(5)]
    assert lv == sorted(lv) and lv[0] in (0, 3) and lv[-1] == 7, lv
    assert emu.beam_level(c, rot, fov, 3840, 2160, 5) == 5
    assert emu.beam_level(c, rot, fov, 64, 36, 12) == 0
    assert emu.beam_level(c, rot * 1.01, fov, 3840, 2160, 12) == 0
    assert emu.beam_level((2.5, 1.5, 1.5), rot, fov, 3840, 2160, 12) == 0
    assert emu.beam_level(c, rot, 0.0, 3840, 2160, 12) == 0


def test_beam_on_the_pool_octree_layout(emu, ort, oc):
    """och::octree pool rows (root = row 0, MISS time 0.0F) take the same beam start."""
    from oracle import oracle as ocm
    depth = 7
    P = ocm.OracleOctree(depth, 1 << 16)
    rs = np.random.RandomState(3)
    for _ in range(300):
        x, y, z = rs.randint(0, 1 << depth, 3)
        P.set(int(x), int(y), int(z) // 3, int(rs.randint(1, 5)))
    pool = P.nodes().copy()
    pos, rot_fov = (1.5, 1.5, 1.8), oc.camera_coeffs(0.4, -0.8)
    rot, fov = rot_fov
    W, H = 640, 360
    k = emu.beam_level(pos, rot, fov, W, H, depth)
    assert k > 0
    grid = emu.beam_grid(pool, 0, k, pool=True)
    want = emu.trace_frame(pool, 0, depth, pos, rot, fov, W, H, walker=0, miss_t=0.0, pool=True)
    got = emu.trace_frame(pool, 0, depth, pos, rot, fov, W, H, walker=13, miss_t=0.0, pool=True, beam=grid, want_stats=True)
    assert_same_hits(got, want, "pool layout with beam start")
    assert got[3]["beam_rays"] + got[3]["beam_misses"] > 0 and got[3]["beam_guard"] == 0


def test_beam_frames_on_random_dags(emu, ort, oc):
    """The scenes of tests/test_fuzz_random_dags.py -- isolated voxels (the thing a non-conservative beam would fly past),
    boxes, dense noise, a solid cube with holes (every ray starts inside: no beam start may happen) -- seen from random
    cameras, on-grid origins included, at frame sizes that select every grid level: frames with the beam start equal the
    oracle, the tile starts never exceed a hit time, the guard never fires, no tile is certified wrongly."""
    from test_fuzz_random_dags import random_scene
    tab = emu.default_rcp_table()
    levels = set()
    frames = tile_misses = beam_rays = 0
    for seed in range(120):
        depth, nodes8, root, _, _ = random_scene(ort, 5000 + seed, depth=3 + seed % 6)
        if root == 0:
            continue
        rs = np.random.RandomState(seed)
        grids = {}
        for cam in range(3):
            pos = rs.uniform(1.02, 1.98, 3).astype(np.float32)
            if cam == 1:
                q = 1 << int(rs.randint(1, depth + 1))
                pos = (np.floor((pos - 1.0) * q) / q + 1.0 + (1.0 / q if rs.rand() < 0.5 else 0.0)).astype(np.float32)      # on some level's grid
                pos = np.clip(pos, np.float32(1.0) + np.float32(1.0 / q), np.float32(2.0) - np.float32(1.0 / q))
            rot, fov = oc.camera_coeffs(float(rs.uniform(-3.1, 3.1)), float(rs.uniform(-1.5, 1.5)))
            W, H = [(256, 144), (640, 360), (1600, 900)][(seed + cam) % 3]
            rows = min(H, 64)
            y0 = int(rs.randint(0, H - rows + 1)) & ~3
            k = emu.beam_level(pos, rot, fov, W, H, depth)
            if k == 0:
                continue
            levels.add(k)
            if k not in grids:
                grids[k] = emu.beam_grid(nodes8, root, k)
            d = oc.gen_rays(rot, fov, W, H, y0, y0 + rows)
            want = oc.trace_rays(nodes8, root, depth, pos, d, rcp_tab=tab, nthreads=NCPU)
            got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, y0=y0, rows=rows, walker=13, want_stats=True, beam=grids[k], want_tau=True)
            what = f"seed {seed} (kind {(5000 + seed) % 4}, depth {depth}), camera {cam} at {pos.tolist()}, {W}x{H}, level {k}"
            assert_same_hits(got, want, what)
            st, tau = got[3], got[4]
            assert st["beam_guard"] == 0 and st["beam_cert_wrong"] == 0, what
            hit = want[0] != 0
            assert (tau[hit] <= want[2][hit]).all(), f"{what}: a tile start later than a hit time of the tile"
            frames += 1; tile_misses += st["beam_tile_misses"]; beam_rays += st["beam_rays"]
    assert frames > 150 and len(levels) >= 3, (frames, levels)
    assert tile_misses > 0 and beam_rays > 0


def test_needles_and_sheets_are_not_flown_past(emu, ort, oc):
    """What a beam bound must never do: start rays behind something thinner than the beam.  Depth 12, 4K pixels (grid level 7,
    cells 32 voxels wide): single voxels floating in empty space at several distances, a wall one voxel thick and a
    one-voxel-wide rod along the view direction, each seen by a handful of pixels at most."""
    depth = 12
    dim = 1 << depth
    T = ort.HOctree(16, depth, device=None)
    rs = np.random.RandomState(9)
    needles = [(2100 + 37 * i, 2048 + int(rs.randint(-300, 300)), 2048 + int(rs.randint(-200, 200))) for i in range(40)]
    for x, y, z in needles:
        T.set(x, y, z, 1 + (x % 5))
    T.fill_box((3500, 1000, 1000), (3501, 3000, 3000), 3)            # a sheet, one voxel thick, far away
    T.fill_box((2300, 2047, 1500), (3400, 2048, 1501), 4)            # a rod along x
    nodes8, root, _ = T.flatten()
    tab = emu.default_rcp_table()
    W, H = 3840, 2160
    pos = np.array([1.5, 1.5, 1.5], np.float32)                      # voxel (2048, 2048, 2048): on every grid
    for yaw, pitch in [(0.0, 0.0), (0.03, -0.12), (-0.05, 0.04)]:
        rot, fov = oc.camera_coeffs(yaw, pitch)
        k = emu.beam_level(pos, rot, fov, W, H, depth)
        assert k == 7
        grid = emu.beam_grid(nodes8, root, k)
        y0, rows = 760, 640
        d = oc.gen_rays(rot, fov, W, H, y0, y0 + rows)
        want = oc.trace_rays(nodes8, root, depth, pos, d, rcp_tab=tab, nthreads=NCPU)
        got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, y0=y0, rows=rows, walker=13, want_stats=True, beam=grid, want_tau=True)
        assert_same_hits(got, want, f"needles, yaw {yaw} pitch {pitch}")
        hit = want[0] != 0
        assert hit.sum() > 100 and len(np.unique(want[0][hit])) >= 4, "the scene should be visible: needles of several kinds, the sheet, the rod"
        assert (got[4][hit] <= want[2][hit]).all()
        assert got[3]["beam_guard"] == 0 and got[3]["beam_cert_wrong"] == 0 and got[3]["beam_tile_misses"] > 0
