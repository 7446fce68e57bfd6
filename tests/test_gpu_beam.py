"""Beam start of camera frames on the GPU (csrc/ort_beam.cuh, the beam_* kernels and the BEAM instantiations of the frame
kernels), through the C ABI.  The claims the CPU suite holds the host emulation to (tests/test_beam.py) are repeated here
on the device: same grid bytes, same outputs with the option on and off, equal to the reference / the oracle, fewer
rounds, grids rebuilt after edits."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_same_hits

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))

pytestmark = pytest.mark.gpu

POSES = {"A": ((1.5, 1.5, 1.5), 0.0, 0.0), "B": ((1.5, 1.5, 1.5), 0.7, -0.6), "C": ((1.1, 1.1, 1.4), 0.785, -0.3)}
NCPU = max(1, min(32, os.cpu_count() or 1))


@pytest.fixture(scope="module")
def emu():
    import emu as m
    m.lib()
    return m


def test_gpu_grid_equals_host_grid_and_follows_edits(ort, emu):
    depth = 9
    T = ort.HOctree(21, depth, device=0)
    ort.harness.build_terrain(T, tunnels=True)
    T.sync()
    nodes8, root, _ = T.flatten()
    for k in (3, 6, 7):
        assert np.array_equal(T.ctx.beam_grid(k), emu.beam_grid(nodes8, root, k)), f"level {k}"
    b0 = T.ctx.beam_builds
    assert np.array_equal(T.ctx.beam_grid(6), emu.beam_grid(nodes8, root, 6)) and T.ctx.beam_builds == b0, "an unchanged DAG keeps its grid"
    # a floating block high above the terrain: cells that were free at level 1 are marked now
    T.set_box(300, 300, 470, 12, 3)
    T.sync()
    nodes8, root, _ = T.flatten()
    for k in (6, 7, 3):
        g = T.ctx.beam_grid(k)
        assert np.array_equal(g, emu.beam_grid(nodes8, root, k)), f"level {k} after the edit"
        assert g[(470 * (1 << k)) >> depth, (300 * (1 << k)) >> depth, (300 * (1 << k)) >> depth] == 0
    assert T.ctx.beam_builds == b0 + 3
    T.ctx.close()


def test_bench_frames_beam_on_equals_beam_off_and_the_reference(ort, oc, emu):
    """Depth 12, 3840x2160, poses A/B/C: every ray with the beam start == every ray without, sampled rows == the CPU
    checker (the reference's own sse_trace where it travelled), and the rounds actually issued are the host emulation's."""
    depth = 12
    T = ort.HOctree(24, depth, device=0)
    ort.harness.build_terrain(T)
    T.sync()
    ctx = T.ctx
    ctx.set_option("beam_after", 0)          # a grid at the first frame of a DAG version (default: after two frames without)
    nodes8, root, _ = T.flatten()
    tab = emu.default_rcp_table()
    W, H = 3840, 2160
    if oc.have_ref():
        R = oc.RefTree(24, depth)
        R.import_compact(nodes8, root)
        cpu = lambda o, d: R.trace(o, d, nthreads=NCPU)
    else:
        cpu = lambda o, d: oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=NCPU)
    tot_on = tot_off = 0
    for name, (pos, yaw, pitch) in POSES.items():
        rot, fov = oc.camera_coeffs(yaw, pitch)
        assert ctx.beam_level(pos, rot, fov, W, H) == 7
        ctx.set_option("beam", 0)
        off = ctx.trace_frame(pos, rot, fov, W, H, want_npush=True)
        ctx.set_option("beam", 1)
        l0 = ctx.launch_count
        on = ctx.trace_frame(pos, rot, fov, W, H)
        assert ctx.launch_count - l0 >= 2 * 8, "every chunk of the host-buffer frame is a march + a trace launch"
        assert_same_hits(on, off, f"pose {name}: beam on vs off")
        ys = np.arange(5, H, 54)
        d = np.concatenate([oc.gen_rays(rot, fov, W, H, int(y), int(y) + 1) for y in ys])
        want = cpu(np.array(pos, np.float32), d)
        sel = (ys[:, None] * W + np.arange(W)[None, :]).ravel()
        assert_same_hits(tuple(a[sel] for a in on), want, f"pose {name}: beam on vs the CPU checker")
        # rounds: counted with the beam start, against the emulation of the same rows
        ctx.set_option("count_beam", 1)
        cnt = ctx.trace_frame(pos, rot, fov, W, H, want_npush=True)
        ctx.set_option("count_beam", 0)
        assert_same_hits(cnt, off, f"pose {name}: counting launch with beam")
        tot_on += int(cnt[3].astype(np.int64).sum()); tot_off += int(off[3].astype(np.int64).sum())
        if name == "C":
            grid = emu.beam_grid(nodes8, root, 7)
            e = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, y0=1024, rows=64, walker=13, want_npush=True, beam=grid)
            assert np.array_equal(e[3], cnt[3][1024 * W:(1024 + 64) * W]), "per-ray rounds with the beam start differ from the host emulation's"
    assert tot_on < 0.6 * tot_off, (tot_on, tot_off)
    ctx.close()


def test_strips_rgba_batched_launches_and_small_frames(ort, oc, emu):
    """Every frame entry point with the beam start against itself without: cyclic strips (tile heights 8 and 12; 6 is not
    a multiple of 4 and runs without), shaded frames, the batched launch, device buffers, a depth-7 DAG (grid level =
    depth), frames too coarse for any grid."""
    import torch
    depth = 10
    T = ort.HOctree(22, depth, device=0)
    ort.harness.build_terrain(T, tunnels=True)
    T.sync()
    ctx = T.ctx
    ctx.set_option("beam_after", 0)
    cols, _ = ort.harness.parse_voxels(ort.harness.DEMO_VOXELS)
    ctx.set_palette(cols)
    W, H = 1920, 1080
    cams = [(np.array(p, np.float32),) + tuple(oc.camera_coeffs(y, pt)) for p, y, pt in POSES.values()]
    cams.append((np.array([1.03, 1.96, 1.93], np.float32),) + tuple(oc.camera_coeffs(5.5, -0.7)))

    def run(fn):
        ctx.set_option("beam", 0)
        a = fn()
        ctx.set_option("beam", 1)
        b = fn()
        return a, b

    for ci, (pos, rot, fov) in enumerate(cams):
        for (y0, rows, tr, ts) in [(0, H, 1, 1), (8, 536, 8, 2), (24, 360, 12, 3), (6, 180, 6, 5), (17, 333, 1, 1)]:
            a, b = run(lambda: ctx.trace_frame(pos, rot, fov, W, H, y0=y0, rows=rows, tile_rows=tr, tile_step=ts))
            assert_same_hits(b, a, f"camera {ci}, rows {y0}+{rows}, tiles {tr}/{ts}")
        a, b = run(lambda: ctx.trace_frame_rgba(pos, rot, fov, W, H))
        assert np.array_equal(a, b), f"camera {ci}: shaded frame"
        a, b = run(lambda: ctx.trace_frame_rgba(pos, rot, fov, W, H, y0=16, rows=256, tile_rows=16, tile_step=4))
        assert np.array_equal(a, b), f"camera {ci}: shaded strip"
        a, b = run(lambda: ctx.trace_frame(pos, rot, fov, 96, 54))
        assert_same_hits(b, a, f"camera {ci}: 96x54 (no grid is coarse enough)")
    assert ctx.beam_level(cams[0][0], cams[0][1], cams[0][2], 96, 54) == 0

    # batched launch on device buffers: strips of all cameras in one launch, twice (beam off / on)
    def batched():
        outs, jobs = [], []
        for k, (pos, rot, fov) in enumerate(cams):
            rows = 8 * (((H // 8) - k % 2 + 1) // 2)
            n = rows * W
            o = (torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.float32, device="cuda"))
            outs.append(o)
            jobs.append((pos, rot, fov, W, H, 8 * (k % 2), rows, 8, 2, o[0], o[1], o[2]))
        ctx.trace_frames_async(jobs)
        ctx.sync()
        return [(o[0].cpu().numpy().view(np.uint32), o[1].cpu().numpy(), o[2].cpu().numpy()) for o in outs]
    a, b = run(batched)
    for k in range(len(cams)):
        assert_same_hits(b[k], a[k], f"batched job {k}")
    ctx.close()

    depth = 7
    T = ort.HOctree(18, depth, device=0)
    ort.harness.build_terrain(T, tunnels=True)
    T.sync()
    pos, rot, fov = cams[1]
    assert T.ctx.beam_level(pos, rot, fov, 3840, 2160) == 7
    T.ctx.set_option("beam_after", 0)
    T.ctx.set_option("beam", 0)
    a = T.ctx.trace_frame(pos, rot, fov, 3840, 2160)
    T.ctx.set_option("beam", 1)
    b = T.ctx.trace_frame(pos, rot, fov, 3840, 2160)
    assert_same_hits(b, a, "depth 7, grid level = depth")
    T.ctx.close()


def test_edit_loop_with_beam_start_vs_oracle(ort, oc, emu):
    """Config 4's loop at 1080p: edit bursts (add above the terrain, carve below), delta upload, frame -- the grid must follow
    every DAG version (a stale grid would start rays behind the new voxels)."""
    depth = 8
    T = ort.HOctree(20, depth, device=0)
    ort.harness.build_terrain(T, tunnels=False)
    tab = emu.default_rcp_table()
    pos, yaw, pitch = POSES["B"]
    rot, fov = oc.camera_coeffs(yaw, pitch)
    W, H = 1920, 1080
    assert T.ctx.beam_level(pos, rot, fov, W, H) > 0
    T.ctx.set_option("beam_after", 0)
    d = oc.gen_rays(rot, fov, W, H)
    rs = np.random.RandomState(2)
    builds = T.ctx.beam_builds
    for it in range(6):
        x, y = (int(v) for v in rs.randint(40, 200, 2))
        if it % 2 == 0:
            T.set_box(x, y, 150 + 10 * it, 14, 2)          # a block in the sky, between the camera and the terrain
        else:
            T.set_box(x, y, 60, 30, 0)                     # a pit
        T.sync()
        got = T.ctx.trace_frame(pos, rot, fov, W, H)
        nodes8, root, _ = T.flatten()
        want = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=NCPU)
        assert_same_hits(got, want, f"edit {it}")
        assert T.ctx.beam_builds == builds + it + 1
    T.ctx.close()


def test_band_schedule_changes_no_output(ort, oc):
    """The band schedule of a view (recorded costs -> most expensive 16-row bands first) is a permutation of the launch's
    blocks: frames traced before a schedule exists, while one is being recorded and after it is applied are the same,
    for whole frames and strips, with and without the beam start, and other views keep their own schedules."""
    import torch
    depth = 10
    T = ort.HOctree(22, depth, device=0)
    ort.harness.build_terrain(T, tunnels=True)
    T.sync()
    ctx = T.ctx
    W, H = 1920, 1080
    views = [(np.array(p, np.float32),) + tuple(oc.camera_coeffs(y, pt)) for p, y, pt in POSES.values()]
    n = W * H
    dv = torch.empty(n, dtype=torch.int32, device="cuda"); df = torch.empty(n, dtype=torch.uint8, device="cuda"); dt = torch.empty(n, dtype=torch.float32, device="cuda")

    def frame(pos, rot, fov, y0=0, rows=H, tr=1, ts=1):
        m = rows * W
        ctx.trace_frame_async(pos, rot, fov, W, H, y0, rows, tr, ts, dv, df, dt)
        ctx.sync()
        torch.cuda.synchronize()
        return dv[:m].cpu().numpy().view(np.uint32), df[:m].cpu().numpy(), dt[:m].cpu().numpy()

    ctx.set_option("band_order", 0)
    want = [frame(*v) for v in views]
    want_strip = [frame(*v, y0=8, rows=536, tr=8, ts=2) for v in views]
    assert ctx.band_schedules == 0
    ctx.set_option("band_order", 1)
    for rep in range(5):
        for k, v in enumerate(views):
            assert_same_hits(frame(*v), want[k], f"view {k}, repetition {rep}")
            assert_same_hits(frame(*v, y0=8, rows=536, tr=8, ts=2), want_strip[k], f"strip of view {k}, repetition {rep}")
    assert ctx.band_schedules >= 2 * len(views), "every repeated view should have had schedules applied"
    s0 = ctx.band_schedules
    ctx.set_option("beam", 0)
    for rep in range(3):
        for k, v in enumerate(views):
            assert_same_hits(frame(*v), want[k], f"view {k} without beam, repetition {rep}")
    # a camera that never repeats a view records nothing and applies nothing
    s1 = ctx.band_schedules
    for i in range(70):
        rot, fov = oc.camera_coeffs(0.01 * i, -0.5)
        frame(np.array([1.5, 1.5, 1.6], np.float32), rot, fov)
    assert ctx.band_schedules == s1
    assert s1 >= s0
    ctx.close()


def test_beam_experiment_walkers_equal_the_product(ort, oc):
    """Measurement build only (ORT_B200_EXPERIMENTS=1; tests/test_experiments.py runs this in a child process): the walkers
    and loop shapes of csrc/ort_experiments.cuh with the beam start (variants 24-28, 30, 31; 29 runs without the guard and is
    a measurement only) against the product kernel."""
    if b"experiments" not in ort.lib().ort_version():
        pytest.skip("the product library carries no experiment kernels")
    depth = 10
    T = ort.HOctree(22, depth, device=0)
    ort.harness.build_terrain(T, tunnels=True)
    T.sync()
    ctx = T.ctx
    ctx.set_option("beam_after", 0)
    W, H = 1920, 1080
    for name, (pos, yaw, pitch) in POSES.items():
        rot, fov = oc.camera_coeffs(yaw, pitch)
        ctx.set_option("variant", 13)
        want = ctx.trace_frame(pos, rot, fov, W, H)
        ctx.set_option("count_beam", 1)
        want_n = ctx.trace_frame(pos, rot, fov, W, H, want_npush=True)[3]
        for v in (24, 25, 26, 27, 28, 30, 31):
            ctx.set_option("variant", v)
            assert_same_hits(ctx.trace_frame(pos, rot, fov, W, H), want, f"pose {name}, variant {v}")
            got = ctx.trace_frame(pos, rot, fov, W, H, y0=8, rows=536, tile_rows=8, tile_step=2, want_npush=True)
            ctx.set_option("variant", 13)
            ref = ctx.trace_frame(pos, rot, fov, W, H, y0=8, rows=536, tile_rows=8, tile_step=2, want_npush=True)
            assert_same_hits(got, ref, f"pose {name}, variant {v}, strip")
            assert np.array_equal(got[3], ref[3]), f"pose {name}, variant {v}: rounds differ from the product's"
        ctx.set_option("count_beam", 0)
        assert want_n.astype(np.int64).sum() > 0
    ctx.close()


def test_an_edit_loop_never_builds_a_grid_and_a_steady_scene_does(ort, oc, emu):
    """Default policy ("beam_after" = 2): a DAG version gets its grid at the third frame call that meets it.  Edits every
    other frame (BASELINE config 4) therefore never build one; frames of an unchanged DAG do from the third on; either
    way every frame equals the oracle."""
    depth = 8
    T = ort.HOctree(20, depth, device=0)
    ort.harness.build_terrain(T, tunnels=False)
    tab = emu.default_rcp_table()
    pos, yaw, pitch = POSES["B"]
    rot, fov = oc.camera_coeffs(yaw, pitch)
    W, H = 1920, 1080
    d = oc.gen_rays(rot, fov, W, H)

    def check(what):
        got = T.ctx.trace_frame(pos, rot, fov, W, H)            # a host-buffer frame: one call, several chunk launches
        nodes8, root, _ = T.flatten()
        assert_same_hits(got, oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=NCPU), what)

    T.sync()
    for it in range(6):
        T.set_box(60 + 20 * it, 80, 150 + 10 * (it % 3), 14, 2)
        T.sync()
        check(f"edit {it}, first frame")
        check(f"edit {it}, second frame")
    assert T.ctx.beam_builds == 0, "an edit every other frame must not build grids"
    for k in range(4):
        check(f"steady frame {k}")
    assert T.ctx.beam_builds == 1
    T.ctx.close()


def test_pool_octree_frames_with_beam_start_vs_oracle(ort, oc):
    """och::octree pool layout (raw rows, root = row 0, MISS time 0.0F) with the beam start, grids following the row deltas;
    per-frame syncs without edits must not throw the grid away."""
    from test_oracle import _builtin_table
    rs = np.random.RandomState(11)
    depth, cap = 7, 1 << 16
    A, T = oc.OracleOctree(depth, cap), ort.Octree(depth, cap)
    ops = []
    for _ in range(30):
        c = rs.randint(8, 120, 3)
        e = rs.randint(2, 10)
        ops += [(c[0] + x, c[1] + y, (c[2] + z) // 2, int(rs.randint(1, 6)), 0) for x in range(e) for y in range(e) for z in range(e)]
    ops = np.array(ops, np.int32)
    A.apply(ops); T.apply(ops)
    T.sync()
    T.ctx.set_option("beam_after", 0)
    tab = _builtin_table()
    pos, yaw, pitch = (1.5, 1.5, 1.9), 0.3, -1.1
    rot, fov = oc.camera_coeffs(yaw, pitch)
    W, H = 1280, 720
    assert T.ctx.beam_level(pos, rot, fov, W, H) > 0
    d = oc.gen_rays(rot, fov, W, H)
    for step in range(3):
        got = T.trace_frame(pos, yaw, pitch, W, H)
        want = A.trace(np.array(pos, np.float32), d, rcp_tab=tab, nthreads=NCPU)
        assert_same_hits(got, want, f"pool frame {step}")
        b = T.ctx.beam_builds
        assert_same_hits(T.trace_frame(pos, yaw, pitch, W, H), want, f"pool frame {step}, again")      # (trace_frame syncs first: an empty delta)
        assert T.ctx.beam_builds == b and b == step + 1
        more = np.array([(int(x), int(y), int(z), 3, 0) for x, y, z in rs.randint(20, 100, (200, 3))], np.int32)
        A.apply(more); T.apply(more)
        T.sync()
    T.ctx.close()
