"""Parity of the CUDA path (through the C ABI) against the reference's golden outputs and the CPU
oracle.  Bar: voxel ids, faces AND hit times bit-exact (north_star allows 1e-5 relative on t; the
kernel reproduces the reference's arithmetic exactly, so the tests hold it to 0 ulp)."""
import numpy as np
import pytest

from conftest import assert_same_hits

pytestmark = pytest.mark.gpu

POSES = {"A": ((1.5, 1.5, 1.5), 0.0, 0.0), "B": ((1.5, 1.5, 1.5), 0.7, -0.6), "C": ((1.1, 1.1, 1.4), 0.785, -0.3)}


def builtin_table():
    from test_oracle import _builtin_table
    return _builtin_table()


@pytest.fixture(scope="module")
def ncpu():
    import os
    return max(1, min(32, os.cpu_count() or 1))


@pytest.mark.parametrize("name", ["d6_tunnels", "d8_tunnels"])
def test_golden_from_the_real_reference(ort, golden, name):
    """Inputs and expected outputs were produced by the unmodified reference (tests/golden/make_golden.py)."""
    g = golden(name)
    depth = int(g["depth"])
    ctx = ort.TraceContext(depth)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    W, H = int(g["W"]), int(g["H"])
    for p in "ABC":
        got = ctx.trace_frame(g[f"pose{p}_pos"], g[f"pose{p}_rot"], float(g[f"pose{p}_fov"]), W, H)
        assert_same_hits(got, (g[f"pose{p}_vox"], g[f"pose{p}_face"], g[f"pose{p}_t"]), f"{name} frame {p}")
    for k in ("rand", "edge"):
        got = ctx.trace_rays(g[f"{k}_o"], g[f"{k}_d"])
        assert_same_hits(got, (g[f"{k}_vox"], g[f"{k}_face"], g[f"{k}_t"]), f"{name} {k} rays")
    ctx.close()


def test_host_rcp_table_override_matches_oracle_hw_mode(ort, oc, golden):
    """With the table derived from THIS box's RCPSS the GPU equals the oracle running the real
    instruction (i.e. the reference as it would run on this host)."""
    tab, bad = oc.rcp_table_from_hw(11)
    if bad:
        pytest.skip(f"host RCPSS is not an 11-bit table function ({bad} mismatches); built-in table covered elsewhere")
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.set_rcp_table(tab)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    for k in ("rand", "edge"):
        want = oc.trace_rays(g["nodes8"], int(g["root"]), 8, g[f"{k}_o"], g[f"{k}_d"])     # hardware RCPSS
        assert_same_hits(ctx.trace_rays(g[f"{k}_o"], g[f"{k}_d"]), want, k)


def test_in_kernel_rays_equal_explicit_rays_and_oracle_rays(ort, oc, golden):
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    for W, H in ((317, 203), (1280, 720)):
        rot, fov = ort.camera_coeffs(0.7, -0.6)
        d = oc.gen_rays(rot, fov, W, H)
        pos = np.array([1.5, 1.5, 1.5], np.float32)
        a = ctx.trace_frame(pos, rot, fov, W, H)
        b = ctx.trace_rays(pos, d)
        assert_same_hits(a, b, f"{W}x{H} generated vs explicit")


def test_strips_and_cyclic_tiles_reassemble_the_frame(ort, golden):
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    W, H = 640, 360
    rot, fov = ort.camera_coeffs(0.785, -0.3)
    pos = np.array([1.1, 1.1, 1.4], np.float32)
    full = [x.reshape(H, W) for x in ctx.trace_frame(pos, rot, fov, W, H)]
    # contiguous strips of odd heights
    y = 0
    for rows in (1, 37, 100, 222):
        part = ctx.trace_frame(pos, rot, fov, W, H, y0=y, rows=rows)
        assert_same_hits(part, [f[y:y + rows].ravel() for f in full], f"strip {y}+{rows}")
        y += rows
    assert y == H
    # cyclic 8-row tiles over 4 "ranks" (H = 360 = 45 tiles: ranks get 12,11,11,11 tiles); tile heights that are
    # powers of two take the shift/mask row mapping, the others the division (ort::frame_row)
    ctx.set_palette((np.arange(6 * 8, dtype=np.uint32).reshape(8, 6) * 0x010305) | 0xFF000000)
    rgba_full = ctx.trace_frame_rgba(pos, rot, fov, W, H).reshape(H, W)
    assert len(np.unique(rgba_full)) > 6
    for tr, N in ((8, 4), (16, 3), (1, 5), (5, 3), (12, 2)):
        for rank in range(N):
            tiles = list(range(rank, H // tr, N))
            rows = len(tiles) * tr
            want_rows = np.concatenate([np.arange(t * tr, (t + 1) * tr) for t in tiles])
            part = ctx.trace_frame(pos, rot, fov, W, H, y0=rank * tr, rows=rows, tile_rows=tr, tile_step=N)
            assert_same_hits(part, [f[want_rows].ravel() for f in full], f"cyclic {tr}-row tiles, rank {rank} of {N}")
            rgba = ctx.trace_frame_rgba(pos, rot, fov, W, H, y0=rank * tr, rows=rows, tile_rows=tr, tile_step=N)
            assert np.array_equal(rgba.ravel(), rgba_full[want_rows].ravel()), f"rgba, cyclic {tr}-row tiles, rank {rank} of {N}"


def test_empty_tree_and_tiny_inputs(ort):
    t = ort.HOctree(12, 4)
    v, f, tt = t.trace_rays(np.array([1.5, 1.5, 1.5], np.float32), np.array([[0, 0, -1], [1, 0, 0]], np.float32))
    assert list(v) == [0, 0] and list(f) == [6, 6] and np.isinf(tt).all()
    d, vox, tm = t.sse_trace((1.5, 1.5, 1.5), (0.0, 0.0, -1.0))
    assert d == ort.Direction.exit and vox == 0 and tm == float("inf")
    t.set(8, 8, 3, 5)
    d, vox, tm = t.sse_trace((1.53, 1.53, 1.9), (0.0, 0.0, -1.0))     # axis-parallel: the reference's NaN path
    ctx = t.ctx
    assert ctx.trace_rays(np.zeros(3, np.float32), np.zeros((0, 3), np.float32))[0].size == 0
    v, f, tt = t.trace_frame((1.5, 1.5, 1.9), 0.0, -1.5, 1, 1)
    assert v.size == 1


def test_depth10_terrain_frames_vs_oracle(ort, oc, ncpu):
    """BASELINE config 1 shape: depth-10 (1024^3) terrain, 1280x720 primary rays, poses A/B/C."""
    depth = 10
    T = ort.HOctree(22, depth)
    ort.harness.build_terrain(T)
    nodes8, root, _ = T.flatten()
    tab = builtin_table()
    for p, (pos, yaw, pitch) in POSES.items():
        got = T.trace_frame(pos, yaw, pitch, 1280, 720, want_npush=True)
        rot, fov = oc.camera_coeffs(yaw, pitch)
        d = oc.gen_rays(rot, fov, 1280, 720)
        wv, wf, wt, wn, tot = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=ncpu, want_counts=True)
        assert_same_hits(got, (wv, wf, wt), f"depth 10 pose {p}")
        assert np.array_equal(got[3], wn), "per-ray PUSH counts differ from the oracle's"
        assert (wv != 0).sum() > 1000


def test_incoherent_rays_vs_oracle(ort, oc, ncpu):
    """BASELINE config 3 shape (scaled to what the oracle does in seconds): random-direction rays."""
    depth = 10
    T = ort.HOctree(22, depth)
    ort.harness.build_terrain(T)
    nodes8, root, _ = T.flatten()
    o, d = ort.harness.random_rays(1 << 20)
    got = T.trace_rays(o, d)
    want = oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=builtin_table(), nthreads=ncpu)
    assert_same_hits(got, want, "incoherent")
    assert 0.05 < (got[0] != 0).mean() < 0.6


def test_edit_loop_with_delta_uploads(ort, oc, ncpu):
    """BASELINE config 4 shape: frames interleaved with 40^3 place/remove edits at the crosshair hit
    (test_och_h_octree.cpp:366-433), device mirror updated by deltas only; every frame is compared
    with the oracle tracing a plain oracle table that received the same set() calls."""
    depth, log2cap = 8, 19
    h, g = oc.heightmap(depth), oc.grass_bits(depth)
    A = oc.OracleTree(log2cap, depth)
    A.initialize_terrain(h, g, False)
    T = ort.HOctree(log2cap, depth)
    ort.harness.build_terrain(T, h, g)
    dim = 1 << depth
    pos = np.array([1.5, 1.5, 1.0 + (h[dim // 2, dim // 2] + 0.2 * dim * 0.25) / dim], np.float32)
    yaw, pitch = 0.4, -1.2
    W, H = 320, 180
    rot, fov = oc.camera_coeffs(yaw, pitch)
    d = oc.gen_rays(rot, fov, W, H)
    tab = builtin_table()
    n_full = n_delta = 0
    for frame in range(12):
        # pick ray = camera forward (test_och_h_octree.cpp:527, :535-536)
        dir3 = np.array([np.cos(np.float32(yaw)) * np.cos(np.float32(pitch)), np.sin(np.float32(yaw)) * np.cos(np.float32(pitch)), np.sin(np.float32(pitch))], np.float32)
        face, vox, t = T.sse_trace(pos, dir3)
        ov, of, ot = A.trace(pos, dir3.reshape(1, 3), rcp_tab=tab)
        assert (int(face), vox) == (int(of[0]), int(ov[0])) and np.float32(t).view(np.uint32) == ot.view(np.uint32)[0]
        if vox and t < 0.5:
            off = np.zeros(3, np.float32)
            if int(face) < 6:
                off[int(face) % 3] = (T.voxel_dim / 2) * (1 if int(face) < 3 else -1)      # :487-502
            place = frame % 2 == 0
            cp = pos + dir3 * np.float32(t) + (off if place else -off) - np.float32(1.0)   # :399-402 (T) / :418-421 (Z)
            c = (cp * np.float32(dim)).astype(np.uint16)
            ext, v = 40, (1 if place else 0)
            T.set_box(int(c[0]), int(c[1]), int(c[2]), ext, v)
            ops = np.array([((int(c[0]) + x) & 0xFFFF, (int(c[1]) + y) & 0xFFFF, (int(c[2]) + z) & 0xFFFF, v)
                            for z in range(-20, 20) for y in range(-20, 20) for x in range(-20, 20)], np.uint32)
            A.set_many(ops)
        n, full = T.sync()
        n_full += full
        n_delta += (not full) and n > 0
        got = T.trace_frame(pos, yaw, pitch, W, H)
        assert_same_hits(got, A.trace(pos, d, rcp_tab=tab, nthreads=ncpu), f"frame {frame}")
    assert n_full <= 1 and n_delta >= 3, (n_full, n_delta)
    assert (T.get_fillcnt(), T.get_nodecnt()) == (A.fillcnt, A.nodecnt)


def test_device_pointer_entry_points(ort, oc, golden):
    """ort_trace_*_async with device buffers (torch is only the allocator here)."""
    import torch
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    W, H = int(g["W"]), int(g["H"])
    n = W * H
    dv = torch.empty(n, dtype=torch.int32, device="cuda")
    df = torch.empty(n, dtype=torch.uint8, device="cuda")
    dt = torch.empty(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.trace_frame_async(g["poseB_pos"], g["poseB_rot"], float(g["poseB_fov"]), W, H, 0, H, 1, 1, dv, df, dt)
    ctx.sync()
    got = (dv.cpu().numpy().view(np.uint32), df.cpu().numpy(), dt.cpu().numpy())
    assert_same_hits(got, (g["poseB_vox"], g["poseB_face"], g["poseB_t"]), "async frame")
    o = torch.from_numpy(g["rand_o"]).cuda()
    d = torch.from_numpy(g["rand_d"]).cuda()
    m = o.shape[0]
    torch.cuda.synchronize()
    ctx.trace_rays_async(o, 3, d, m, dv, df, dt)
    ctx.sync()
    got = (dv[:m].cpu().numpy().view(np.uint32), df[:m].cpu().numpy(), dt[:m].cpu().numpy())
    assert_same_hits(got, (g["rand_vox"], g["rand_face"], g["rand_t"]), "async rays")
    # one frame kernel + one rays kernel; with the beam start the frame also has its march and, once per DAG version, the
    # grid build (occupancy, dilation, k - 1 pyramid levels, skip levels)
    # (default policy: the first two frame calls of a DAG version run without a grid, see ort_set_option "beam_after")
    assert ctx.launch_count == 2
    ctx.set_option("beam_after", 0)
    k = ctx.beam_level(g["poseB_pos"], g["poseB_rot"], float(g["poseB_fov"]), W, H)
    ctx.trace_frame_async(g["poseB_pos"], g["poseB_rot"], float(g["poseB_fov"]), W, H, 0, H, 1, 1, dv, df, dt)
    ctx.sync()
    assert ctx.launch_count == 3 + ((1 + 3 + (k - 1)) if k else 0)
    got = (dv.cpu().numpy().view(np.uint32), df.cpu().numpy(), dt.cpu().numpy())
    assert_same_hits(got, (g["poseB_vox"], g["poseB_face"], g["poseB_t"]), "async frame with the beam start")


def test_depth12_4k_properties(ort, oc, ncpu):
    """BASELINE config 2 at full size: depth-12 (4096^3) terrain DAG, 3840x2160.  The oracle is run
    on a sample of rows; the whole frame is checked through size-independent properties."""
    depth = 12
    T = ort.HOctree(24, depth)
    ort.harness.build_terrain(T)
    nodes8, root, lo = T.flatten()
    assert 1_000_000 < nodes8.shape[0] < 2_000_000
    tab = builtin_table()
    W, H = 3840, 2160
    pos, yaw, pitch = POSES["B"]
    v, f, t, npush = T.trace_frame(pos, yaw, pitch, W, H, want_npush=True)
    # (1) oracle on every 27th row
    rot, fov = oc.camera_coeffs(yaw, pitch)
    rows = np.arange(5, H, 27)
    d = np.concatenate([oc.gen_rays(rot, fov, W, H, int(r), int(r) + 1) for r in rows])
    wv, wf, wt, wn, tot = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=ncpu, want_counts=True)
    sel = (rows[:, None] * W + np.arange(W)[None, :]).ravel()
    assert_same_hits((v[sel], f[sel], t[sel]), (wv, wf, wt), "4K sample rows")
    assert np.array_equal(npush[sel], wn)
    # (2) idempotence + strip independence: bottom half alone equals the bottom half of the frame
    v2, f2, t2 = T.trace_frame(pos, yaw, pitch, W, H, y0=H // 2, rows=H // 2)
    assert_same_hits((v2, f2, t2), (v[W * (H // 2):], f[W * (H // 2):], t[W * (H // 2):]), "half frame")
    # (3) physical sanity of every hit: the hit point lies on the entry face of a voxel cell (within float error)
    hit = np.flatnonzero((v != 0) & (f < 6))
    samp = hit[:: max(1, hit.size // 200000)]
    ys, xs = samp // W, samp % W
    dd = np.concatenate([oc.gen_rays(rot, fov, W, H, int(y), int(y) + 1)[x][None] for y, x in zip(ys[:2000], xs[:2000])])
    p = np.array(pos, np.float64)[None] + dd.astype(np.float64) * t[samp[:2000]].astype(np.float64)[:, None]
    ax = f[samp[:2000]] % 3
    coord = (p[np.arange(len(ax)), ax] - 1.0) * (1 << depth)
    assert np.abs(coord - np.round(coord)).max() < 0.6       # RCPPS-grade t: well within one voxel of a cell plane
    assert set(np.unique(v)) <= {0, 1, 2, 3, 4}


def test_kernel_variants_agree(ort, golden):
    """Baseline walk (variant 0), round 1's fast walk (1), persistent lane refill (2), the round-2 tiers (13) -- and, in
    the measurement build, every experiment kernel -- are the same function."""
    from conftest import frame_variants
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    W, H = 333, 217                                    # not multiples of the 8x4 tile
    pos, rot, fov = g["poseC_pos"], g["poseC_rot"], float(g["poseC_fov"])
    ref = None
    ctx.set_option("smem_levels", 215)                  # variant 3 stages the first 215 nodes (levels 1-4 of this DAG)
    for variant in frame_variants(ort):
        ctx.set_option("variant", variant)
        got = ctx.trace_frame(pos, rot, fov, W, H, want_npush=True)
        part = ctx.trace_frame(pos, rot, fov, W, H, y0=8, rows=64, tile_rows=8, tile_step=3)
        rows = np.concatenate([np.arange(8 + 24 * k, 16 + 24 * k) for k in range(8)])
        if ref is None:
            ref = got
        assert_same_hits(got, ref, f"frame variant {variant}")
        assert np.array_equal(got[3], ref[3]), f"npush differs in variant {variant}"
        assert_same_hits(part, [x.reshape(H, W)[rows].ravel() for x in ref[:3]], f"tiles variant {variant}")
    for k in ("rand", "edge"):
        outs = []
        for variant, rv in ((0, 1), (1, 1), (13, 1), (13, 2)):
            ctx.set_option("variant", variant)
            ctx.set_option("rays_variant", rv)
            for lw in ((20,) if rv == 1 else (0, 12, 31)):
                ctx.set_option("low_water", lw)
                outs.append(ctx.trace_rays(g[f"{k}_o"], g[f"{k}_d"], want_npush=True))
        for o in outs:
            assert_same_hits(o, (g[f"{k}_vox"], g[f"{k}_face"], g[f"{k}_t"]), k)
            assert np.array_equal(o[3], outs[0][3])


@pytest.mark.gpu
@pytest.mark.parametrize("depth,log2cap", [(6, 16), (8, 19), (10, 22)])
def test_gpu_fixture_kernels_match_host_and_oracle(ort, oc, depth, log2cap):
    """the fixture's noise on the GPU -- heightmap and tunnel bitmap -- is bit-identical to the host
    builder's (and the heightmap to the oracle's), so both routes intern the same nodes in the same order."""
    ctx_tree = ort.HOctree(log2cap, depth, device=0)
    h_gpu = ort.harness.heightmap_gpu(ctx_tree.ctx, depth)
    h_host = ort.harness.heightmap(depth, 4)
    assert np.array_equal(h_gpu, h_host)
    if depth <= 8:
        assert np.array_equal(h_gpu, oc.heightmap(depth))
    g = ort.harness.grass_bits(depth)
    ort.harness.build_terrain(ctx_tree, h_gpu, g, tunnels=True, gpu=True)
    host_tree = ort.HOctree(log2cap, depth, device=None)
    ort.harness.build_terrain(host_tree, h_host, g, tunnels=True, gpu=False)
    assert (ctx_tree.get_fillcnt(), ctx_tree.get_nodecnt(), ctx_tree.get_root()) == (host_tree.get_fillcnt(), host_tree.get_nodecnt(), host_tree.get_root())
    assert np.array_equal(ctx_tree.cashes(), host_tree.cashes())
    assert np.array_equal(ctx_tree.nodes(), host_tree.nodes())
    assert np.array_equal(ctx_tree.refcounts(), host_tree.refcounts())


@pytest.mark.gpu
def test_host_buffer_pipelines_keep_results(ort, golden):
    """The host-buffer paths are pipelines (chunk kernels on several streams, staged D2H, optional deferred completion,
    chunked H2D/trace/D2H for explicit rays): every chunking / completion mode returns the same bits."""
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    W, H = 419, 301
    poses = [(g["poseC_pos"], g["poseC_rot"], float(g["poseC_fov"]))]
    rot2, fov2 = ort.camera_coeffs(0.3, -0.9)
    poses.append((np.array([1.3, 1.6, 1.7], np.float32), rot2, fov2))
    cols, _ = ort.harness.parse_voxels(ort.harness.DEMO_VOXELS)
    ctx.set_palette(cols)
    ref = [ctx.trace_frame(p, r, f, W, H, want_npush=True) for p, r, f in poses]
    ref_rgba = [ctx.trace_frame_rgba(p, r, f, W, H).copy() for p, r, f in poses]
    for chunks in (1, 2, 5, 8):
        ctx.set_option("frame_chunks", chunks)
        for (p, r, f), want, want_rgba in zip(poses, ref, ref_rgba):
            got = ctx.trace_frame(p, r, f, W, H, want_npush=True)
            assert_same_hits(got, want, f"{chunks} chunks")
            assert np.array_equal(got[3], want[3])
            assert np.array_equal(ctx.trace_frame_rgba(p, r, f, W, H), want_rgba)
    # deferred completion: queue everything, collect once; cyclic tile strips as well
    ctx.set_option("frame_chunks", 0)
    ctx.set_option("defer_sync", 1)
    outs = [(np.zeros(W * H, np.uint32), np.zeros(W * H, np.uint8), np.zeros(W * H, np.float32), None) for _ in range(6)]
    rgbas = [np.zeros(W * H, np.uint32) for _ in range(6)]
    for k in range(6):
        p, r, f = poses[k % 2]
        ctx.trace_frame(p, r, f, W, H, out=outs[k])
        ctx.trace_frame_rgba(p, r, f, W, H, out=rgbas[k])
    ctx.sync()
    ctx.set_option("defer_sync", 0)
    for k in range(6):
        assert_same_hits(outs[k], ref[k % 2], f"deferred frame {k}")
        assert np.array_equal(rgbas[k], ref_rgba[k % 2])
    # explicit rays through many small pipeline stages, mixed with a frame call in between
    o, d = g["rand_o"], g["rand_d"]
    want = (g["rand_vox"], g["rand_face"], g["rand_t"])
    for chunk in (4096, 5000, 1 << 20):
        ctx.set_option("rays_chunk", chunk)
        assert_same_hits(ctx.trace_rays(o, d), want, f"rays_chunk {chunk}")
        assert_same_hits(ctx.trace_frame(*poses[0], W, H), ref[0], "frame after rays")
    shared = np.array([1.5, 1.5, 1.9], np.float32)
    dd = np.ascontiguousarray(d[:20000])
    a = ctx.trace_rays(shared, dd)
    ctx.set_option("rays_chunk", 4096)
    b = ctx.trace_rays(shared, dd)
    assert_same_hits(a, b, "shared origin, chunked")


@pytest.mark.gpu
@pytest.mark.parametrize("depth,log2cap", [(1, 8), (2, 8), (16, 18)])
def test_extreme_depths_trace_vs_oracle(ort, oc, depth, log2cap):
    """Depth 1 (one node) and depth 16 (65536^3, a 16-entry parent stack, voxel-size 2^-16 steps): GPU trace of random,
    axis-parallel and corner-grazing rays equals the oracle on the same table."""
    rs = np.random.RandomState(100 + depth)
    dim = 1 << depth
    A, T = oc.OracleTree(log2cap, depth), ort.HOctree(log2cap, depth)
    n = 6 if depth <= 2 else 4000
    pts = rs.randint(0, dim, (n, 3))
    if depth == 16:                                           # a dense blob so that rays actually hit something
        pts[: n // 2] = 32768 + rs.randint(-40, 40, (n // 2, 3))
    pts[:2] = [[0, 0, 0], [dim - 1, dim - 1, dim - 1]]
    ops = np.concatenate([pts, rs.randint(1, 5, (n, 1))], 1).astype(np.uint32)
    A.set_many(ops)
    T.set_many(ops)
    m = 30000
    o = rs.uniform(1.001, 1.999, (m, 3)).astype(np.float32)
    target = (1.0 + (pts[rs.randint(0, n, m)] + rs.uniform(0, 1, (m, 3))) / dim).astype(np.float32)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    d[:300, 1:] = 0.0                                        # axis-parallel
    d[300:600, 2] = 0.0
    o[600:900] = np.float32(1.5)                             # origins on cell planes
    got = T.trace_rays(o, d, want_npush=True)
    want = A.trace(o, d)
    assert_same_hits(got, want, f"depth {depth}")
    assert (got[0] != 0).sum() > m // 10


@pytest.mark.gpu
def test_edits_between_deferred_frames(ort, oc, golden):
    """A frame loop with deferred completion and edits in between: frame k is still crossing PCIe when the delta of
    edit k+1 is staged and frame k+1 is queued.  Every frame must show exactly the tree as it was when it was queued."""
    depth, log2cap = 8, 19
    h, g = oc.heightmap(depth), oc.grass_bits(depth)
    A = oc.OracleTree(log2cap, depth)
    A.initialize_terrain(h, g, False)
    T = ort.HOctree(log2cap, depth)
    ort.harness.build_terrain(T, h, g, gpu=False)
    T.sync()
    tab = builtin_table()
    W, H = 640, 360
    pos = np.array([1.5, 1.5, 1.62], np.float32)
    rot, fov = ort.camera_coeffs(0.4, -1.0)
    d = oc.gen_rays(rot, fov, W, H)
    rs = np.random.RandomState(77)
    frames, wants = [], []
    T.ctx.set_option("defer_sync", 1)
    for k in range(6):
        c = rs.randint(60, 190, 3)
        c[2] = int(h[c[1], c[0]])
        v = k % 2
        T.fill_box(c - 12, c + 12, v)
        box = np.array([(x, y, z, v) for z in range(c[2] - 12, c[2] + 12) for y in range(c[1] - 12, c[1] + 12) for x in range(c[0] - 12, c[0] + 12)
                        if 0 <= z < 256], np.uint32)
        A.set_many(box)
        n_up, full = T.sync()                                        # delta upload while earlier frames drain
        assert n_up > 0 and not full
        out = (np.zeros(W * H, np.uint32), np.zeros(W * H, np.uint8), np.zeros(W * H, np.float32), None)
        T.ctx.trace_frame(pos, rot, fov, W, H, out=out)              # returns once queued
        frames.append(out)
        wants.append(A.trace(pos, d, rcp_tab=tab, nthreads=4))       # the straight table that saw the same edits
        assert (A.fillcnt, A.nodecnt) == (T.get_fillcnt(), T.get_nodecnt())
    T.ctx.sync()
    T.ctx.set_option("defer_sync", 0)
    for k in range(6):
        assert_same_hits(frames[k], wants[k], f"deferred frame {k}")
    assert any(not np.array_equal(frames[k][0], frames[k + 1][0]) for k in range(5))     # the edits are visible


@pytest.mark.gpu
def test_batched_frames_equal_separate_launches(ort, golden):
    """ort_trace_frames_async: jobs of different poses, sizes and strip layouts in one launch give what one
    ort_trace_frame per job gives (20 jobs: more than one parameter batch)."""
    import torch
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    rs = np.random.RandomState(4)
    jobs, wants, outs = [], [], []
    for k in range(20):
        W, H = int(rs.randint(40, 300)), int(rs.randint(30, 200))
        pos = rs.uniform(1.05, 1.95, 3).astype(np.float32)
        rot, fov = ort.camera_coeffs(float(rs.uniform(-3, 3)), float(rs.uniform(-1.4, 0.6)))
        if k % 3 == 0:
            y0, rows, tr, ts = 8 * (k % 2), None, 8, 2
            rows = len([y for y in range(H) if (y // 8) % 2 == (k % 2)])
        else:
            y0, rows, tr, ts = int(rs.randint(0, H // 2)), None, 1, 1
            rows = int(rs.randint(1, H - y0 + 1))
        want = ctx.trace_frame(pos, rot, fov, W, H, y0=y0, rows=rows, tile_rows=tr, tile_step=ts, want_npush=(k % 4 == 0))
        n = rows * W
        dv = torch.zeros(n, dtype=torch.int32, device="cuda")
        df = torch.zeros(n, dtype=torch.uint8, device="cuda")
        dt = torch.zeros(n, dtype=torch.float32, device="cuda")
        dn = torch.zeros(n, dtype=torch.int16, device="cuda") if k % 4 == 0 else None
        jobs.append((pos, rot, fov, W, H, y0, rows, tr, ts, dv, df, dt, dn))
        wants.append(want)
        outs.append((dv, df, dt, dn))
    ctx.trace_frames_async(jobs)
    ctx.sync()
    torch.cuda.synchronize()
    for k, (want, (dv, df, dt, dn)) in enumerate(zip(wants, outs)):
        got = (dv.cpu().numpy().view(np.uint32), df.cpu().numpy(), dt.cpu().numpy())
        assert_same_hits(got, want, f"job {k}")
        if dn is not None:
            assert np.array_equal(dn.cpu().numpy().view(np.uint16), want[3])


@pytest.mark.gpu
def test_degenerate_rays_and_corner_cameras_vs_oracle(ort, oc, ncpu):
    """Rays outside the fast walkers' preconditions must take the baseline walk in every kernel: zero / denormal
    direction components, origins outside [1,2)^3, on-plane origins and -- found by tests/test_host_emu.py -- a
    coordinate of exactly 1.0f travelled in the positive direction (mirrored to 2.0f, masked position bits 0).
    Explicit-ray kernels (one-shot and persistent) and frame kernels (camera on the cube's corner / faces)."""
    from conftest import degenerate_rays, frame_variants
    depth = 8
    T = ort.HOctree(19, depth)
    ort.harness.build_terrain(T, tunnels=True)
    nodes8, root, _ = T.flatten()
    tab = builtin_table()
    O, D = degenerate_rays(ort, depth)
    assert (O == 1.0).any(axis=1).sum() > 100
    want = oc.trace_rays(nodes8, root, depth, O, D, rcp_tab=tab, nthreads=ncpu, want_counts=True)
    T.sync()
    ctx = T.ctx
    for variant, rays_variant in ((0, 1), (1, 1), (13, 1), (13, 2)):
        ctx.set_option("variant", variant)
        ctx.set_option("rays_variant", rays_variant)
        got = ctx.trace_rays(O, D, want_npush=True)
        assert_same_hits(got, want[:3], f"variant {variant}, rays_variant {rays_variant}")
        assert np.array_equal(got[3], want[3]), f"variant {variant}, rays_variant {rays_variant}: PUSH counts"
    ctx.set_option("rays_variant", 1)
    W, H = 320, 200
    for pos, yaw, pitch in [((1.0, 1.0, 1.0), 0.785, 0.6), ((1.0, 1.5, 1.75), 0.0, 0.0), ((1.5, 1.0, 1.5), 1.5708, -0.2), ((1.25, 1.5, 1.0), 0.3, 1.2)]:
        rot, fov = oc.camera_coeffs(yaw, pitch)
        d = oc.gen_rays(rot, fov, W, H)
        wv, wf, wt, wn, _ = oc.trace_rays(nodes8, root, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=ncpu, want_counts=True)
        for variant in frame_variants(ort):
            ctx.set_option("variant", variant)
            got = ctx.trace_frame(np.array(pos, np.float32), rot, fov, W, H, want_npush=True)
            assert_same_hits(got, (wv, wf, wt), f"camera at {pos}, variant {variant}")
            assert np.array_equal(got[3], wn), f"camera at {pos}, variant {variant}: PUSH counts"
    ctx.set_option("variant", 13)
