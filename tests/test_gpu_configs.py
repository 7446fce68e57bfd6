"""BASELINE.json's configurations at their STATED sizes against the oracle (round-1 review: rows N2 / N3), plus the
round-2 entry points: per-launch work counters (two caller streams), device-side deltas at any word offset, the
library's multi-GPU gather.  All comparisons bitwise (voxel, face, t), through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import assert_same_hits

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSES = {"A": ((1.5, 1.5, 1.5), 0.0, 0.0), "B": ((1.5, 1.5, 1.5), 0.7, -0.6), "C": ((1.1, 1.1, 1.4), 0.785, -0.3)}


def builtin_table():
    from test_oracle import _builtin_table
    return _builtin_table()


def cpu_reference(oc, log2cap, depth, nodes8, root):
    """(kind, trace(o, d)) -- the reference's own sse_trace where oracle/_ref carries the instantiation, else the port."""
    ncpu = max(1, min(64, os.cpu_count() or 1))
    if oc.have_ref() and (log2cap, depth) in oc.REF_CONFIGS:
        R = oc.RefTree(log2cap, depth)
        R.import_compact(nodes8, root)
        return "reference", lambda o, d: R.trace(o, d, nthreads=ncpu)
    tab = builtin_table()
    return "port", lambda o, d: oc.trace_rays(nodes8, root, depth, o, d, rcp_tab=tab, nthreads=ncpu)


@pytest.fixture(scope="module")
def depth12(ort):
    T = ort.HOctree(24, 12, node_capacity=1 << 21)
    ort.harness.build_terrain(T)
    T.sync()
    nodes8, root, _ = T.flatten()
    return T, nodes8, root


def test_config3_full_size_all_outputs_vs_reference(ort, oc, depth12):
    """Config 3 as BASELINE.json states it: depth-12 terrain, 2^24 = 16.7 M incoherent random-direction rays, every
    output of every ray against the reference's CPU trace -- through the default (persistent lane-refill) kernel and the
    one-thread-per-ray kernel."""
    T, nodes8, root = depth12
    n = 1 << 24
    o, d = ort.harness.random_rays(n)
    kind, trace = cpu_reference(oc, 24, 12, nodes8, root)
    want = trace(o, d)
    assert 0.05 < (want[0] != 0).mean() < 0.6
    ctx = T.ctx
    for rays_variant in (2, 1):
        ctx.set_option("rays_variant", rays_variant)
        got = ctx.trace_rays(o, d)
        assert_same_hits(got, want, f"config 3, 2^24 rays, rays_variant {rays_variant} vs {kind}")
    ctx.set_option("rays_variant", 2)


def test_config2_all_three_poses_4k_sampled_rows_vs_reference(ort, oc, depth12):
    """The bench's own frames (depth 12, 3840x2160, poses A, B, C): 108 rows of each against the reference."""
    T, nodes8, root = depth12
    kind, trace = cpu_reference(oc, 24, 12, nodes8, root)
    W, H = 3840, 2160
    for p, (pos, yaw, pitch) in POSES.items():
        v, f, t = T.trace_frame(pos, yaw, pitch, W, H)
        rot, fov = oc.camera_coeffs(yaw, pitch)
        rows = np.arange(3, H, 20)
        d = np.concatenate([oc.gen_rays(rot, fov, W, H, int(r), int(r) + 1) for r in rows])
        sel = (rows[:, None] * W + np.arange(W)[None, :]).ravel()
        assert_same_hits((v[sel], f[sel], t[sel]), trace(np.array(pos, np.float32), d), f"4K pose {p} vs {kind}")


@pytest.mark.parametrize("depth,log2cap", [(13, 26), (14, 27)])
def test_depth13_and_14_sampled_rows_vs_reference(ort, oc, depth, log2cap):
    """Config 5's DAG sizes: depth 13 (5 M nodes) and depth 14 (17.7 M nodes, 540 MiB, 4x the L2; reference counts
    saturate at 2^32 - 1, ids need 25 bits), 7680x4320 frames, sampled rows of poses A, B, C against the reference's own
    sse_trace (oracle/_ref instantiates h_octree<25,13> / <25,14> for exactly this; the compact array is imported)."""
    T = ort.HOctree(log2cap, depth, node_capacity=(6 << 20) if depth == 13 else (19 << 20))
    ort.harness.build_terrain(T)
    n_up, full = T.sync()
    nodes8, root, _ = T.flatten()
    assert full and n_up == nodes8.shape[0] and (4_000_000 if depth == 13 else 15_000_000) < n_up < (1 << 25)
    if depth == 14:
        assert int(T.refcounts().max()) == 0xFFFFFFFF, "depth 14 is expected to saturate reference counts"
    # the host table answers at() like the heightmap it was built from
    h = ort.harness.heightmap(depth)
    rs = np.random.RandomState(depth)
    for x, y in rs.randint(0, 1 << depth, (64, 2)):
        z = int(h[y, x])
        assert T.at(int(x), int(y), z) != 0 and T.at(int(x), int(y), z + 1) == 0
    kind, trace = cpu_reference(oc, 25, depth, nodes8, root)
    W, H = 7680, 4320
    for p, (pos, yaw, pitch) in POSES.items():
        rot, fov = oc.camera_coeffs(yaw, pitch)
        rows = np.arange(11, H, 173)
        got = [T.ctx.trace_frame(np.array(pos, np.float32), rot, fov, W, H, y0=int(r), rows=1) for r in rows]
        v, f, t = (np.concatenate([g[k] for g in got]) for k in range(3))
        d = np.concatenate([oc.gen_rays(rot, fov, W, H, int(r), int(r) + 1) for r in rows])
        assert_same_hits((v, f, t), trace(np.array(pos, np.float32), d), f"depth {depth} 8K pose {p} vs {kind}")
        assert (v != 0).sum() > 1000
    # a strip launch (cyclic 8-row tiles, rank 3 of 8) of the whole frame equals the same rows traced one by one
    pos, yaw, pitch = POSES["C"]
    rot, fov = oc.camera_coeffs(yaw, pitch)
    rows8 = ort.lib().ort_mg_strip_rows(3, 8, H, 8)
    sv, sf, st = T.ctx.trace_frame(np.array(pos, np.float32), rot, fov, W, H, y0=3 * 8, rows=rows8, tile_rows=8, tile_step=8)
    for k in (0, 17, rows8 // 8 - 1):
        y = (3 + 8 * k) * 8 + 5
        one = T.ctx.trace_frame(np.array(pos, np.float32), rot, fov, W, H, y0=y, rows=1)
        r = k * 8 + 5
        assert_same_hits((sv[r * W:(r + 1) * W], sf[r * W:(r + 1) * W], st[r * W:(r + 1) * W]), one, f"depth {depth} strip row {y}")


def test_persistent_launches_on_two_streams_do_not_share_a_counter(ort, golden):
    """ort_set_stream + ort_trace_rays_async on two caller streams at once (the default explicit-ray kernel is the
    persistent one): every launch draws from its own work counter, so both results equal the golden vectors."""
    import torch
    g = golden("d8_tunnels")
    ctx = ort.TraceContext(8)
    ctx.upload_full(g["nodes8"], int(g["root"]))
    o = torch.from_numpy(np.tile(g["rand_o"], (24, 1))).cuda()
    d = torch.from_numpy(np.tile(g["rand_d"], (24, 1))).cuda()
    m = o.shape[0]
    want = tuple(np.tile(g[f"rand_{k}"], 24) for k in ("vox", "face", "t"))
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [(torch.zeros(m, dtype=torch.int32, device="cuda"), torch.zeros(m, dtype=torch.uint8, device="cuda"), torch.zeros(m, dtype=torch.float32, device="cuda")) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(6):
        for k in range(2):
            ctx.set_stream(s[k])
            ctx.trace_rays_async(o, 3, d, m, *outs[k])
    ctx.set_stream(None)
    torch.cuda.synchronize()
    for k in range(2):
        got = (outs[k][0].cpu().numpy().view(np.uint32), outs[k][1].cpu().numpy(), outs[k][2].cpu().numpy())
        assert_same_hits(got, want, f"stream {k}")
    # frames through the persistent kernel (variant 2) on two streams as well
    ctx.set_option("variant", 2)
    W, H = int(g["W"]), int(g["H"])
    fo = [(torch.zeros(W * H, dtype=torch.int32, device="cuda"), torch.zeros(W * H, dtype=torch.uint8, device="cuda"), torch.zeros(W * H, dtype=torch.float32, device="cuda")) for _ in range(2)]
    for rep in range(4):
        for k, p in enumerate("BC"):
            ctx.set_stream(s[k])
            ctx.trace_frame_async(g[f"pose{p}_pos"], g[f"pose{p}_rot"], float(g[f"pose{p}_fov"]), W, H, 0, H, 1, 1, *fo[k])
    ctx.set_stream(None)
    torch.cuda.synchronize()
    for k, p in enumerate("BC"):
        got = (fo[k][0].cpu().numpy().view(np.uint32), fo[k][1].cpu().numpy(), fo[k][2].cpu().numpy())
        assert_same_hits(got, (g[f"pose{p}_vox"], g[f"pose{p}_face"], g[f"pose{p}_t"]), f"persistent frame {p} on stream {k}")


def test_device_side_delta_at_any_word_offset(ort, oc):
    """ort_upload_delta with DEVICE pointers whose rows are only 4-byte aligned (a broadcast payload [ids | rows] with
    n % 4 != 0): the scatter must not assume 16-byte alignment (round-1 advisor finding).  A numpy mirror of the compact
    array takes the same deltas and feeds the oracle."""
    import torch
    from octree_ray_tracing_b200._lib import check
    depth = 6
    T = ort.HOctree(16, depth, device=None)
    ort.harness.build_terrain(T, tunnels=True)
    ctx_host, ctx_dev = ort.TraceContext(depth), ort.TraceContext(depth)
    ids, rows, root, is_full = T.take_delta()
    assert is_full
    mirror = np.zeros((max(4096, 2 * rows.shape[0]), 8), np.uint32)
    mirror[: rows.shape[0]] = rows
    ctx_host.upload_full(mirror, root)                         # (spare rows: deltas may hand out new ids)
    ctx_dev.upload_full(mirror, root)
    pos, yaw, pitch = POSES["C"]
    rot, fov = oc.camera_coeffs(yaw, pitch)
    d = oc.gen_rays(rot, fov, 320, 200)
    seen_offsets = set()
    for k, ext in enumerate((3, 5, 2, 7, 1, 4)):
        T.set_box(20 + 3 * k, 30, 24, ext, 2 + k % 3)
        ids, rows, root, is_full = T.take_delta()
        assert not is_full and rows.shape[0] > 0
        n = rows.shape[0]
        mirror[ids - 1] = rows
        off = 1 + k % 3                                           # ids at word `off`, rows at word off + n: any alignment
        seen_offsets.add((off + n) % 4)
        payload = torch.from_numpy(np.concatenate([np.zeros(off, np.uint32), ids.astype(np.uint32), rows.reshape(-1)]).view(np.int32)).cuda()
        base = payload.data_ptr()
        check(ort.lib().ort_upload_delta(ctx_dev.h, base + 4 * off, base + 4 * (off + n), n, int(root)), ctx_dev.h)
        del payload                                               # the call returns after the scatter has read its source
        ctx_host.upload_delta(ids, rows, int(root))               # host-pointer path
        w = oc.trace_rays(mirror, int(root), depth, np.array(pos, np.float32), d, rcp_tab=builtin_table(), nthreads=4)
        a = ctx_host.trace_frame(np.array(pos, np.float32), rot, fov, 320, 200)
        b = ctx_dev.trace_frame(np.array(pos, np.float32), rot, fov, 320, 200)
        assert_same_hits(a, w, f"edit {k}: host-pointer delta")
        assert_same_hits(b, w, f"edit {k}: device-pointer delta, rows at word offset {off + n}")
    assert len(seen_offsets) > 1, "the test must exercise several row alignments"


def test_frame_argument_checks_come_before_any_work(ort):
    """ort_trace_frame with host outputs used to divide by tile_rows before validating it (round-1 advisor finding)."""
    ctx = ort.TraceContext(4)
    nodes8 = np.zeros((1, 8), np.uint32)
    ctx.upload_full(nodes8, 1)
    rot, fov = ort.camera_coeffs(0.1, 0.1)
    pos = np.array([1.5, 1.5, 1.5], np.float32)
    for kw in (dict(tile_rows=0), dict(tile_rows=-8), dict(tile_step=0), dict(y0=-1)):
        with pytest.raises(ort.OrtError):
            ctx.trace_frame(pos, rot, fov, 64, 32, **kw)
    v, f, t = ctx.trace_frame(pos, rot, fov, 64, 32)
    assert (f == 6).all()


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_multi_gpu_gather_in_the_library_vs_reference():
    """ort_mg_* on real devices, world size = all visible GPUs (max 4): DAG broadcast from rank 0 (full, then two edit
    deltas), strips traced by the CUDA kernels, gathered over NCCL inside the library, assembled frames compared on
    rank 0 with the ORACLE on every row -- and with the single-GPU trace of the same frame."""
    world = min(_n_gpus(), 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(ROOT, "tests", "mg_worker.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MG-OK" in out.stdout, out.stdout[-2000:]
