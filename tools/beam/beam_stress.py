"""Offline stress of the beam start on the host emulation (CPU only): 160 random cameras -- in the air, just above the ground,
inside the ground band / tunnels, on grids of random levels -- at four frame sizes on depth-10 (tunnels) and depth-11 terrain;
every frame must equal the oracle, no tile start may exceed a hit time, the guard must stay silent.  python tools/beam/beam_stress.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'host_emu'))
import numpy as np, emu
import octree_ray_tracing_b200 as ort
from octree_ray_tracing_b200 import harness
from oracle import oracle as oc
from conftest import assert_same_hits
tab = emu.default_rcp_table()
bad = 0
for depth, tunnels in [(10, True), (11, False)]:
    T = ort.HOctree(14 + depth, depth, device=None)
    heights, _ = harness.build_terrain(T, tunnels=tunnels)
    nodes8, root, _ = T.flatten()
    dim = 1 << depth
    rs = np.random.RandomState(depth)
    grids = {}
    n_frames = 0
    for i in range(80):
        # cameras: random in the air, just above the ground, inside the ground band (tunnels), on grids of random levels
        x, y = rs.randint(8, dim - 8, 2)
        h = int(heights[y, x])
        kind = i % 4
        z = {0: rs.uniform(h + 2, dim - 2), 1: h + rs.uniform(1.01, 3.0), 2: rs.uniform(2, max(3, h)), 3: rs.uniform(h + 1, dim - 2)}[kind]
        pos = np.array([1 + (x + rs.rand()) / dim, 1 + (y + rs.rand()) / dim, 1 + z / dim], np.float32)
        if kind == 3:
            q = 1 << int(rs.randint(1, depth + 1))
            pos = np.clip((np.floor((pos - 1.0) * q) / q + 1.0), 1.0 + 1.0 / q, 2.0 - 1.0 / q).astype(np.float32)
        rot, fov = oc.camera_coeffs(float(rs.uniform(-3.14, 3.14)), float(rs.uniform(-1.55, 1.0)))
        W, H = [(640, 360), (1920, 1080), (3840, 2160), (320, 180)][i % 4]
        rows = 48
        y0 = int(rs.randint(0, H - rows)) & ~3
        k = emu.beam_level(pos, rot, fov, W, H, depth)
        if k == 0:
            continue
        if k not in grids:
            grids[k] = emu.beam_grid(nodes8, root, k)
        d = oc.gen_rays(rot, fov, W, H, y0, y0 + rows)
        want = oc.trace_rays(nodes8, root, depth, pos, d, rcp_tab=tab, nthreads=16)
        got = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, y0=y0, rows=rows, walker=13, want_stats=True, beam=grids[k], want_tau=True)
        try:
            assert_same_hits(got, want, f"depth {depth} cam {i}")
            hit = want[0] != 0
            assert (got[4][hit] <= want[2][hit]).all(), "tau > t_hit"
            assert got[3]["beam_guard"] == 0 and got[3]["beam_cert_wrong"] == 0
        except AssertionError as e:
            bad += 1
            print("FAIL", depth, i, pos.tolist(), W, H, y0, k, str(e)[:200], flush=True)
        n_frames += 1
    print(f"depth {depth}: {n_frames} frames checked, levels {sorted(grids)}", flush=True)
print("failures:", bad)
