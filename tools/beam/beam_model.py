"""Offline model of a conservative beam start (CPU only, statistics): for the bench frames (depth 12, 4K, poses A/B/C),
how many PUSH rounds of the reference walk lie before a tile-wide conservative start time t0 obtained from a dilated
level-k occupancy grid, and at which level the walk would be re-entered.  python tools/beam/beam_model.py [k] [tile_w tile_h]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import octree_ray_tracing_b200 as ort
from octree_ray_tracing_b200 import harness
from oracle import oracle as oc

K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
TW, TH = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 4)
DEPTH, W, H = 12, 3840, 2160
so = "/tmp/libbeam_model.so"
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, os.path.join(ROOT, "tools/beam/beam_model.c"), "-lm"])
L = C.CDLL(so)
p = lambda a: a.ctypes.data_as(C.c_void_p)

tree = ort.HOctree(24, DEPTH, device=None)
harness.build_terrain(tree)
nodes8, root, _ = tree.flatten()
nodes8 = np.ascontiguousarray(nodes8, np.uint32)

# occupancy of level k: ids[z, y, x] of the level-l cells, expanded level by level
ids = np.array([[[root]]], np.uint32)
for l in range(K):
    n = ids.shape[0]
    nxt = np.zeros((2 * n, 2 * n, 2 * n), np.uint32)
    ch = np.where(ids[..., None] != 0, nodes8[np.maximum(ids, 1) - 1], 0)          # (n, n, n, 8)
    for s in range(8):
        nxt[(s >> 2 & 1)::2, (s >> 1 & 1)::2, (s & 1)::2] = ch[..., s]
    ids = nxt
occ = ids != 0
N = 1 << K
print(f"level {K}: {occ.sum()} of {N**3} cells occupied")
dil = np.zeros_like(occ)
pad = np.pad(occ, 1)
for dz in range(3):
    for dy in range(3):
        for dx in range(3):
            dil |= pad[dz:dz + N, dy:dy + N, dx:dx + N]
print(f"dilated: {dil.sum()} cells")
dil8 = np.ascontiguousarray(dil, np.uint8)

step = 4       # every step-th tile in x and y
for name in "ABC":
    pos, yaw, pitch = harness.POSES[name]
    o = np.array(pos, np.float32)
    rot, fov = oc.camera_coeffs(yaw, pitch)
    tot_r = tot_after = tot_skip = tot_n = 0
    miss_free = 0
    worst = 0.0
    lv_hist = np.zeros(16, np.int64)
    for ty in range(0, H // TH, step):
        rows = oc.gen_rays(rot, fov, W, H, ty * TH, ty * TH + TH).reshape(TH, W, 3)
        txs = np.arange(0, W // TW, step)
        d = np.stack([rows[:, tx * TW:(tx + 1) * TW, :].reshape(-1, 3) for tx in txs])     # (tiles, TW*TH, 3)
        dc = d.astype(np.float64).mean(axis=1)
        dc /= np.linalg.norm(dc, axis=1, keepdims=True)
        dev = np.linalg.norm(d.astype(np.float64) / np.linalg.norm(d.astype(np.float64), axis=2, keepdims=True) - dc[:, None, :], axis=2).max(axis=1)
        te = np.zeros(len(txs), np.float64)
        L.beam_dda(p(dil8), K, p(o), p(np.ascontiguousarray(dc)), C.c_size_t(len(txs)), p(te))
        fin = np.isfinite(te)
        worst = max(worst, float((np.where(fin, te, 1.8) * (dev + 4e-4)).max() * N))       # beam radius at t0 in level-k cells (must stay < 1)
        t0 = np.repeat((te * (1 - 1e-3)).astype(np.float32), TW * TH)
        dd = np.ascontiguousarray(d.reshape(-1, 3), np.float32)
        n = len(dd)
        rounds = np.zeros(n, np.int32); at_r = np.zeros(n, np.int32); at_l = np.zeros(n, np.int32); th = np.zeros(n, np.float32)
        L.beam_walk(p(nodes8), C.c_uint32(root), DEPTH, p(o), p(dd), p(t0), C.c_size_t(n), p(rounds), p(at_r), p(at_l), p(th))
        assert not (np.isfinite(th) & (th < t0)).any(), "t0 not conservative"
        found = at_r >= 0
        # rays whose tile start lies beyond their walk: a MISS without any round
        new = np.where(found, at_l + (rounds - at_r - 1), 0)
        new = np.where(~found & np.isfinite(th), rounds, new)      # (hit before any exit time >= t0 cannot happen when conservative)
        miss_free += int((~found & ~np.isfinite(th)).sum())
        tot_r += int(rounds.sum()); tot_after += int(new.sum()); tot_n += n
        np.add.at(lv_hist, at_l[found], 1)
    print(f"pose {name}: rounds/ray {tot_r / tot_n:.2f} -> {tot_after / tot_n:.2f} ({tot_after / tot_r:.3f}); rays that become a MISS without a round: {miss_free / tot_n:.3f}; "
          f"beam radius at t0 <= {worst:.2f} level-{K} cells; re-entry level histogram {lv_hist[1:13].tolist()}")
