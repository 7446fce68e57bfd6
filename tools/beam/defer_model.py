"""Offline model (CPU, host emulation): would handing rays that exceed a round budget to a second, densely packed launch save
warp-rounds?  Depth 12, 4K, poses A/B/C, beam start on.  Answer: no.  python tools/beam/defer_model.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'host_emu'))
import numpy as np, emu
import octree_ray_tracing_b200 as ort
from octree_ray_tracing_b200 import harness
from oracle import oracle as oc
DEPTH=12; W,H=3840,2160
tree = ort.HOctree(24, DEPTH, device=None); harness.build_terrain(tree)
nodes8, root, _ = tree.flatten()
grid = emu.beam_grid(nodes8, root, 7)
for name in "ABC":
    pos, yaw, pitch = harness.POSES[name]
    rot, fov = oc.camera_coeffs(yaw, pitch)
    got = emu.trace_frame(nodes8, root, DEPTH, pos, rot, fov, W, H, walker=13, want_npush=True, beam=grid)
    n = got[3].astype(np.int64).reshape(H//4,4,W//8,8).transpose(0,2,1,3).reshape(-1,32)   # per warp lanes
    wr = n.max(axis=1)
    tot_wr = wr.sum(); lane = n.sum()
    print(f"pose {name}: warps {len(n)}, active warps {(wr>0).sum()}, warp-rounds {tot_wr/1e6:.2f} M, lane-rounds {lane/1e6:.2f} M, efficiency {lane/(32*tot_wr):.3f}")
    for T in (48, 64, 96, 128, 192, 256):
        cut = np.minimum(wr, T).sum()                      # phase-1 warp rounds
        rem = np.maximum(n - T, 0)
        nd = (rem > 0).sum()
        # phase 2: deferred rays sorted by remaining length, packed 32 per warp
        r = np.sort(rem[rem > 0])[::-1]
        pad = (-len(r)) % 32
        r2 = np.concatenate([r, np.zeros(pad, np.int64)]).reshape(-1, 32)
        p2 = r2.max(axis=1).sum()
        # unsorted packing (arrival order ~ random): estimate with random permutation
        rp = np.random.RandomState(0).permutation(rem[rem > 0]); rp = np.concatenate([rp, np.zeros(pad, np.int64)]).reshape(-1, 32)
        p2r = rp.max(axis=1).sum()
        print(f"   T={T:3d}: phase-1 warp-rounds {cut/1e6:.2f} M ({cut/tot_wr:.3f}), deferred rays {nd} ({nd/n.size*100:.2f} %), phase-2 warp-rounds sorted {p2/1e6:.3f} M / unsorted {p2r/1e6:.3f} M, total {(cut+p2r)/tot_wr:.3f}, longest phase-1 warp {min(T, wr.max())}, longest deferred {r[0] if len(r) else 0}")
