"""Design aid (CPU only): how do the lanes of a warp spend the rounds of the frame kernel?  Runs the device walker
compiled for the host (tests/host_emu) with the 32 rays of every 8x4 tile in lockstep and reports, per pose of the bench
step, the share of warp-rounds in which only the descend branch, only the advance branch or both have takers -- a warp
pays for each branch that has at least one lane in it -- and the rounds per tree level.
Usage: python tools/simt_model.py [depth] [W] [H]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "host_emu"))
import octree_ray_tracing_b200 as ort          # noqa: E402
import emu                                      # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 12
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
H = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
T = ort.HOctree({8: 19, 10: 22, 12: 24}.get(depth, 24), depth, device=None)
ort.harness.build_terrain(T, tunnels=False)
nodes8, root, _ = T.flatten()
for name, (pos, yaw, pitch) in ort.harness.POSES.items():
    rot, fov = ort.camera_coeffs(yaw, pitch)
    m = emu.warp_model(nodes8, root, depth, pos, rot, fov, W, H)
    st = emu.trace_frame(nodes8, root, depth, pos, rot, fov, W, H, walker=1, want_stats=True)[3]
    wr = m["warp_rounds"]
    print(f"pose {name}: {m['lane_rounds'] / (W * H):.2f} rounds per ray, {wr / m['warps']:.1f} per warp; "
          f"descend only {100 * m['descend_only'] / wr:.1f} %, advance only {100 * m['advance_only'] / wr:.1f} %, both {100 * m['both'] / wr:.1f} %; "
          f"{m['active_lanes'] / wr:.1f} lanes active per round ({m['descend_lanes'] / wr:.1f} descend, {m['advance_lanes'] / wr:.1f} advance)")
    print("   rounds per ray by level:", np.round(st["rounds_by_level"][1:depth + 1] / (W * H), 2))
