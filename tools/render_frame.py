"""Render the demo terrain from one of the survey's poses to a PNG (needs a B200): the headless stand-in for the
reference's window, for eyeball parity with its README screenshot.
Usage: python tools/render_frame.py [--depth 10] [--pose A|B|C] [--size 1280x720] [--tunnels] out.png"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import octree_ray_tracing_b200 as ort  # noqa: E402
from octree_ray_tracing_b200 import harness  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("out")
ap.add_argument("--depth", type=int, default=10)
ap.add_argument("--pose", default="A", choices=list(harness.POSES))
ap.add_argument("--size", default="1280x720")
ap.add_argument("--tunnels", action="store_true")
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
log2cap = {8: 19, 10: 22, 12: 24, 13: 26, 14: 27}[args.depth]
tree = ort.HOctree(log2cap, args.depth, device=0)
harness.build_terrain(tree, tunnels=args.tunnels)
tree.sync()
cols, names = harness.parse_voxels(harness.DEMO_VOXELS)
tree.ctx.set_palette(cols)
pos, yaw, pitch = harness.POSES[args.pose]
rot, fov = ort.camera_coeffs(yaw, pitch)
rgba = tree.ctx.trace_frame_rgba(np.array(pos, np.float32), rot, fov, W, H)
harness.save_png(args.out, rgba, W, H)
print(f"{args.out}: {W}x{H}, pose {args.pose}, depth {args.depth}, {tree.get_fillcnt()} nodes, voxel types {names}")
