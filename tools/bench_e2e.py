"""End-to-end (host-buffer) frame paths on the depth-12 terrain at 3840x2160, poses A/B/C:
   staged  = chunked kernel -> device staging -> copy engine (the default of ort_trace_frame with host outputs)
   zero    = option zero_copy: the kernel stores straight into the mapped pinned host buffers
for the 9-byte outputs (voxel, face, t) and for shaded RGBA frames.  Wall clock around the public calls, results of
the two modes compared.  Usage: python tools/bench_e2e.py [reps]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import octree_ray_tracing_b200 as ort  # noqa: E402
from octree_ray_tracing_b200 import harness  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
depth, log2cap = 12, 24
W, H = 3840, 2160
n = W * H
tree = ort.HOctree(log2cap, depth, device=0, node_capacity=1 << 21)
harness.build_terrain(tree)
tree.sync()
ctx = tree.ctx
cols, _ = harness.parse_voxels(harness.DEMO_VOXELS)
ctx.set_palette(cols)
cams = [(np.array(p[0], np.float32),) + ort.camera_coeffs(p[1], p[2]) for p in (harness.POSES[k] for k in "ABC")]

hv = torch.empty(n, dtype=torch.int32).pin_memory()
hf = torch.empty(n, dtype=torch.uint8).pin_memory()
ht = torch.empty(n, dtype=torch.float32).pin_memory()
hrgba = torch.empty(n, dtype=torch.int32).pin_memory()
out = (hv.numpy().view(np.uint32), hf.numpy(), ht.numpy(), None)


def run9():
    for cam in cams:
        ctx.trace_frame(cam[0], cam[1], cam[2], W, H, out=out)


def run4():
    for cam in cams:
        ctx.trace_frame_rgba(cam[0], cam[1], cam[2], W, H, out=hrgba)


def wall(fn):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


res = {"rays_per_run": 3 * n, "modes": {}}
keep = {}
for variant in (1,):
    ctx.set_option("variant", variant)
    for zc in (0, 1):
        ctx.set_option("zero_copy", zc)
        s9 = wall(run9)
        keep[(variant, zc, 9)] = tuple(x.copy() for x in out[:3])
        s4 = wall(run4)
        keep[(variant, zc, 4)] = hrgba.numpy().copy()
        res["modes"][f"variant{variant}_{'zero_copy' if zc else 'staged'}"] = {
            "out9_ms_per_frame": round(s9 / 3 * 1e3, 4), "out9_Mrays_s": round(3 * n / s9 / 1e6, 1),
            "rgba_ms_per_frame": round(s4 / 3 * 1e3, 4), "rgba_Mrays_s": round(3 * n / s4 / 1e6, 1)}
# deferred completion: the three frames of a run are enqueued back to back, ort_sync() collects them -- frame k+1's
# kernels overlap frame k's D2H.  Separate host buffers per frame.
ctx.set_option("variant", 1)
ctx.set_option("zero_copy", 0)
outs3 = [(torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory())
         for _ in cams]
outs3 = [(a.numpy().view(np.uint32), b.numpy(), c.numpy(), None) for a, b, c in outs3]
rgba3 = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in cams]


def run9_deferred():
    for cam, o in zip(cams, outs3):
        ctx.trace_frame(cam[0], cam[1], cam[2], W, H, out=o)
    ctx.sync()


def run4_deferred():
    for cam, o in zip(cams, rgba3):
        ctx.trace_frame_rgba(cam[0], cam[1], cam[2], W, H, out=o)
    ctx.sync()


ctx.set_option("defer_sync", 1)
for fc in (1, 2, 3, 6):
    ctx.set_option("frame_chunks", fc)
    s9, s4 = wall(run9_deferred), wall(run4_deferred)
    res["modes"][f"variant1_staged_deferred_sync_chunks{fc}"] = {
        "out9_ms_per_frame": round(s9 / 3 * 1e3, 4), "out9_Mrays_s": round(3 * n / s9 / 1e6, 1),
        "rgba_ms_per_frame": round(s4 / 3 * 1e3, 4), "rgba_Mrays_s": round(3 * n / s4 / 1e6, 1)}
ctx.set_option("frame_chunks", 0)
ctx.set_option("defer_sync", 0)
for fc in (3, 4, 8):
    ctx.set_option("frame_chunks", fc)
    s9, s4 = wall(run9), wall(run4)
    res["modes"][f"variant1_staged_chunks{fc}"] = {
        "out9_ms_per_frame": round(s9 / 3 * 1e3, 4), "out9_Mrays_s": round(3 * n / s9 / 1e6, 1),
        "rgba_ms_per_frame": round(s4 / 3 * 1e3, 4), "rgba_Mrays_s": round(3 * n / s4 / 1e6, 1)}
ctx.set_option("frame_chunks", 0)
keep[(1, 2, 9)] = tuple(x.copy() for x in outs3[2][:3])        # pose C, like the synchronous runs' last frame
keep[(1, 2, 4)] = rgba3[2].numpy().copy()

# explicit rays (config 3): 16.7 M incoherent rays from pinned host memory and back
nr = 1 << 24
o, d = harness.random_rays(nr)
ho, hd = torch.from_numpy(o).pin_memory(), torch.from_numpy(d).pin_memory()
rv = torch.empty(nr, dtype=torch.int32).pin_memory()
rf = torch.empty(nr, dtype=torch.uint8).pin_memory()
rt = torch.empty(nr, dtype=torch.float32).pin_memory()
L = ctx.L
from octree_ray_tracing_b200.tree import _p  # noqa: E402
res["rays_e2e"] = {"n": nr}
rref = None
for chunk in (nr, 1 << 21, 1 << 20):
    ctx.set_option("rays_chunk", chunk)
    sec = wall(lambda: ctx._ck(L.ort_trace_rays(ctx.h, _p(ho.numpy()), 3, _p(hd.numpy()), nr, _p(rv.numpy()), _p(rf.numpy()), _p(rt.numpy()), None)))
    got = (rv.numpy().copy(), rf.numpy().copy(), rt.numpy().view(np.uint32).copy())
    if rref is None:
        rref = got
    res["rays_e2e"][f"chunk_{chunk}"] = {"ms": round(sec * 1e3, 3), "Mrays_s": round(nr / sec / 1e6, 1),
                                         "same": bool(all(np.array_equal(a, b) for a, b in zip(got, rref)))}
ctx.set_option("rays_chunk", 1 << 21)

ref9, ref4 = keep[(1, 0, 9)], keep[(1, 0, 4)]
res["all_modes_identical"] = bool(all(
    (np.array_equal(v[0], ref9[0]) and np.array_equal(v[1], ref9[1]) and np.array_equal(v[2].view(np.uint32), ref9[2].view(np.uint32)))
    if k[2] == 9 else np.array_equal(v, ref4) for k, v in keep.items()))
print(json.dumps(res, indent=1))
