"""BASELINE config 3: depth-12 terrain, 16 M incoherent random-direction rays, device-resident timing of
ort_trace_rays_async for the one-thread-per-ray kernel vs the persistent lane-refill kernel.
Also times the frame kernels per variant.  Usage: python tools/bench_rays.py [n_rays_log2]"""
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

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
depth, log2cap = 12, 24
tree = ort.HOctree(log2cap, depth, device=0, node_capacity=1 << 21)
harness.build_terrain(tree)
tree.sync()
ctx = tree.ctx
n = 1 << log2n
o, d = harness.random_rays(n)
do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
dv = torch.empty(n, dtype=torch.int32, device="cuda")
df = torch.empty(n, dtype=torch.uint8, device="cuda")
dt = torch.empty(n, dtype=torch.float32, device="cuda")
dn = torch.empty(n, dtype=torch.int16, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
torch.cuda.synchronize()


def timeit(fn, reps=5):
    with torch.cuda.stream(stream):
        fn(); fn()
        stream.synchronize()
        evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            evs.append((a, b))
        stream.synchronize()
    return min(a.elapsed_time(b) for a, b in evs)


ctx.trace_rays_async(do, 3, dd, n, dv, df, dt, dn)
ctx.sync()
pushes = float((dn.to(torch.int64) & 0xFFFF).sum().item()) / n
ref = (dv.clone(), df.clone(), dt.clone())
res = {"n_rays": n, "pushes_per_ray": round(pushes, 3), "hit_fraction": round(float((dv != 0).float().mean().item()), 4), "rays": {}, "frames": {}}
for pb in (1, 8):
    ctx.set_option("rays_variant", 2)
    ctx.set_option("persist_blocks", pb)
    for lw in (16, 20, 24):
        ctx.set_option("low_water", lw)
        ms = timeit(lambda: ctx.trace_rays_async(do, 3, dd, n, dv, df, dt))
        ctx.sync()
        same = bool((dv == ref[0]).all().item() and (df == ref[1]).all().item() and (dt.view(torch.int32) == ref[2].view(torch.int32)).all().item())
        res["rays"][f"variant2_blocks{pb}_lw{lw}"] = {"ms": round(ms, 3), "Mrays/s": round(n / ms / 1e3, 1), "same_as_ref": same}
ctx.set_option("persist_blocks", 6)
for rv, lws in ((1, (0,)), (2, (0, 8, 16, 20, 24, 28))):
    ctx.set_option("rays_variant", rv)
    for lw in lws:
        ctx.set_option("low_water", lw)
        ms = timeit(lambda: ctx.trace_rays_async(do, 3, dd, n, dv, df, dt))
        ctx.sync()
        same = bool((dv == ref[0]).all().item() and (df == ref[1]).all().item() and (dt.view(torch.int32) == ref[2].view(torch.int32)).all().item())
        res["rays"][f"variant{rv}_lw{lw}"] = {"ms": round(ms, 3), "Mrays/s": round(n / ms / 1e3, 1), "same_as_ref": same}
W, H = 3840, 2160
fv = torch.empty(W * H, dtype=torch.int32, device="cuda")
ff = torch.empty(W * H, dtype=torch.uint8, device="cuda")
ft = torch.empty(W * H, dtype=torch.float32, device="cuda")
for variant, lws in ((0, (20,)), (1, (20,)), (2, (8, 16, 24))):
    ctx.set_option("variant", variant)
    for lw in lws:
        ctx.set_option("low_water", lw)
        out = {}
        for pn, (pos, yaw, pitch) in harness.POSES.items():
            rot, fov = ort.camera_coeffs(yaw, pitch)
            p = np.array(pos, np.float32)
            out[pn] = round(timeit(lambda: ctx.trace_frame_async(p, rot, fov, W, H, 0, H, 1, 1, fv, ff, ft)), 4)
        res["frames"][f"variant{variant}_lw{lw}"] = out
print(json.dumps(res, indent=1))
