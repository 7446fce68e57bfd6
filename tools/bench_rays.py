"""BASELINE config 3: depth-12 terrain, 2^24 incoherent random-direction rays.  Device-resident timing of
ort_trace_rays_async for the one-thread-per-ray kernels (variants 0 / 1 / 13) and the persistent lane-refill kernel
(low-water sweep, register caps), every result compared with the ORACLE (the reference's own sse_trace where
oracle/_ref exists) on all rays: voxel, face and t bitwise.  Usage: python tools/bench_rays.py [n_rays_log2]"""
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
from oracle import oracle as oc  # noqa: E402  (checker only)

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

# the oracle's answer for every ray
nodes8, root, _ = tree.flatten()
t0 = time.time()
if oc.have_ref():
    R = oc.RefTree(log2cap, depth)
    R.import_compact(nodes8, root)
    wv, wf, wt = R.trace(o, d, nthreads=os.cpu_count())
    kind = "reference (oracle/_ref: the reference's own sse_trace)"
else:
    wv, wf, wt = oc.trace_rays(nodes8, root, depth, o, d, nthreads=os.cpu_count())
    kind = "oracle port"
oracle_s = time.time() - t0
ref = (torch.from_numpy(wv.view(np.int32)).cuda(), torch.from_numpy(wf).cuda(), torch.from_numpy(wt.view(np.int32)).cuda())


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


def mismatches():
    return {"voxel": int((dv != ref[0]).sum().item()), "face": int((df != ref[1]).sum().item()), "t_bitwise": int((dt.view(torch.int32) != ref[2]).sum().item())}


ctx.trace_rays_async(do, 3, dd, n, dv, df, dt, dn)
ctx.sync()
pushes = float((dn.to(torch.int64) & 0xFFFF).sum().item()) / n
res = {"n_rays": n, "pushes_per_ray": round(pushes, 3), "hit_fraction": round(float((dv != 0).float().mean().item()), 4),
       "oracle": {"kind": kind, "seconds": round(oracle_s, 2), "threads": os.cpu_count(), "Mrays/s": round(n / oracle_s / 1e6, 1)}, "rays": {}}


def run(name):
    dv.zero_(); df.zero_(); dt.zero_()
    ms = timeit(lambda: ctx.trace_rays_async(do, 3, dd, n, dv, df, dt))
    ctx.sync()
    res["rays"][name] = {"ms": round(ms, 3), "Mrays/s": round(n / ms / 1e3, 1), "mismatches_vs_oracle": mismatches()}


ctx.set_option("rays_variant", 1)
for v in (0, 1, 13):
    ctx.set_option("variant", v)
    run(f"one_thread_per_ray_variant{v}")
ctx.set_option("variant", 13)
ctx.set_option("rays_variant", 2)
for pb in (1, 6, 8):
    ctx.set_option("persist_blocks", pb)
    for lw in ((0, 8, 16, 20, 24, 28) if pb == 6 else (16, 20, 24)):
        ctx.set_option("low_water", lw)
        run(f"persistent_blocks{pb}_lw{lw}")
print(json.dumps(res, indent=1))
