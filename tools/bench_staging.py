"""Does staging the DAG's upper levels in shared memory pay?  Frame kernel variant 1 (global/L1 only) against the
staged experiment kernel (variant 3) for several numbers of staged levels, depth-12 terrain, 4K, poses A/B/C."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import octree_ray_tracing_b200 as ort
from octree_ray_tracing_b200 import harness
tree = ort.HOctree(24, 12, device=0, node_capacity=1 << 21); harness.build_terrain(tree)
_, _, lo = tree.flatten()
tree.sync(); ctx = tree.ctx
W, H = 3840, 2160
bufs = [torch.empty(W * H, dtype=torch.int32, device="cuda"), torch.empty(W * H, dtype=torch.uint8, device="cuda"), torch.empty(W * H, dtype=torch.float32, device="cuda")]
ref = [b.clone() for b in bufs]
stream = torch.cuda.ExternalStream(ctx.stream)
def timeit(fn_, reps=7):
    with torch.cuda.stream(stream):
        fn_(); fn_(); stream.synchronize(); evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn_(); b.record(stream); evs.append((a, b))
        stream.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
res = {"level_first_ids": [int(x) for x in lo]}
for label, variant, levels in (("global_only_256thr", 1, 0), ("staged_1024thr_0_levels", 3, 0), ("staged_3_levels", 3, 3), ("staged_4_levels", 3, 4),
                               ("staged_5_levels", 3, 5), ("staged_6_levels", 3, 6), ("staged_7_levels(capped)", 3, 7)):
    ctx.set_option("variant", variant)
    n_staged = int(lo[levels]) - 1 if levels else 0
    ctx.set_option("smem_levels", n_staged)
    out = {"n_staged_nodes": min(n_staged, 6144), "smem_kb": round(min(n_staged, 6144) * 32 / 1024, 1)}
    for pn, (pos, yaw, pitch) in harness.POSES.items():
        rot, fov = ort.camera_coeffs(yaw, pitch); p = np.array(pos, np.float32)
        out[pn] = round(timeit(lambda: ctx.trace_frame_async(p, rot, fov, W, H, 0, H, 1, 1, *bufs)), 4)
        if variant == 1:
            ctx.sync(); ref = [b.clone() for b in bufs]
    ctx.sync()
    out["same_as_global_only(pose C)"] = bool(all((a.view(torch.uint8) == b.view(torch.uint8)).all().item() for a, b in zip(bufs, ref)))
    res[label] = out
print(json.dumps(res, indent=1))
