"""How does the frame kernel's time scale with the share of the frame a GPU gets?  (contiguous vs cyclic strips)"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import octree_ray_tracing_b200 as ort
from octree_ray_tracing_b200 import harness, multi_gpu
tree = ort.HOctree(24, 12, device=0, node_capacity=1 << 21); harness.build_terrain(tree); tree.sync(); ctx = tree.ctx
W, H = 3840, 2160
fv = torch.empty(W * H, dtype=torch.int32, device="cuda"); ff = torch.empty(W * H, dtype=torch.uint8, device="cuda"); ft = torch.empty(W * H, dtype=torch.float32, device="cuda")
fn = torch.empty(W * H, dtype=torch.int16, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
def timeit(fn_, reps=7):
    with torch.cuda.stream(stream):
        fn_(); fn_(); stream.synchronize(); evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn_(); b.record(stream); evs.append((a, b))
        stream.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
res = {}
for pn, (pos, yaw, pitch) in harness.POSES.items():
    rot, fov = ort.camera_coeffs(yaw, pitch); p = np.array(pos, np.float32)
    r = {}
    ctx.trace_frame_async(p, rot, fov, W, H, 0, H, 1, 1, fv, ff, ft, fn); ctx.sync()
    np_ = (fn.to(torch.int32) & 0xFFFF).reshape(H, W)
    r["max_push"] = int(np_.max().item()); r["mean_push"] = round(float(np_.float().mean().item()), 2)
    r["full"] = round(timeit(lambda: ctx.trace_frame_async(p, rot, fov, W, H, 0, H, 1, 1, fv, ff, ft)), 4)
    for world in (2, 4, 8):
        cyc = []; con = []
        for rank in range(world):
            y0, rows, _ = multi_gpu.strip_rows(rank, world, H, 8)
            cyc.append(timeit(lambda: ctx.trace_frame_async(p, rot, fov, W, H, y0, rows, 8, world, fv, ff, ft)))
            c0 = rank * (H // world)
            con.append(timeit(lambda: ctx.trace_frame_async(p, rot, fov, W, H, c0, H // world, 1, 1, fv, ff, ft)))
        r[f"cyclic_{world}"] = [round(x, 4) for x in cyc]; r[f"contig_{world}"] = [round(x, 4) for x in con]
    res[pn] = r
print(json.dumps(res, indent=1))
