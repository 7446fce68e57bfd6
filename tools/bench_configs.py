"""The BASELINE.json configurations that bench.py does not cover (it runs config 2):
  --config 4   edit-then-trace loop: 100 frames at 1920x1080, every 2nd frame a 40^3 place/remove (T/Z) edit at the
               crosshair hit, device mirror updated by deltas; host edit / delta upload / trace time per frame
  --config 5   depth-14 (16384^3) terrain DAG, 7680x4320 frames cut into cyclic tile strips over the ranks
               (launch with torchrun for N > 1), NCCL gather of the strips to rank 0
One JSON line per config on rank 0.  Device-resident timing with CUDA events; see bench.py for the headline metric."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import octree_ray_tracing_b200 as ort  # noqa: E402
from octree_ray_tracing_b200 import harness, multi_gpu  # noqa: E402


def config4(args):
    depth, log2cap = args.depth, {8: 19, 10: 22, 12: 24}[args.depth]
    W, H = 1920, 1080
    tree = ort.HOctree(log2cap, depth, device=0, node_capacity=1 << 21)
    heights, _ = harness.build_terrain(tree)
    dim = 1 << depth
    n0, _ = tree.sync()
    ctx = tree.ctx
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    # camera 0.2 * (terrain amplitude) above the ground at the map centre, looking steeply down (t < 0.5 rule holds)
    pos = np.array([1.5, 1.5, 1.0 + (heights[dim // 2, dim // 2] + 0.05 * dim) / dim], np.float32)
    yaw, pitch = 0.4, -1.2
    rot, fov = ort.camera_coeffs(yaw, pitch)
    dir3 = np.array([np.cos(np.float32(yaw)) * np.cos(np.float32(pitch)), np.sin(np.float32(yaw)) * np.cos(np.float32(pitch)), np.sin(np.float32(pitch))], np.float32)
    dv = torch.empty(W * H, dtype=torch.int32, device="cuda")
    df = torch.empty(W * H, dtype=torch.uint8, device="cuda")
    dt = torch.empty(W * H, dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream)
    t_edit = t_sync = t_trace = 0.0
    n_edits = delta_nodes = fulls = 0
    for frame in range(args.frames):
        face, vox, t = tree.sse_trace(pos, dir3)                       # pick ray (test_och_h_octree.cpp:535)
        if frame % 2 == 1 and vox and t < 0.5 and not args.no_edits:
            place = (frame // 2) % 2 == 0
            off = np.zeros(3, np.float32)
            if int(face) < 6:
                off[int(face) % 3] = (tree.voxel_dim / 2) * (1 if int(face) < 3 else -1)
            cp = pos + dir3 * np.float32(t) + (off if place else -off) - np.float32(1.0)
            c = (cp * np.float32(dim)).astype(np.uint16)
            a = time.perf_counter()
            if args.bulk:
                ci = c.astype(np.int64)
                tree.fill_box(ci - 20, ci + 20, 1 if place else 0)
            else:
                tree.set_box(int(c[0]), int(c[1]), int(c[2]), 40, 1 if place else 0)
            t_edit += time.perf_counter() - a
            n_edits += 1
        a = time.perf_counter()
        n, full = tree.sync()
        t_sync += time.perf_counter() - a
        delta_nodes += 0 if full else n
        fulls += full
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.trace_frame_async(pos, rot, fov, W, H, 0, H, 1, 1, dv, df, dt)
            e1.record(stream)
            stream.synchronize()
            t_trace += e0.elapsed_time(e1) * 1e-3
    print(json.dumps({"config": 4, "depth": depth, "frames": args.frames, "resolution": [W, H], "edits": n_edits, "edit_call": "ort_tree_fill_box (bulk)" if args.bulk else "ort_tree_set_box (64000 x set)",
                      "ms_per_edit_host_40^3": round(t_edit / max(n_edits, 1) * 1e3, 2),
                      "ms_per_sync_delta_build_upload": round(t_sync / args.frames * 1e3, 4),
                      "delta_nodes_per_edit": round(delta_nodes / max(n_edits, 1), 1), "full_uploads_after_first": fulls,
                      "ms_per_frame_trace": round(t_trace / args.frames * 1e3, 4),
                      "Mrays_per_s_trace": round(W * H * args.frames / t_trace / 1e6, 1),
                      "options": args.opt, "beam_level": ctx.beam_level(pos, rot, fov, W, H), "beam_grids_built": ctx.beam_builds,
                      "dag_nodes": tree.get_fillcnt(), "hits_last_frame": int((dv != 0).sum().item())}), flush=True)


def config5(args):
    rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    depth, log2cap = args.depth, {12: 24, 13: 26, 14: 27}[args.depth]
    W, H, tr = 7680, 4320, 8
    ctx = ort.TraceContext(depth, device=local, node_capacity=1 << 16)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    t0 = time.time()
    update = None
    if rank == 0:
        tree = ort.HOctree(log2cap, depth, device=None)
        harness.build_terrain(tree)
        update = tree.take_delta()
    t_build = time.time() - t0
    n_nodes, _ = multi_gpu.broadcast_update(update, multi_gpu.context_applier(ctx), device=torch.device("cuda", local))
    ctx.sync()
    if world > 1:                                   # all replicas must trace the same picture before anything is timed
        p0 = harness.POSES["B"]
        rot0, fov0 = ort.camera_coeffs(p0[1], p0[2])
        v, f, t = ctx.trace_frame(np.array(p0[0], np.float32), rot0, fov0, 480, 270)
        mine = torch.tensor([int(v.astype(np.uint64).sum()), int(f.astype(np.uint64).sum()), int((v != 0).sum())], dtype=torch.int64, device="cuda")
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi) or int(mine[2]) < 1000:
            raise SystemExit(f"[rank {rank}] replica check FAILED: ranks do not trace the same DAG")
    y0, rows, _ = multi_gpu.strip_rows(rank, world, H, tr)
    n_local = rows * W
    dv = torch.empty(n_local, dtype=torch.int32, device="cuda")
    df = torch.empty(n_local, dtype=torch.uint8, device="cuda")
    dt = torch.empty(n_local, dtype=torch.float32, device="cuda")
    dn = torch.empty(n_local, dtype=torch.int16, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream)
    cams = [(np.array(p[0], np.float32),) + ort.camera_coeffs(p[1], p[2]) for p in harness.POSES.values()]
    out = {}
    with torch.cuda.stream(stream):
        pushes = 0
        for cam in cams:
            ctx.trace_frame_async(cam[0], cam[1], cam[2], W, H, y0, rows, tr, world, dv, df, dt, dn)
            stream.synchronize()
            pushes += int((dn.to(torch.int64) & 0xFFFF).sum().item())
    # frames in flight on NS streams (own output buffers each) so that launch tails overlap, as in bench.py
    NS = args.streams
    streams = [torch.cuda.Stream(device=local) for _ in range(NS)]
    outs = [(torch.empty(n_local, dtype=torch.int32, device="cuda"), torch.empty(n_local, dtype=torch.uint8, device="cuda"),
             torch.empty(n_local, dtype=torch.float32, device="cuda")) for _ in range(NS)]
    torch.cuda.synchronize()
    for name, gather in (("trace", False), ("trace+gather", True)):
        for rep in range(2 + args.steps):
            if rep == 2:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                a = time.perf_counter()
            for k, cam in enumerate(cams):
                st = streams[k % NS]
                o = outs[k % NS]
                ctx.set_stream(st)
                ctx.trace_frame_async(cam[0], cam[1], cam[2], W, H, y0, rows, tr, world, o[0], o[1], o[2])
                if gather:
                    with torch.cuda.stream(st):
                        for buf in (o[0], o[2], o[1]):
                            multi_gpu.gather_strips(buf, world, H, W, tr, dst=0)
        ctx.set_stream(None)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sec = (time.perf_counter() - a) / args.steps
        tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out[name] = float(tt.item())
    pt = torch.tensor([float(pushes), float(n_local * len(cams))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(pt)
    if rank == 0:
        rays = W * H * len(cams)
        print(json.dumps({"config": 5, "depth": depth, "n_gpus": world, "resolution": [W, H], "frames_per_step": len(cams),
                          "dag_nodes": n_nodes, "dag_mib": round(n_nodes * 32 / 2**20, 1), "host_build_s": round(t_build, 1),
                          "pushes_per_ray": round(float(pt[0] / pt[1]), 2), "options": args.opt,
                          "beam_level": ctx.beam_level(cams[0][0], cams[0][1], cams[0][2], W, H),
                          "Mrays_per_s_trace": round(rays / out["trace"] / 1e6, 1),
                          "Mrays_per_s_trace_plus_gather": round(rays / out["trace+gather"] / 1e6, 1),
                          "ms_per_frame_trace": round(out["trace"] / len(cams) * 1e3, 3),
                          "timing": "wall clock around queued launches + final synchronize, max over ranks (strong scaling: one 8K frame split over all ranks)"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--streams", type=int, default=3, help="config 5: frames in flight")
    ap.add_argument("--bulk", action="store_true", help="config 4: use the bulk box edit instead of the 64000-set() loop")
    ap.add_argument("--opt", action="append", default=[], help="key=value passed to ort_set_option (e.g. beam=0)")
    ap.add_argument("--no-edits", action="store_true", help="config 4: the same frames without the edits (a steady scene)")
    args = ap.parse_args()
    if args.depth is None:
        args.depth = 10 if args.config == 4 else 14
    (config4 if args.config == 4 else config5)(args)
