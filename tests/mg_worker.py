"""torchrun worker of tests/test_gpu_configs.py::test_multi_gpu_gather_in_the_library_vs_reference (one process per GPU).
Rank 0 owns the host table; the DAG and its edit deltas reach the other ranks through ort_mg_broadcast_update, frames are
traced in cyclic strips and gathered by ort_mg_trace_frame_gather; rank 0 (and once rank 1) checks every assembled frame
against the oracle, row for row, bitwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import octree_ray_tracing_b200 as ort  # noqa: E402
from octree_ray_tracing_b200 import harness, multi_gpu  # noqa: E402
from oracle import oracle as oc  # noqa: E402  (checker)
from conftest import assert_same_hits  # noqa: E402
from test_oracle import _builtin_table  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
depth = 8
ctx = ort.TraceContext(depth, device=local)
mg = multi_gpu.MultiGpu(ctx)
assert mg.rank == rank and mg.world == world
tree = None
if rank == 0:
    tree = ort.HOctree(19, depth, device=None)
    harness.build_terrain(tree, tunnels=True)
tab = _builtin_table()


def ship():
    mg.broadcast_update(tree.take_delta() if rank == 0 else None)
    if rank == 0:
        nodes8, root, _ = tree.flatten()
        assert ctx.node_count >= nodes8.shape[0] and ctx.root == root
        return nodes8, root
    return None, None


def check_frames(nodes8, root, what):
    for (W, H, tile_rows) in ((1280, 720, 8), (640, 1001, 8), (320, 200, 16)):
        for dst in ((0, 1) if (W, H) == (1280, 720) else (0,)):
            bufs = []
            for p, (pos, yaw, pitch) in harness.POSES.items():
                rot, fov = ort.camera_coeffs(yaw, pitch)
                out = None
                if rank == dst:
                    out = (torch.full((W * H,), -1, dtype=torch.int32, device="cuda"), torch.full((W * H,), 255, dtype=torch.uint8, device="cuda"),
                           torch.full((W * H,), -1.0, dtype=torch.float32, device="cuda"))
                mg.trace_frame_gather(np.array(pos, np.float32), rot, fov, W, H, tile_rows=tile_rows, dst=dst, d_vox=out and out[0], d_face=out and out[1], d_t=out and out[2])
                bufs.append(out)
            mg.sync()
            # the consumer checks; for dst != 0 it gets the DAG from rank 0 first (plain torch broadcast of the checker's input)
            meta = torch.zeros(2, dtype=torch.int64, device="cuda")
            if rank == 0:
                meta[0], meta[1] = nodes8.shape[0], root
            dist.broadcast(meta, src=0)
            nn = torch.from_numpy(nodes8.view(np.int32)).cuda() if rank == 0 else torch.empty((int(meta[0]), 8), dtype=torch.int32, device="cuda")
            dist.broadcast(nn, src=0)
            if rank == dst:
                n8, rt = nn.cpu().numpy().view(np.uint32), int(meta[1])
                for (p, (pos, yaw, pitch)), out in zip(harness.POSES.items(), bufs):
                    rot, fov = oc.camera_coeffs(yaw, pitch)
                    d = oc.gen_rays(rot, fov, W, H)
                    want = oc.trace_rays(n8, rt, depth, np.array(pos, np.float32), d, rcp_tab=tab, nthreads=8)
                    got = (out[0].cpu().numpy().view(np.uint32), out[1].cpu().numpy(), out[2].cpu().numpy())
                    assert_same_hits(got, want, f"{what}: {W}x{H} tile_rows {tile_rows} pose {p} gathered on rank {dst} of {world}")
                    assert (got[0] != 0).sum() > 100


nodes8, root = ship()
check_frames(nodes8, root, "full upload")
transports = {mg.transport}
assert os.environ.get("ORT_MG_TRANSPORT") is not None or mg.transport in (0, 1)
for k in range(2):
    if rank == 0:
        tree.set_box(100 + 20 * k, 128, 70, 17 + k, 1 + k)
    nodes8, root = ship()
    mg.set_group((2, 5)[k])           # grouped wire operations: 2 frames per NCCL group (one partial group per sync), then 5 (every sync sends a partial group)
    check_frames(nodes8, root, f"delta {k}, group {(2, 5)[k]}")
# the other transport (NCCL send / recv where the default was peer copies), grouped and per frame
mg.set_transport(0)
mg.set_group(3)
check_frames(nodes8, root, "NCCL send/recv transport, group 3")
transports.add(mg.transport)
assert mg.transport == 0
mg.set_group(1)
check_frames(nodes8, root, "NCCL send/recv transport, group 1")
mg.set_transport(1)
check_frames(nodes8, root, "peer-copy transport again, group 1")
transports.add(mg.transport)
wire = mg.wire_bytes
dist.barrier()
if rank == 0:
    print(f"MG-OK world={world} nccl={ort.lib().ort_mg_nccl_version()} transports_used={sorted(transports)} wire_bytes_rank0={wire:.0f}")
mg.close()
dist.destroy_process_group()
