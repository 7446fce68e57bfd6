"""World-size-2 tests of the multi-GPU plumbing on CPU (gloo): strip partition, DAG / delta broadcast from the
rank that owns the host table, strip gather.  The GPU is replaced by the oracle tracing each rank's mirror."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_strip_partition_covers_every_row_once():
    sys.path.insert(0, ROOT)
    from octree_ray_tracing_b200 import multi_gpu
    for H, tr in ((2160, 8), (4320, 8), (723, 8), (100, 16), (7, 8)):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(H, int)
            for r in range(world):
                y0, rows, fr = multi_gpu.strip_rows(r, world, H, tr)
                assert rows == fr.size and (rows == 0 or y0 == fr[0])
                # the kernel's mapping (ort_trace_frame): local row -> frame row
                loc = np.arange(rows)
                assert np.array_equal(y0 + (loc // tr) * tr * world + loc % tr, fr)
                seen[fr] += 1
            assert (seen == 1).all()


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path.insert(0, ROOT)
        import octree_ray_tracing_b200 as ort
        from octree_ray_tracing_b200 import harness, multi_gpu
        from oracle import oracle as oc

        depth, log2cap, W, H, tr = 6, 16, 96, 52, 8          # H not a multiple of the tile height
        mirror = {"nodes": np.zeros((0, 8), np.uint32), "root": 0}

        def apply(ids, nodes8, root, is_full):
            n8 = nodes8.numpy().view(np.uint32).reshape(-1, 8)
            if is_full:
                mirror["nodes"] = n8.copy()
            else:
                i = ids.numpy().view(np.uint32)
                need = int(i.max()) if i.size else 0
                if need > mirror["nodes"].shape[0]:
                    mirror["nodes"] = np.concatenate([mirror["nodes"], np.zeros((need - mirror["nodes"].shape[0], 8), np.uint32)])
                mirror["nodes"][i - 1] = n8
            mirror["root"] = root

        tree = A = None
        if rank == 0:
            tree = ort.HOctree(log2cap, depth, device=None)
            harness.build_terrain(tree)
            A = oc.OracleTree(log2cap, depth)
            A.initialize_terrain(oc.heightmap(depth), oc.grass_bits(depth), False)
        rot, fov = oc.camera_coeffs(0.4, -0.8)
        pos = np.array([1.5, 1.5, 1.7], np.float32)
        y0, rows, frame_rows = multi_gpu.strip_rows(rank, world, H, tr)
        rs = np.random.RandomState(5)
        kinds = []
        for step in range(6):
            update = None
            if rank == 0:
                if step:
                    cx, cy, cz, ext, v = 20 + 4 * step, 30, 20 + step, 9, step % 2
                    tree.set_box(cx, cy, cz, ext, v)
                    A.set_many(np.array([((cx + x) & 0xFFFF, (cy + y) & 0xFFFF, (cz + z) & 0xFFFF, v)
                                         for z in range(-4, 5) for y in range(-4, 5) for x in range(-4, 5)], np.uint32))
                update = tree.take_delta()
            n, full = multi_gpu.broadcast_update(update, apply)
            kinds.append(full)
            # every rank traces its strip over its own replica
            d = np.concatenate([oc.gen_rays(rot, fov, W, H, int(y), int(y) + 1) for y in frame_rows])
            v, f, t = oc.trace_rays(mirror["nodes"], mirror["root"], depth, pos, d)
            frame_v = multi_gpu.gather_strips(torch.from_numpy(v.view(np.int32)), world, H, W, tr)
            frame_t = multi_gpu.gather_strips(torch.from_numpy(t), world, H, W, tr)
            if rank == 0:
                wv, wf, wt = A.trace(pos, oc.gen_rays(rot, fov, W, H))
                assert np.array_equal(frame_v.numpy().view(np.uint32).ravel(), wv), f"step {step}: voxels"
                assert np.array_equal(frame_t.numpy().view(np.uint32).ravel(), wt.view(np.uint32)), f"step {step}: t"
        assert kinds[0] and not any(kinds[1:]), kinds       # one full upload, then deltas only
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
        raise


def test_broadcast_and_gather_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
