"""Diagnostic (torchrun, 2+ ranks): NCCL send/recv bandwidth rank r -> rank 0 for a few message sizes, through
torch.distributed (the same libnccl the library's gather loads)."""
import json
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for mb in (8, 37, 128, 512):
    n = mb << 20
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(world)]
    def step():
        if rank == 0:
            ops = [dist.P2POp(dist.irecv, bufs[r], r) for r in range(1, world)]
        else:
            ops = [dist.P2POp(dist.isend, buf, 0)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    step(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        step()
    b.record(); torch.cuda.synchronize()
    out[f"{mb}MiB_per_sender"] = {"ms": round(a.elapsed_time(b) / 5, 3), "rank0_ingest_gbs": round((world - 1) * n * 5 / (a.elapsed_time(b) * 1e-3) / 1e9, 1)}
    dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "send_to_rank0": out}, indent=1))
dist.destroy_process_group()
