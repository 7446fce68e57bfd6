"""Multi-GPU plumbing of the headless harness: one process per GPU (torch.distributed), the DAG replicated,
frames cut into cyclic tile strips.  The trace itself needs no collective; NCCL is used for exactly two things,
as in BASELINE.json's north_star: broadcasting the DAG / its edit deltas from the rank that owns the host table,
and gathering finished strips where the assembled frame is wanted.

Everything here is backend-agnostic (nccl with CUDA tensors on the GPU box, gloo with CPU tensors in the CPU
tests): the functions take an `apply` callback instead of touching a device themselves.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


# ---- host placement ------------------------------------------------------------------------------------

def bind_to_gpu_numa(physical_gpu_index: int) -> bool:
    """Pin the calling process to the CPUs next to its GPU (NVML's ideal-CPU mask), so that the pinned host buffers it
    allocates afterwards land on that socket's memory and the GPU's D2H writes do not cross the inter-socket link.
    With 8 ranks each moving 9 B/ray to the host this is the difference between a per-GPU PCIe bound and a shared
    inter-socket bound.  Returns False (and changes nothing) where NVML or the affinity call is unavailable."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(physical_gpu_index)
        nv.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


# ---- partition -----------------------------------------------------------------------------------------

def strip_rows(rank: int, world: int, H: int, tile_rows: int = 8):
    """Rank's share of a frame of H rows cut into cyclic tile strips: tiles rank, rank+world, ...
    Returns (y0, rows, frame_rows) with frame_rows[r] = frame row of local row r -- the arguments
    ort_trace_frame wants are (y0, rows, tile_rows, tile_step=world).  H need not divide evenly: the last,
    shorter tile belongs to whoever its index falls to."""
    n_tiles = (H + tile_rows - 1) // tile_rows
    mine = np.arange(rank, n_tiles, world)
    frame_rows = (mine[:, None] * tile_rows + np.arange(tile_rows)[None, :]).ravel()
    frame_rows = frame_rows[frame_rows < H]
    return rank * tile_rows, int(frame_rows.size), frame_rows


def max_strip_rows(world: int, H: int, tile_rows: int = 8) -> int:
    return max(strip_rows(r, world, H, tile_rows)[1] for r in range(world))


def assemble(parts, world: int, H: int, W: int, tile_rows: int = 8):
    """parts[r] = rank r's strip (rows_r * W entries, possibly padded at the end) -> the H x W frame."""
    out = None
    for r in range(world):
        _, rows, frame_rows = strip_rows(r, world, H, tile_rows)
        p = parts[r][: rows * W].reshape(rows, W)
        if out is None:
            out = torch.empty((H, W), dtype=p.dtype, device=p.device) if isinstance(p, torch.Tensor) else np.empty((H, W), p.dtype)
        idx = torch.as_tensor(frame_rows, device=p.device) if isinstance(p, torch.Tensor) else frame_rows
        out[idx] = p
    return out


def gather_strips(local: torch.Tensor, world: int, H: int, W: int, tile_rows: int = 8, dst: int = 0):
    """Gather every rank's strip to `dst` and assemble the frame there (None elsewhere).  `local` holds this
    rank's rows * W results; strips are padded to the longest one so the collective is regular."""
    if world == 1:
        return local.reshape(H, W)
    rank = dist.get_rank()
    pad = max_strip_rows(world, H, tile_rows) * W
    send = local if local.numel() == pad else torch.cat([local, local.new_zeros(pad - local.numel())])
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst)
    return assemble(bufs, world, H, W, tile_rows) if rank == dst else None


# ---- DAG / delta broadcast -----------------------------------------------------------------------------

def broadcast_update(update, apply, device="cpu", src: int = 0):
    """Ship one device update from `src` to every rank and apply it everywhere.

    update (on src; ignored elsewhere): (ids | None, nodes8[n,8], root, is_full) as HOctree.take_delta()
    returns it.  apply(ids_or_None, nodes8, root, is_full) is called on every rank (src included) with
    tensors on `device` -- e.g. ctx.upload_full / ctx.upload_delta with device pointers.
    One header broadcast + one payload broadcast: [ids[n] | nodes8[n*8]] as uint32 (int32 on the wire)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    hdr = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == src:
        ids, nodes8, root, is_full = update
        n = int(nodes8.shape[0])
        hdr[0], hdr[1], hdr[2] = n, int(root), int(bool(is_full))
    if world > 1:
        dist.broadcast(hdr, src=src)
    n, root, is_full = int(hdr[0]), int(hdr[1]), bool(hdr[2])
    words = n * 8 + (0 if is_full else n)
    if rank == src:
        parts = [] if is_full else [np.ascontiguousarray(ids, np.uint32)]
        parts.append(np.ascontiguousarray(nodes8, np.uint32).reshape(-1))
        payload = torch.from_numpy(np.concatenate(parts).view(np.int32)).to(device) if words else torch.zeros(0, dtype=torch.int32, device=device)
    else:
        payload = torch.empty(words, dtype=torch.int32, device=device)
    if world > 1 and words:
        dist.broadcast(payload, src=src)
    if payload.is_cuda:
        # apply() hands raw pointers to the context, which copies on ITS stream: the payload must be complete first
        # (the NCCL broadcast and the host-to-device copy above only order torch's current stream, not the host)
        torch.cuda.current_stream(payload.device).synchronize()
    if is_full:
        apply(None, payload.view(-1, 8), root, True)
    else:
        apply(payload[:n], payload[n:].view(-1, 8), root, False)
    return n, is_full


def context_applier(ctx):
    """apply-callback that feeds a TraceContext from CUDA (or host) tensors without a host round trip."""
    def apply(ids, nodes8, root, is_full):
        lib = ctx.L
        from ._lib import check
        if is_full:
            check(lib.ort_upload_full(ctx.h, nodes8.data_ptr(), nodes8.shape[0], root), ctx.h)
        else:
            n = nodes8.shape[0]
            check(lib.ort_upload_delta(ctx.h, ids.data_ptr() if n else None, nodes8.data_ptr() if n else None, n, root), ctx.h)
    return apply
