"""Multi-GPU plumbing of the headless harness: one process per GPU (torch.distributed), the DAG replicated,
frames cut into cyclic tile strips.  The trace itself needs no collective; NCCL is used for exactly two things,
as in BASELINE.json's north_star: broadcasting the DAG / its edit deltas from the rank that owns the host table,
and gathering finished strips where the assembled frame is wanted.

Two layers:
  * `MultiGpu` -- the library's own communicator (include/ort_b200.h, ort_mg_*): NCCL inside libort_b200.so, strips
    traced into a ring of slots, sent on a dedicated stream, unpacked at their final rows on the consumer.  This is
    the product path on GPUs; torch.distributed only carries the 128-byte NCCL id at start-up.
  * the functions below it -- backend-agnostic helpers (nccl with CUDA tensors, gloo with CPU tensors in the CPU
    tests) that take an `apply` callback instead of touching a device themselves; `gather_strips` is the simple
    torch.distributed gather kept as the checker of the library's.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


# ---- the library's communicator ------------------------------------------------------------------------

class MultiGpu:
    """ort_mg_* (include/ort_b200.h): the communicator of one rank's TraceContext.  The NCCL id travels over the
    torch.distributed process group that launched the job (any backend); everything after that is the library's."""

    def __init__(self, ctx, rank: int | None = None, world: int | None = None):
        import ctypes as C
        from ._lib import check, lib
        self.L, self.ctx = lib(), ctx
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                raw = (C.c_char * 128)()
                check(self.L.ort_mg_unique_id(raw))
                idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
            dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = idbuf.to(dev)
            dist.broadcast(t, src=0)
            idbuf = t.cpu()
        self._id = bytes(idbuf.numpy().tobytes())
        h = C.c_void_p()
        check(self.L.ort_mg_create(C.byref(h), ctx.h, self.rank, self.world, self._id), ctx.h)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.ort_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        from ._lib import check
        check(rc, self.ctx.h)

    def strip_rows(self, H: int, tile_rows: int = 8, rank: int | None = None) -> int:
        return self.L.ort_mg_strip_rows(self.rank if rank is None else rank, self.world, H, tile_rows)

    def broadcast_update(self, update, src: int = 0):
        """update (on src; None elsewhere): (ids | None, nodes8, root, is_full) as HOctree.take_delta() returns it."""
        if self.rank == src:
            ids, nodes8, root, is_full = update
            nodes8 = np.ascontiguousarray(nodes8, np.uint32)
            ids = None if ids is None else np.ascontiguousarray(ids, np.uint32)
            n = int(nodes8.shape[0])
            self._ck(self.L.ort_mg_broadcast_update(self.h, None if ids is None else ids.ctypes.data, nodes8.ctypes.data if n else None, n, int(root), int(bool(is_full)), src))
        else:
            self._ck(self.L.ort_mg_broadcast_update(self.h, None, None, 0, 0, 0, src))
        return self.ctx.node_count

    def trace_frame_gather(self, pos, rot, fov_factor, W, H, tile_rows=8, dst=0, d_vox=None, d_face=None, d_t=None):
        """Enqueue: trace this rank's strips and gather the frame into dst's device tensors (None elsewhere)."""
        pos = np.ascontiguousarray(pos, np.float32)
        rot = np.ascontiguousarray(rot, np.float32)
        ptr = lambda x: None if x is None else x.data_ptr()
        self._ck(self.L.ort_mg_trace_frame_gather(self.h, pos.ctypes.data, rot.ctypes.data, float(fov_factor), W, H, tile_rows, dst, ptr(d_vox), ptr(d_face), ptr(d_t)))

    def make_jobs(self, jobs):
        """jobs: iterable of (pos, rot, fov_factor, W, H, tile_rows, dst, d_vox, d_face, d_t) -> a reusable job array for
        trace_frames_gather (the tensors must stay alive while the array is in use)."""
        import ctypes as C

        class Job(C.Structure):
            _fields_ = [("pos", C.c_float * 3), ("rot", C.c_float * 9), ("fov_factor", C.c_float), ("W", C.c_int), ("H", C.c_int), ("tile_rows", C.c_int), ("dst", C.c_int),
                        ("voxel", C.c_void_p), ("face", C.c_void_p), ("t", C.c_void_p)]
        jobs = list(jobs)
        arr = (Job * len(jobs))()
        for a, (pos, rot, fov, W, H, tile_rows, dst, dv, df, dt) in zip(arr, jobs):
            a.pos[:] = [float(x) for x in np.asarray(pos, np.float32).ravel()]
            a.rot[:] = [float(x) for x in np.asarray(rot, np.float32).ravel()]
            a.fov_factor, a.W, a.H, a.tile_rows, a.dst = float(fov), W, H, tile_rows, dst
            a.voxel = None if dv is None else dv.data_ptr()
            a.face = None if df is None else df.data_ptr()
            a.t = None if dt is None else dt.data_ptr()
        return arr

    def trace_frames_gather(self, job_array):
        """Enqueue a sequence of frames (make_jobs) in one call."""
        self._ck(self.L.ort_mg_trace_frames_gather(self.h, job_array, len(job_array)))

    def set_group(self, frames: int):
        """Frames per wire operation (the same on every rank): 1 = lowest latency, n = one NCCL group per n frames."""
        self._ck(self.L.ort_mg_set_group(self.h, frames))

    def flush(self):
        self._ck(self.L.ort_mg_flush(self.h))

    def set_trace_streams(self, n: int):
        """Streams the strips of consecutive frames are traced on (1..8; default 4, 8 beyond four ranks)."""
        self._ck(self.L.ort_mg_set_trace_streams(self.h, n))

    def set_transport(self, transport: int):
        """1: peer copies on the copy engines (CUDA IPC), 0: NCCL send / recv.  Collective; takes effect at the next frame."""
        self._ck(self.L.ort_mg_set_transport(self.h, transport))

    @property
    def transport(self) -> int:
        return self.L.ort_mg_transport(self.h)

    def sync(self):
        self._ck(self.L.ort_mg_sync(self.h))

    @property
    def stream(self):
        return self.L.ort_mg_stream(self.h)

    @property
    def wire_bytes(self) -> float:
        return float(self.L.ort_mg_wire_bytes(self.h))


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
    One header broadcast + one payload broadcast: [ids[n] padded to 4 words | nodes8[n*8]] as uint32 (int32 on the
    wire).  ort_upload_delta returns only after its scatter has read the payload, so the tensor may be recycled."""
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
    n_ids = 0 if is_full else (n + 3) // 4 * 4          # the ids section is padded to 16 bytes so that the rows stay 16-byte aligned
    words = n * 8 + n_ids
    if rank == src:
        parts = [] if is_full else [np.ascontiguousarray(ids, np.uint32), np.zeros(n_ids - n, np.uint32)]
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
        apply(payload[:n], payload[n_ids:].view(-1, 8), root, False)
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
