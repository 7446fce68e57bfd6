"""TEST INFRASTRUCTURE: ctypes wrapper of tests/host_emu/emu.cpp -- the device traversal header (csrc/ort_trace.cuh)
compiled for the host.  Lets the CPU-only suite check the kernels' per-ray code against the oracle; it is not a product
path (nothing under octree_ray_tracing_b200/ imports it)."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "octree_ray_tracing_b200", "csrc")
SANITIZE = os.environ.get("ORT_EMU_SANITIZE", "") == "1"      # tools/sanitize_host.sh: ASan + UBSan over the walkers' local arrays
SO = os.path.join(HERE, "_build", "libort_emu_san.so" if SANITIZE else "libort_emu.so")
_lib = None

STAT_FIELDS = ["rays", "slow_path_rays", "oob_loads", "lean_rays", "beam_rays", "beam_misses", "beam_guard", "beam_tile_misses", "beam_cert_wrong"]


def build(force: bool = False) -> str:
    deps = [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "cuda_shim.h"), os.path.join(CSRC, "ort_trace.cuh"), os.path.join(CSRC, "ort_trace_experiments.cuh"), os.path.join(CSRC, "ort_beam.cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-mfma", "-mavx2", "-ffp-contract=off", "-fPIC", "-shared",
           "-Wall", "-Wno-unused-function", deps[0], "-o", SO]
    if SANITIZE:
        cmd[1:2] = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode:
        raise RuntimeError("host emulation build failed:\n" + out.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_stats_words.restype = C.c_int
    return _lib


def default_rcp_table() -> np.ndarray:
    """The table the product ships (csrc/ort_rcp_table.h)."""
    txt = open(os.path.join(CSRC, "ort_rcp_table.h")).read()
    return np.array([int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)], np.uint32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _stats(raw):
    nlev = len(raw) - len(STAT_FIELDS)
    d = {"rounds_by_level": raw[:nlev].copy()}
    d.update({k: int(v) for k, v in zip(STAT_FIELDS, raw[nlev:])})
    return d


def trace_rays(nodes8, root, depth, o, d, walker=1, rcp_tab=None, miss_t=np.inf, want_npush=False, want_stats=False, nthreads=None, pool=False, allow_oob=False, tau=None):
    """pool=True: nodes8 is an och::octree pool (raw rows, root = row 0; the caller passes miss_t=0.0).
    tau: per-ray re-entry times of the beam start (walker 13 only; 0 = ordinary start)."""
    nodes8 = np.ascontiguousarray(nodes8, np.uint32)
    tab = default_rcp_table() if rcp_tab is None else np.ascontiguousarray(rcp_tab, np.uint32)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    n = len(d)
    o = np.ascontiguousarray(o, np.float32)
    o_stride = 0 if o.size == 3 else 3
    vox = np.empty(n, np.uint32); face = np.empty(n, np.uint8); t = np.empty(n, np.float32)
    npush = np.empty(n, np.uint16) if want_npush else None
    stats = np.zeros(lib().emu_stats_words(), np.uint64) if want_stats else None
    oob = lib().emu_trace_rays(_p(nodes8), C.c_size_t(nodes8.size // 8), 0 if pool else 1, 1 if (pool or root) else 0, C.c_uint32(root), depth, C.c_float(miss_t), _p(tab), int(np.log2(len(tab))),
                         _p(o), o_stride, _p(d), C.c_size_t(n), walker,
                         _p(vox), _p(face), _p(t), _p(npush), _p(stats), nthreads or os.cpu_count() or 1,
                         _p(None if tau is None else np.ascontiguousarray(tau, np.float32)))
    if oob and not allow_oob:
        raise MemoryError("host emulation: the walk loaded from outside the node array / reciprocal table")
    out = [vox, face, t]
    if want_npush:
        out.append(npush)
    if want_stats:
        out.append(_stats(stats))
    return tuple(out)


def beam_grid(nodes8, root, k, pool=False) -> np.ndarray:
    """The level-k skip grid of the beam start (csrc/ort_beam.cuh) for a DAG, computed on the host: (N, N, N) bytes [z, y, x]."""
    nodes8 = np.ascontiguousarray(nodes8, np.uint32)
    n = 1 << k
    skip = np.zeros((n, n, n), np.uint8)
    lib().emu_beam_grid(_p(nodes8), 0 if pool else 1, C.c_uint32(root), k, _p(skip))
    return skip


def beam_level(pos, rot, fov, W, H, depth, rcp_tab=None) -> int:
    """The grid level the library would pick for this camera (0: no beam start)."""
    tab = default_rcp_table() if rcp_tab is None else np.ascontiguousarray(rcp_tab, np.uint32)
    pos = np.ascontiguousarray(pos, np.float32)
    rot = np.ascontiguousarray(rot, np.float32)
    return int(lib().emu_beam_level(_p(pos), _p(rot), C.c_float(fov), W, H, depth, _p(tab), int(np.log2(len(tab)))))


def trace_frame(nodes8, root, depth, pos, rot, fov, W, H, y0=0, rows=None, tile_rows=1, tile_step=1, walker=1, rcp_tab=None, miss_t=np.inf,
                want_npush=False, want_stats=False, nthreads=None, pool=False, allow_oob=False, beam=None, want_tau=False):
    """beam: a skip grid from beam_grid() -- the frame is traced with the beam start of every 8 x 4 tile (walker 13)."""
    nodes8 = np.ascontiguousarray(nodes8, np.uint32)
    tab = default_rcp_table() if rcp_tab is None else np.ascontiguousarray(rcp_tab, np.uint32)
    rows = H - y0 if rows is None else rows
    n = rows * W
    pos = np.ascontiguousarray(pos, np.float32); rot = np.ascontiguousarray(rot, np.float32)
    vox = np.empty(n, np.uint32); face = np.empty(n, np.uint8); t = np.empty(n, np.float32)
    npush = np.empty(n, np.uint16) if want_npush else None
    stats = np.zeros(lib().emu_stats_words(), np.uint64) if want_stats else None
    tau = np.zeros(n, np.float32) if want_tau else None
    if beam is not None:
        beam = np.ascontiguousarray(beam, np.uint8)
    oob = lib().emu_trace_frame(_p(nodes8), C.c_size_t(nodes8.size // 8), 0 if pool else 1, 1 if (pool or root) else 0, C.c_uint32(root), depth, C.c_float(miss_t), _p(tab), int(np.log2(len(tab))),
                          _p(pos), _p(rot), C.c_float(fov), W, H, y0, rows, tile_rows, tile_step, walker,
                           _p(vox), _p(face), _p(t), _p(npush), _p(stats), nthreads or os.cpu_count() or 1,
                           _p(beam), 0 if beam is None else int(np.log2(beam.shape[0])), _p(tau))
    if oob and not allow_oob:
        raise MemoryError("host emulation: the walk loaded from outside the node array / reciprocal table")
    out = [vox, face, t]
    if want_npush:
        out.append(npush)
    if want_stats:
        out.append(_stats(stats))
    if want_tau:
        out.append(tau)
    return tuple(out)


def warp_model(nodes8, root, depth, pos, rot, fov, W, H, rcp_tab=None, nthreads=None):
    """SIMT cost model of the frame kernel (emu.cpp: emu_warp_model): the rays of each 8x4 tile advance in lockstep, one
    round per step.  Returns a dict of totals over the frame."""
    nodes8 = np.ascontiguousarray(nodes8, np.uint32)
    tab = default_rcp_table() if rcp_tab is None else np.ascontiguousarray(rcp_tab, np.uint32)
    pos = np.ascontiguousarray(pos, np.float32); rot = np.ascontiguousarray(rot, np.float32)
    out = np.zeros(10, np.uint64)
    lib().emu_warp_model(_p(nodes8), C.c_size_t(nodes8.size // 8), C.c_uint32(root), depth, _p(tab), int(np.log2(len(tab))),
                         _p(pos), _p(rot), C.c_float(fov), W, H, _p(out), nthreads or os.cpu_count() or 1)
    keys = ["warp_rounds", "descend_only", "advance_only", "both", "active_lanes", "descend_lanes", "advance_lanes", "warps", "_", "lane_rounds"]
    return {k: int(v) for k, v in zip(keys, out) if k != "_"}
