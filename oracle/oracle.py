"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two libraries live here:

* ``liboch_oracle.so``  -- the plain-C restatement (``och_oracle.c``) of the reference's
  ``och::h_octree`` path (och_h_octree.h:17-452, test_och_h_octree.cpp:87-138, :561-787).
* ``_ref/libochref.so`` -- the UNMODIFIED reference compiled in place from /root/reference by
  ``oracle/Makefile`` (``ref_build/ref_wrap.cpp``).  Present only if it was built in the
  authoring container (it travels to the GPU box as a prebuilt file).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboch_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libochref.so")

_vp = C.c_void_p


def build(verbose: bool = False) -> None:
    """(Re)build liboch_oracle.so and, when /root/reference exists, _ref/libochref.so."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout + out.stderr)
    if out.returncode:
        raise RuntimeError("oracle build failed")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Counters(C.Structure):
    _fields_ = [("push", C.c_uint64), ("step", C.c_uint64), ("pop", C.c_uint64)]


class _OcTree(C.Structure):
    _fields_ = [
        ("log2cap", C.c_int), ("depth", C.c_int), ("cap", C.c_uint32), ("idx_mask", C.c_uint32),
        ("cashes", C.POINTER(C.c_uint8)), ("refcounts", C.POINTER(C.c_uint32)), ("nodes", C.POINTER(C.c_uint32)),
        ("root_idx", C.c_uint32), ("fillcnt", C.c_uint32), ("nodecnt", C.c_uint32), ("table_full", C.c_int),
    ]


class _OcOctree(C.Structure):
    _fields_ = [("depth", C.c_int), ("cap", C.c_uint32), ("nodes", C.POINTER(C.c_uint32)), ("head", C.c_uint32),
                ("node_cnt", C.c_int), ("failed", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.oc_rcp_hw_bits.restype = C.c_uint32
        L.oc_rcp_hw_bits.argtypes = [C.c_uint32]
        L.oc_rcp_table_from_hw.restype = C.c_long
        L.oc_rcp_table_from_hw.argtypes = [_vp, C.c_int]
        L.oc_rcp_table_bits.restype = C.c_uint32
        L.oc_rcp_table_bits.argtypes = [_vp, C.c_int, C.c_uint32]
        L.oc_tree_create.restype = C.POINTER(_OcTree)
        L.oc_tree_create.argtypes = [C.c_int, C.c_int]
        L.oc_tree_destroy.argtypes = [C.POINTER(_OcTree)]
        L.oc_node_hash.restype = C.c_uint32
        L.oc_node_hash.argtypes = [_vp]
        L.oc_register_node.restype = C.c_uint32
        L.oc_register_node.argtypes = [C.POINTER(_OcTree), _vp]
        L.oc_remove_node.argtypes = [C.POINTER(_OcTree), C.c_uint32]
        L.oc_set.argtypes = [C.POINTER(_OcTree), C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint32]
        L.oc_set_many.argtypes = [C.POINTER(_OcTree), _vp, C.c_size_t]
        L.oc_at.restype = C.c_uint32
        L.oc_at.argtypes = [C.POINTER(_OcTree), C.c_int, C.c_int, C.c_int]
        L.oc_clear.argtypes = [C.POINTER(_OcTree)]
        L.oc_z_encode_16.restype = C.c_uint64
        L.oc_z_encode_16.argtypes = [C.c_uint16, C.c_uint16, C.c_uint16]
        L.oc_trace_rays.argtypes = [_vp, C.c_uint32, C.c_int, _vp, C.c_int, _vp, C.c_size_t, _vp, C.c_int,
                                    _vp, _vp, _vp, _vp, C.POINTER(Counters), C.c_int]
        L.oc_camera_coeffs.argtypes = [C.c_float, C.c_float, _vp, _vp]
        L.oc_gen_rays.argtypes = [_vp, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
        L.oc_simplex2_many.argtypes = [C.c_float, _vp, C.c_size_t, _vp]
        L.oc_simplex3_many.argtypes = [C.c_float, _vp, C.c_size_t, _vp]
        L.oc_heightmap.argtypes = [C.c_int, _vp]
        L.oc_initialize_terrain.argtypes = [C.POINTER(_OcTree), _vp, _vp, C.c_int]
        L.oc_octree_create.restype = C.POINTER(_OcOctree)
        L.oc_octree_create.argtypes = [C.c_int, C.c_uint32]
        L.oc_octree_destroy.argtypes = [C.POINTER(_OcOctree)]
        L.oc_octree_set.argtypes = [C.POINTER(_OcOctree), C.c_int16, C.c_int16, C.c_int16, C.c_uint32]
        L.oc_octree_unset.argtypes = [C.POINTER(_OcOctree), C.c_int16, C.c_int16, C.c_int16]
        L.oc_octree_at.restype = C.c_uint32
        L.oc_octree_at.argtypes = [C.POINTER(_OcOctree), C.c_int16, C.c_int16, C.c_int16]
        L.oc_octree_apply.argtypes = [C.POINTER(_OcOctree), _vp, C.c_size_t]
        L.oc_octree_trace_rays.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, C.c_size_t, _vp, C.c_int,
                                           _vp, _vp, _vp, _vp, C.POINTER(Counters), C.c_int]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------------
# reciprocal table
# ------------------------------------------------------------------------------------------------

def rcp_table_from_hw(log2n: int = 11):
    """(table, mismatches): table model of this host's RCPSS and how often it disagrees with it."""
    tab = np.zeros(1 << log2n, np.uint32)
    bad = lib().oc_rcp_table_from_hw(_ptr(tab), log2n)
    return tab, int(bad)


# ------------------------------------------------------------------------------------------------
# tree
# ------------------------------------------------------------------------------------------------

class OracleTree:
    """Restated h_octree<log2cap, depth> node store (och_h_octree.h:17-288)."""

    def __init__(self, log2cap: int, depth: int):
        self.L = lib()
        self.h = self.L.oc_tree_create(log2cap, depth)
        self.log2cap, self.depth, self.cap = log2cap, depth, 1 << log2cap

    def __del__(self):
        try:
            self.L.oc_tree_destroy(self.h)
        except Exception:
            pass

    def register_node(self, c8) -> int:
        a = np.ascontiguousarray(c8, np.uint32)
        return self.L.oc_register_node(self.h, _ptr(a))

    def remove_node(self, idx: int) -> None:
        self.L.oc_remove_node(self.h, idx)

    def set(self, x, y, z, v) -> None:
        self.L.oc_set(self.h, x & 0xFFFF, y & 0xFFFF, z & 0xFFFF, v)

    def set_many(self, xyzv) -> None:
        a = np.ascontiguousarray(xyzv, np.uint32).reshape(-1, 4)
        self.L.oc_set_many(self.h, _ptr(a), a.shape[0])

    def at(self, x, y, z) -> int:
        return self.L.oc_at(self.h, x, y, z)

    def clear(self) -> None:
        self.L.oc_clear(self.h)

    @property
    def root(self) -> int:
        return self.h.contents.root_idx

    @root.setter
    def root(self, r: int) -> None:
        self.h.contents.root_idx = r

    @property
    def fillcnt(self) -> int:
        return self.h.contents.fillcnt

    @property
    def nodecnt(self) -> int:
        return self.h.contents.nodecnt

    @property
    def table_full(self) -> bool:
        return bool(self.h.contents.table_full)

    def nodes(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.h.contents.nodes, shape=(self.cap, 8))

    def cashes(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.h.contents.cashes, shape=(self.cap,))

    def refcounts(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.h.contents.refcounts, shape=(self.cap,))

    def initialize_terrain(self, heights, grass, tunnels: bool) -> None:
        hm = np.ascontiguousarray(heights, np.uint16)
        g = np.ascontiguousarray(grass, np.uint8)
        self.L.oc_initialize_terrain(self.h, _ptr(hm), _ptr(g), int(tunnels))

    def trace(self, o, d, **kw):
        return trace_rays(self.nodes(), self.root, self.depth, o, d, **kw)


def _trace_args(o, d):
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    o = np.ascontiguousarray(o, np.float32)
    n = d.shape[0]
    stride = 0 if o.size == 3 else 3
    assert stride == 0 or o.size == 3 * n
    return o, d, n, stride, np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.float32)


class OracleOctree:
    """Restated och::octree (och_octree.h:10-69, och_octree.cpp:14-320)."""

    def __init__(self, depth: int, table_capacity: int):
        self.L = lib()
        self.h = self.L.oc_octree_create(depth, table_capacity)
        self.depth, self.cap = depth, table_capacity

    def __del__(self):
        try:
            self.L.oc_octree_destroy(self.h)
        except Exception:
            pass

    def set(self, x, y, z, v):
        self.L.oc_octree_set(self.h, x, y, z, v)

    def unset(self, x, y, z):
        self.L.oc_octree_unset(self.h, x, y, z)

    def at(self, x, y, z):
        return self.L.oc_octree_at(self.h, x, y, z)

    def apply(self, ops):
        a = np.ascontiguousarray(ops, np.int32).reshape(-1, 5)
        self.L.oc_octree_apply(self.h, _ptr(a), a.shape[0])

    @property
    def node_cnt(self):
        return self.h.contents.node_cnt

    @property
    def failed(self):
        return bool(self.h.contents.failed)

    def nodes(self):
        return np.ctypeslib.as_array(self.h.contents.nodes, shape=(self.cap, 8))

    def trace(self, o, d, rcp_tab=None, nthreads=1):
        return octree_trace_rays(self.nodes(), self.depth, o, d, rcp_tab, nthreads)


def octree_trace_rays(pool, depth, o, d, rcp_tab=None, nthreads=1):
    pool = np.ascontiguousarray(pool, np.uint32)
    o, d, n, stride, vox, face, t = _trace_args(o, d)
    log2n = 0
    if rcp_tab is not None:
        rcp_tab = np.ascontiguousarray(rcp_tab, np.uint32)
        log2n = int(rcp_tab.size).bit_length() - 1
    tot = Counters()
    lib().oc_octree_trace_rays(_ptr(pool), depth, _ptr(o), stride, _ptr(d), n, _ptr(rcp_tab), log2n,
                               _ptr(vox), _ptr(face), _ptr(t), None, C.byref(tot), nthreads)
    return vox, face, t


def heightmap(depth: int) -> np.ndarray:
    dim = 1 << depth
    h = np.zeros((dim, dim), np.uint16)
    lib().oc_heightmap(depth, _ptr(h))
    return h


def grass_bits(depth: int, seed: int = 1) -> np.ndarray:
    """Deterministic stand-in for the reference's unseeded std::rand() > RAND_MAX/2
    (test_och_h_octree.cpp:780): one bit per column, row-major, from numpy's MT19937(seed)."""
    dim = 1 << depth
    rs = np.random.RandomState(seed)
    return (rs.randint(0, 2, size=(dim, dim))).astype(np.uint8)


def simplex2(freq, xy):
    a = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.zeros(a.shape[0], np.float32)
    lib().oc_simplex2_many(freq, _ptr(a), a.shape[0], _ptr(out))
    return out


def simplex3(freq, xyz):
    a = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.zeros(a.shape[0], np.float32)
    lib().oc_simplex3_many(freq, _ptr(a), a.shape[0], _ptr(out))
    return out


# ------------------------------------------------------------------------------------------------
# rays + trace
# ------------------------------------------------------------------------------------------------

def camera_coeffs(yaw: float, pitch: float):
    rot = np.zeros(9, np.float32)
    fov = C.c_float(0)
    lib().oc_camera_coeffs(yaw, pitch, _ptr(rot), C.byref(fov))
    return rot, float(np.float32(fov.value))


def gen_rays(rot, fov_factor, W, H, y0=0, y1=None) -> np.ndarray:
    y1 = H if y1 is None else y1
    rot = np.ascontiguousarray(rot, np.float32)
    d = np.zeros(((y1 - y0) * W, 3), np.float32)
    lib().oc_gen_rays(_ptr(rot), fov_factor, W, H, y0, y1, _ptr(d))
    return d


def trace_rays(nodes, root, depth, o, d, rcp_tab=None, nthreads=1, want_counts=False):
    """Trace n rays.  o: (3,) shared origin or (n,3); d: (n,3).
    Returns (voxel u32[n], face u8[n], t f32[n]) and, with want_counts, also
    (npush u16[n], Counters)."""
    nodes = np.ascontiguousarray(nodes, np.uint32)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    o = np.ascontiguousarray(o, np.float32)
    n = d.shape[0]
    stride = 0 if o.size == 3 else 3
    assert stride == 0 or o.size == 3 * n
    vox = np.zeros(n, np.uint32)
    face = np.zeros(n, np.uint8)
    t = np.zeros(n, np.float32)
    np16 = np.zeros(n, np.uint16) if want_counts else None
    tot = Counters()
    log2n = 0
    if rcp_tab is not None:
        rcp_tab = np.ascontiguousarray(rcp_tab, np.uint32)
        log2n = int(rcp_tab.size).bit_length() - 1
    lib().oc_trace_rays(_ptr(nodes), root, depth, _ptr(o), stride, _ptr(d), n, _ptr(rcp_tab), log2n,
                        _ptr(vox), _ptr(face), _ptr(t), _ptr(np16), C.byref(tot), nthreads)
    if want_counts:
        return vox, face, t, np16, tot
    return vox, face, t


# ------------------------------------------------------------------------------------------------
# the real reference (oracle/_ref)
# ------------------------------------------------------------------------------------------------

_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_SO)
        R.ochref_tree_create.restype = _vp
        R.ochref_tree_create.argtypes = [C.c_int, C.c_int]
        R.ochref_register_node.restype = C.c_uint32
        R.ochref_register_node.argtypes = [_vp, _vp]
        R.ochref_remove_node.argtypes = [_vp, C.c_uint32]
        R.ochref_set.argtypes = [_vp, C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint32]
        R.ochref_set_many.argtypes = [_vp, _vp, C.c_size_t]
        R.ochref_at.restype = C.c_uint32
        R.ochref_at.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
        R.ochref_set_root.argtypes = [_vp, C.c_uint32]
        R.ochref_get_root.restype = C.c_uint32
        R.ochref_get_root.argtypes = [_vp]
        R.ochref_get_fillcnt.restype = C.c_uint32
        R.ochref_get_fillcnt.argtypes = [_vp]
        R.ochref_get_nodecnt.restype = C.c_uint32
        R.ochref_get_nodecnt.argtypes = [_vp]
        R.ochref_clear.argtypes = [_vp]
        for f in ("ochref_nodes", "ochref_cashes", "ochref_refcounts"):
            getattr(R, f).restype = _vp
            getattr(R, f).argtypes = [_vp]
        R.ochref_import_compact.argtypes = [_vp, _vp, C.c_size_t, C.c_uint32]
        R.ochref_trace_batch.restype = C.c_int
        R.ochref_trace_batch.argtypes = [_vp, _vp, C.c_int, _vp, C.c_size_t, _vp, _vp, _vp, C.c_int]
        R.ochref_z_encode_16.restype = C.c_uint64
        R.ochref_z_encode_16.argtypes = [C.c_uint16, C.c_uint16, C.c_uint16]
        R.ochref_simplex2.argtypes = [C.c_float, _vp, C.c_size_t, _vp]
        R.ochref_simplex3.argtypes = [C.c_float, _vp, C.c_size_t, _vp]
        R.ochref_opensimplex2.argtypes = [C.c_int64, _vp, C.c_size_t, _vp]
        R.ochref_heightmap.argtypes = [C.c_int, _vp]
        R.ochref_initialize_terrain.argtypes = [_vp, _vp, _vp, C.c_int]
        R.ochref_octree_create.restype = _vp
        R.ochref_octree_create.argtypes = [C.c_int, C.c_uint32]
        R.ochref_octree_set.argtypes = [_vp, C.c_int16, C.c_int16, C.c_int16, C.c_uint32]
        R.ochref_octree_unset.argtypes = [_vp, C.c_int16, C.c_int16, C.c_int16]
        R.ochref_octree_at.restype = C.c_uint32
        R.ochref_octree_at.argtypes = [_vp, C.c_int16, C.c_int16, C.c_int16]
        R.ochref_octree_node_cnt.restype = C.c_int
        R.ochref_octree_node_cnt.argtypes = [_vp]
        R.ochref_octree_nodes.restype = _vp
        R.ochref_octree_nodes.argtypes = [_vp]
        R.ochref_octree_apply.argtypes = [_vp, _vp, C.c_size_t]
        R.ochref_octree_trace_batch.argtypes = [_vp, _vp, C.c_int, _vp, C.c_size_t, _vp, _vp, _vp]
        R.ochref_node_hash.restype = C.c_uint32
        R.ochref_node_hash.argtypes = [_vp]
        _ref = R
    return _ref


REF_CONFIGS = [(12, 4), (16, 6), (19, 8), (22, 10), (24, 12), (25, 13), (25, 14)]   # the last two: trace-only (import_compact) for the depth-13 / 14 parity tests


class RefTree:
    """The reference's own och::h_octree<log2cap, depth> (instantiated for REF_CONFIGS)."""

    def __init__(self, log2cap: int, depth: int):
        self.R = ref()
        self.h = _vp(self.R.ochref_tree_create(log2cap, depth))
        if not self.h:
            raise ValueError(f"h_octree<{log2cap},{depth}> is not instantiated in libochref.so")
        self.log2cap, self.depth, self.cap = log2cap, depth, 1 << log2cap

    def register_node(self, c8) -> int:
        a = np.ascontiguousarray(c8, np.uint32)
        return self.R.ochref_register_node(self.h, _ptr(a))

    def remove_node(self, idx):
        self.R.ochref_remove_node(self.h, idx)

    def set(self, x, y, z, v):
        self.R.ochref_set(self.h, x & 0xFFFF, y & 0xFFFF, z & 0xFFFF, v)

    def set_many(self, xyzv):
        a = np.ascontiguousarray(xyzv, np.uint32).reshape(-1, 4)
        self.R.ochref_set_many(self.h, _ptr(a), a.shape[0])

    def at(self, x, y, z):
        return self.R.ochref_at(self.h, x, y, z)

    def clear(self):
        self.R.ochref_clear(self.h)

    @property
    def root(self):
        return self.R.ochref_get_root(self.h)

    @root.setter
    def root(self, r):
        self.R.ochref_set_root(self.h, r)

    @property
    def fillcnt(self):
        return self.R.ochref_get_fillcnt(self.h)

    @property
    def nodecnt(self):
        return self.R.ochref_get_nodecnt(self.h)

    def _arr(self, fn, dtype, shape):
        p = getattr(self.R, fn)(self.h)
        ct = {np.uint32: C.c_uint32, np.uint8: C.c_uint8}[dtype]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=shape)

    def nodes(self):
        return self._arr("ochref_nodes", np.uint32, (self.cap, 8))

    def cashes(self):
        return self._arr("ochref_cashes", np.uint8, (self.cap,))

    def refcounts(self):
        return self._arr("ochref_refcounts", np.uint32, (self.cap,))

    def import_compact(self, nodes8, root):
        a = np.ascontiguousarray(nodes8, np.uint32).reshape(-1, 8)
        assert a.shape[0] <= self.cap
        self.R.ochref_import_compact(self.h, _ptr(a), a.shape[0], root)

    def initialize_terrain(self, heights, grass, tunnels: bool):
        hm = np.ascontiguousarray(heights, np.uint16)
        g = np.ascontiguousarray(grass, np.uint8)
        self.R.ochref_initialize_terrain(self.h, _ptr(hm), _ptr(g), int(tunnels))

    def trace(self, o, d, nthreads=1):
        """The reference's own sse_trace (och_h_octree.h:292) over n rays."""
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        o = np.ascontiguousarray(o, np.float32)
        n = d.shape[0]
        stride = 0 if o.size == 3 else 3
        vox = np.zeros(n, np.uint32)
        face = np.zeros(n, np.uint8)
        t = np.zeros(n, np.float32)
        rc = self.R.ochref_trace_batch(self.h, _ptr(o), stride, _ptr(d), n, _ptr(vox), _ptr(face), _ptr(t), nthreads)
        assert rc == 0
        return vox, face, t


class RefOctree:
    """The reference's own och::octree (och_octree.cpp compiled with the one hoisted declaration, see Makefile)."""

    def __init__(self, depth: int, table_capacity: int):
        self.R = ref()
        self.h = _vp(self.R.ochref_octree_create(depth, table_capacity))
        self.depth, self.cap = depth, table_capacity

    def set(self, x, y, z, v):
        self.R.ochref_octree_set(self.h, x, y, z, v)

    def unset(self, x, y, z):
        self.R.ochref_octree_unset(self.h, x, y, z)

    def at(self, x, y, z):
        return self.R.ochref_octree_at(self.h, x, y, z)

    def apply(self, ops):
        a = np.ascontiguousarray(ops, np.int32).reshape(-1, 5)
        self.R.ochref_octree_apply(self.h, _ptr(a), a.shape[0])

    @property
    def node_cnt(self):
        return self.R.ochref_octree_node_cnt(self.h)

    def nodes(self):
        p = self.R.ochref_octree_nodes(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.cap, 8))

    def trace(self, o, d):
        o, d, n, stride, vox, face, t = _trace_args(o, d)
        self.R.ochref_octree_trace_batch(self.h, _ptr(o), stride, _ptr(d), n, _ptr(vox), _ptr(face), _ptr(t))
        return vox, face, t


def ref_heightmap(depth: int) -> np.ndarray:
    dim = 1 << depth
    h = np.zeros((dim, dim), np.uint16)
    ref().ochref_heightmap(depth, _ptr(h))
    return h
