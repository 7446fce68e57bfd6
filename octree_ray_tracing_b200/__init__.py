"""octree_ray_tracing_b200 -- B200-native (sm_100a) implementation of the hot path of
AlexanderRipar/Octree_Ray_Tracing: och::h_octree<L,D>::sse_trace and the node store behind it.

The product is ``libort_b200.so`` (hand-written CUDA + C++ host side, C ABI in
``include/ort_b200.h``).  This package is the thin Python face of that ABI, shaped like the
reference's own C++ interface (``HOctree`` mirrors ``och::h_octree``), used by the tests, the
headless harness and ``bench.py``.
"""
from ._lib import OrtError, lib, LIB_PATH  # noqa: F401
from .tree import HOctree, Octree, TraceContext, Direction, camera_coeffs, host_rcp_table  # noqa: F401
from . import harness  # noqa: F401
