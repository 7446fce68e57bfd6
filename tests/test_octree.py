"""och::octree row: the pool octree's host side against the oracle on CPU, its trace on the GPU."""
import numpy as np
import pytest

from conftest import assert_same_hits


def random_ops(rs, n, dim, unset_frac=0.3):
    return np.concatenate([rs.randint(0, dim, (n, 3)), rs.randint(1, 9, (n, 1)), (rs.rand(n, 1) < unset_frac).astype(int)], 1).astype(np.int32)


def test_pool_matches_oracle_row_for_row(ort, oc):
    rs = np.random.RandomState(4)
    a, p = oc.OracleOctree(6, 1 << 14), ort.Octree(6, 1 << 14, device=None)
    for _ in range(3):
        ops = random_ops(rs, 5000, 64)
        a.apply(ops)
        p.apply(ops)
        assert a.node_cnt == p.get_node_cnt()
        assert np.array_equal(a.nodes(), p.nodes())
    pts = rs.randint(0, 64, (2000, 3))
    assert [a.at(*q) for q in pts] == [p.at(*q) for q in pts]
    # coordinates are taken modulo dim (no range check in the reference, och_octree.cpp:74-91)
    a.set(64 + 3, 5, 6, 9)
    p.set(64 + 3, 5, 6, 9)
    assert p.at(3, 5, 6) == 9 and np.array_equal(a.nodes(), p.nodes())


def test_pool_exhaustion_is_reported(ort):
    t = ort.Octree(4, 8, device=None)
    with pytest.raises(ort.OrtError):
        for i in range(16):
            t.set(i, i, i, 1)


@pytest.mark.gpu
def test_octree_trace_vs_oracle(ort, oc):
    from test_oracle import _builtin_table
    from golden.make_golden import edge_rays
    rs = np.random.RandomState(8)
    depth, cap = 7, 1 << 16
    A, T = oc.OracleOctree(depth, cap), ort.Octree(depth, cap)
    # a blobby scene: random boxes
    ops = []
    for _ in range(40):
        c = rs.randint(8, 120, 3)
        e = rs.randint(2, 10)
        v = int(rs.randint(1, 6))
        ops += [(c[0] + x, c[1] + y, c[2] + z, v, 0) for x in range(e) for y in range(e) for z in range(e)]
    ops = np.array(ops, np.int32)
    A.apply(ops)
    T.apply(ops)
    tab = _builtin_table()
    n = 200000
    o = rs.uniform(1.001, 1.999, (n, 3)).astype(np.float32)
    d = rs.normal(size=(n, 3)).astype(np.float32)
    got = T.trace_rays(o, d)
    assert_same_hits(got, A.trace(o, d, rcp_tab=tab, nthreads=4), "random rays")
    miss = got[0] == 0
    assert miss.any() and (got[2][miss & (got[1] == 6)] == 0.0).all()       # MISS reports t = 0.0F (och_octree.cpp:302)
    o, d = edge_rays(rs, 300)
    assert_same_hits(T.trace_rays(o, d), A.trace(o, d, rcp_tab=tab), "edge rays")
    # in-place edits -> row deltas
    for step in range(4):
        ops = random_ops(rs, 300, 128, 0.5)
        A.apply(ops)
        T.apply(ops)
        n_up, full = T.sync()
        assert not full and 0 < n_up < 3000
        got = T.trace_frame((1.5, 1.5, 1.95), 0.3, -1.1, 320, 200)
        rot, fov = oc.camera_coeffs(0.3, -1.1)
        want = A.trace(np.array([1.5, 1.5, 1.95], np.float32), oc.gen_rays(rot, fov, 320, 200), rcp_tab=tab, nthreads=4)
        assert_same_hits(got, want, f"frame after edits {step}")
    dr, vox, t = T.sse_trace((1.5, 1.5, 1.95), (0.0, 0.1, 1.0))
    assert dr == ort.Direction.exit and vox == 0 and t == 0.0


def test_octree_live_vs_reference(oc):
    """The restated pool + trace against the real och::octree (oracle/_ref), where it is built."""
    if not oc.have_ref():
        pytest.skip("oracle/_ref/libochref.so not built")
    rs = np.random.RandomState(2)
    A, B = oc.OracleOctree(6, 1 << 14), oc.RefOctree(6, 1 << 14)
    ops = random_ops(rs, 6000, 64)
    A.apply(ops)
    B.apply(ops)
    assert A.node_cnt == B.node_cnt and np.array_equal(A.nodes(), B.nodes())
    o = rs.uniform(1.001, 1.999, (30000, 3)).astype(np.float32)
    d = rs.normal(size=(30000, 3)).astype(np.float32)
    assert_same_hits(A.trace(o, d), B.trace(o, d), "octree restatement vs reference")
