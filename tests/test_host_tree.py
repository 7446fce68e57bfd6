"""CPU tests of the product's host side (C++ node store, flatten/delta, fixture builder, camera)
against the oracle.  No GPU needed: nothing here calls a trace entry point."""
import os

import numpy as np
import pytest

from conftest import ROOT, assert_same_hits, same_bits


def live_mask(tags):
    return (tags != 0) & (tags != 0xFF)


def test_table_matches_oracle_slot_for_slot(ort, oc):
    rs = np.random.RandomState(3)
    a, p = oc.OracleTree(16, 6), ort.HOctree(16, 6, device=None)
    for rnd in range(4):
        ops = np.concatenate([rs.randint(0, 70, (15000, 3)), rs.randint(0, 4, (15000, 1))], 1).astype(np.uint32)
        a.set_many(ops)
        p.set_many(ops)
        assert (a.root, a.fillcnt, a.nodecnt) == (p.get_root(), p.get_fillcnt(), p.get_nodecnt())
        assert np.array_equal(a.cashes(), p.cashes())
        assert np.array_equal(a.refcounts(), p.refcounts())
        lm = live_mask(a.cashes())
        assert np.array_equal(a.nodes()[lm], p.nodes()[lm])
    pts = rs.randint(-2, 66, (3000, 3))
    pts = pts[(pts >= 0).all(1) & (pts < 64).all(1)]
    assert [a.at(*q) for q in pts] == [p.at(*q) for q in pts]


def test_register_remove_api(ort, oc):
    a, p = oc.OracleTree(12, 4), ort.HOctree(12, 4, device=None)
    rs = np.random.RandomState(5)
    ids = []
    for _ in range(300):
        n = rs.randint(0, 3, 8).astype(np.uint32)
        if not n.any():
            n[0] = 1
        i, j = a.register_node(n), p.register_node(n)
        assert i == j
        ids.append(i)
    for i in ids[::2]:
        a.remove_node(i)
        p.remove_node(i)
    assert np.array_equal(a.cashes(), p.cashes()) and np.array_equal(a.refcounts(), p.refcounts())
    assert (a.fillcnt, a.nodecnt) == (p.get_fillcnt(), p.get_nodecnt())
    # re-registering reuses gravestones the same way
    for _ in range(200):
        n = rs.randint(0, 3, 8).astype(np.uint32)
        n[7] = 5
        assert a.register_node(n) == p.register_node(n)
    assert np.array_equal(a.cashes(), p.cashes())
    assert p.get_max_refcnt() == 0      # never maintained by the reference (och_h_octree.h:99)


def test_set_edge_cases(ort):
    t = ort.HOctree(12, 4, device=None)
    assert t.get_root() == 0 and t.at(1, 2, 3) == 0
    t.set(1, 2, 3, 0)                     # removing from an empty tree: no-op
    assert t.get_root() == 0 and t.get_fillcnt() == 0
    t.set(16, 0, 0, 1)                    # out of range (dim = 16): ignored (och_h_octree.h:178)
    t.set(0, 65535, 0, 1)
    assert t.get_root() == 0
    t.set(1, 2, 3, 9)
    assert t.at(1, 2, 3) == 9 and t.get_fillcnt() == 4
    t.set(1, 2, 3, 0)                     # last voxel removed: collapses to the empty tree
    assert t.get_root() == 0 and t.get_fillcnt() == 0 and t.get_nodecnt() == 0
    t.set(15, 15, 15, 2)
    t.clear()                             # clear() only zeroes the tags (och_h_octree.h:285-288)
    assert t.get_root() != 0 and not t.cashes().any()


def test_set_box_is_the_reference_loop(ort, oc):
    a, p = oc.OracleTree(16, 6), ort.HOctree(16, 6, device=None)
    for (cx, cy, cz, ext, v) in ((30, 30, 30, 10, 1), (2, 60, 33, 9, 3), (30, 30, 28, 6, 0)):
        ops = [((cx + x) & 0xFFFF, (cy + y) & 0xFFFF, (cz + z) & 0xFFFF, v)
               for z in range(-(ext // 2), (ext + 1) // 2) for y in range(-(ext // 2), (ext + 1) // 2) for x in range(-(ext // 2), (ext + 1) // 2)]
        a.set_many(np.array(ops, np.uint32))
        p.set_box(cx, cy, cz, ext, v)
        assert np.array_equal(a.cashes(), p.cashes()) and a.root == p.get_root()


def test_camera_coeffs_and_fixture_heightmap(ort, oc):
    for yaw, pitch in ((0.0, 0.0), (0.7, -0.6), (0.785, -0.3), (-2.5, 1.2)):
        r1, f1 = ort.camera_coeffs(yaw, pitch)
        r2, f2 = oc.camera_coeffs(yaw, pitch)
        assert same_bits(r1, r2) and f1 == f2
    for d in (4, 6, 8):
        assert np.array_equal(ort.harness.heightmap(d, 3), oc.heightmap(d))


@pytest.mark.parametrize("depth,log2cap,tunnels", [(6, 16, False), (6, 16, True), (8, 19, False), (8, 19, True)])
def test_fixture_builder_gives_the_canonical_dag(ort, oc, depth, log2cap, tunnels):
    """The memoising builder and the straight restatement of initialize_h_octree end in the same
    content-addressed DAG: same node count, instance counts, voxels, and traced image."""
    h, g = oc.heightmap(depth), oc.grass_bits(depth)
    A = oc.OracleTree(log2cap, depth)
    A.initialize_terrain(h, g, tunnels)
    T = ort.HOctree(log2cap, depth, device=None)
    ort.harness.build_terrain(T, h, g, tunnels=tunnels)
    assert (A.fillcnt, A.nodecnt) == (T.get_fillcnt(), T.get_nodecnt())
    assert np.array_equal(np.sort(A.refcounts()[live_mask(A.cashes())]), np.sort(T.refcounts()[live_mask(T.cashes())]))
    rs = np.random.RandomState(1)
    for q in rs.randint(0, 1 << depth, (4000, 3)):
        assert A.at(*q) == T.at(*q)
    nodes8, root, lo = T.flatten()
    assert nodes8.shape[0] == A.fillcnt and root == 1 and lo[0] == 1 and lo[-1] == nodes8.shape[0] + 1
    rot, fov = oc.camera_coeffs(0.7, -0.6)
    d = oc.gen_rays(rot, fov, 160, 90)
    o = np.array([1.5, 1.5, 1.5], np.float32)
    assert_same_hits(oc.trace_rays(nodes8, root, depth, o, d), A.trace(o, d), "flattened vs hashed table")
    # edits after a memoised build keep the table consistent (refcounts were real instance counts)
    ops = np.concatenate([rs.randint(0, 1 << depth, (3000, 3)), rs.randint(0, 3, (3000, 1))], 1).astype(np.uint32)
    A.set_many(ops)
    T.set_many(ops)
    assert (A.fillcnt, A.nodecnt) == (T.get_fillcnt(), T.get_nodecnt())
    n2, r2, _ = T.flatten()
    assert n2.shape[0] == T.get_fillcnt()
    assert_same_hits(oc.trace_rays(n2, r2, depth, o, d), A.trace(o, d), "after edits")


def apply_delta(mirror, ids, nodes8):
    need = int(ids.max()) if ids.size else 0
    if need > mirror.shape[0]:
        mirror = np.concatenate([mirror, np.zeros((need - mirror.shape[0], 8), np.uint32)])
    mirror[ids - 1] = nodes8
    return mirror


def test_delta_stream_keeps_a_mirror_traceable(ort, oc):
    """Simulated device mirror on the CPU: full flatten once, then only deltas after each burst of
    edits (single voxels and 40^3-style boxes).  The oracle traced over the mirror must always
    equal the oracle traced over a straight oracle table that saw the same edits."""
    depth, log2cap = 6, 16
    h, g = oc.heightmap(depth), oc.grass_bits(depth)
    A = oc.OracleTree(log2cap, depth)
    A.initialize_terrain(h, g, False)
    T = ort.HOctree(log2cap, depth, device=None)
    ort.harness.build_terrain(T, h, g)
    ids, nodes8, root, full = T.take_delta()
    assert full and ids is None
    mirror = nodes8.copy()
    rs = np.random.RandomState(9)
    rot, fov = oc.camera_coeffs(0.3, -0.9)
    d = oc.gen_rays(rot, fov, 96, 54)
    o = np.array([1.5, 1.5, 1.8], np.float32)
    total_delta = 0
    for step in range(30):
        if step % 3 == 2:
            cx, cy, cz = (int(v) for v in rs.randint(8, 56, 3))
            ext, v = int(rs.randint(3, 12)), int(rs.randint(0, 2))
            ops = np.array([((cx + x) & 0xFFFF, (cy + y) & 0xFFFF, (cz + z) & 0xFFFF, v)
                            for z in range(-(ext // 2), (ext + 1) // 2) for y in range(-(ext // 2), (ext + 1) // 2)
                            for x in range(-(ext // 2), (ext + 1) // 2)], np.uint32)
            T.set_box(cx, cy, cz, ext, v)
        else:
            ops = np.concatenate([rs.randint(0, 64, (40, 3)), rs.randint(0, 5, (40, 1))], 1).astype(np.uint32)
            T.set_many(ops)
        A.set_many(ops)
        ids, nodes8, root, full = T.take_delta()
        if full:
            mirror = nodes8.copy()
        else:
            mirror = apply_delta(mirror, ids, nodes8)
            total_delta += ids.size
        assert_same_hits(oc.trace_rays(mirror, root, depth, o, d), A.trace(o, d), f"step {step}")
        assert mirror.shape[0] < 4 * max(T.get_fillcnt(), 256)       # freed ids are recycled
    assert total_delta > 0
    # nothing pending -> empty delta, same root
    ids, nodes8, root2, full = T.take_delta()
    assert not full and ids.size == 0 and root2 == root
    # emptying the tree gives root 0
    for z in range(64):
        T.set_box(32, 32, z, 64, 0)
    ids, nodes8, root3, full = T.take_delta()
    assert T.get_root() == 0 and root3 == 0


def test_flatten_is_level_ordered(ort, oc):
    T = ort.HOctree(19, 8, device=None)
    ort.harness.build_terrain(T, oc.heightmap(8), oc.grass_bits(8))
    nodes8, root, lo = T.flatten()
    assert root == 1 and list(lo) == sorted(lo)
    for level in range(1, 8):                        # interior children point into the next level's id range
        rows = nodes8[lo[level - 1] - 1: lo[level] - 1]
        ch = rows[rows != 0]
        assert ch.min() >= lo[level] and ch.max() < lo[level + 1]
    leaf = nodes8[lo[7] - 1:]
    assert leaf.max() <= 4                           # voxel payloads, untranslated


def test_fill_box_equals_the_set_loop(ort, oc):
    """Bulk box edit: same voxels, live-node count, instance counts and traced image as the
    reference's 64 000-set() loop -- on terrain, on an empty tree, clipped at the cube border, place and remove."""
    depth, log2cap = 7, 18
    h, g = oc.heightmap(depth), oc.grass_bits(depth)
    A = oc.OracleTree(log2cap, depth)
    A.initialize_terrain(h, g, True)
    T = ort.HOctree(log2cap, depth, device=None)
    ort.harness.build_terrain(T, h, g, tunnels=True)
    E, F = oc.OracleTree(log2cap, depth), ort.HOctree(log2cap, depth, device=None)     # start empty
    rs = np.random.RandomState(21)
    rot, fov = oc.camera_coeffs(0.5, -0.7)
    d = oc.gen_rays(rot, fov, 128, 72)
    o = np.array([1.5, 1.5, 1.6], np.float32)
    boxes = [((-5, 60, 20), (9, 75, 40), 1), ((100, 100, 30), (140, 140, 70), 2), ((0, 0, 0), (128, 128, 3), 0), ((64, 64, 0), (65, 65, 128), 7)]
    boxes += [(tuple(c - e // 2), tuple(c + (e + 1) // 2), int(v)) for c, e, v in
              ((rs.randint(0, 128, 3), int(rs.randint(1, 41)), rs.randint(0, 4)) for _ in range(14))]
    for lo, hi, v in boxes:
        ops = np.array([(x, y, z, v) for z in range(max(lo[2], 0), min(hi[2], 128)) for y in range(max(lo[1], 0), min(hi[1], 128))
                        for x in range(max(lo[0], 0), min(hi[0], 128))], np.uint32).reshape(-1, 4)
        for a, t in ((A, T), (E, F)):
            a.set_many(ops)
            t.fill_box(lo, hi, v)
            assert (a.fillcnt, a.nodecnt) == (t.get_fillcnt(), t.get_nodecnt()), (lo, hi, v)
            assert np.array_equal(np.sort(a.refcounts()[live_mask(a.cashes())]), np.sort(t.refcounts()[live_mask(t.cashes())]))
        n8, r8, _ = T.flatten()
        assert_same_hits(oc.trace_rays(n8, r8, depth, o, d), A.trace(o, d), f"box {lo} {hi} {v}")
    for q in rs.randint(0, 128, (5000, 3)):
        assert A.at(*q) == T.at(*q) and E.at(*q) == F.at(*q)
    # single-voxel edits after bulk edits stay consistent (counts were exact)
    ops = np.concatenate([rs.randint(0, 128, (3000, 3)), rs.randint(0, 3, (3000, 1))], 1).astype(np.uint32)
    A.set_many(ops)
    T.set_many(ops)
    assert (A.fillcnt, A.nodecnt) == (T.get_fillcnt(), T.get_nodecnt())


def test_fill_box_delta_stream(ort, oc):
    """fill_box + take_delta drive a mirror exactly like set()-based edits do."""
    depth, log2cap = 6, 16
    h, g = oc.heightmap(depth), oc.grass_bits(depth)
    A = oc.OracleTree(log2cap, depth)
    A.initialize_terrain(h, g, False)
    T = ort.HOctree(log2cap, depth, device=None)
    ort.harness.build_terrain(T, h, g)
    ids, nodes8, root, full = T.take_delta()
    mirror = nodes8.copy()
    rs = np.random.RandomState(3)
    rot, fov = oc.camera_coeffs(0.3, -0.9)
    d = oc.gen_rays(rot, fov, 96, 54)
    o = np.array([1.5, 1.5, 1.8], np.float32)
    for step in range(20):
        c = rs.randint(4, 60, 3)
        e = int(rs.randint(2, 14))
        v = int(rs.randint(0, 3))
        lo, hi = c - e // 2, c + (e + 1) // 2
        T.fill_box(lo, hi, v)
        A.set_many(np.array([(x, y, z, v) for z in range(max(lo[2], 0), min(hi[2], 64)) for y in range(max(lo[1], 0), min(hi[1], 64))
                             for x in range(max(lo[0], 0), min(hi[0], 64))], np.uint32).reshape(-1, 4))
        ids, nodes8, root, full = T.take_delta()
        mirror = nodes8.copy() if full else apply_delta(mirror, ids, nodes8)
        assert_same_hits(oc.trace_rays(mirror, root, depth, o, d), A.trace(o, d), f"step {step}")


def test_table_dump_and_load_round_trip(ort, oc, tmp_path):
    """save()/load(): the restored table is the saved one slot for slot and keeps taking edits like the original."""
    depth, log2cap = 7, 17
    A = ort.HOctree(log2cap, depth, device=None)
    ort.harness.build_terrain(A, tunnels=True, gpu=False)
    rs = np.random.RandomState(5)
    ops = np.concatenate([rs.randint(0, 1 << depth, (2000, 3)), rs.randint(0, 3, (2000, 1))], 1).astype(np.uint32)
    A.set_many(ops)                                     # leaves gravestones behind
    path = str(tmp_path / "table.ort")
    A.save(path)
    B = ort.HOctree(log2cap, depth, device=None)
    B.load(path)
    for get in ("get_root", "get_fillcnt", "get_nodecnt"):
        assert getattr(A, get)() == getattr(B, get)()
    assert np.array_equal(A.cashes(), B.cashes())
    live = A.cashes() != 0
    assert np.array_equal(A.nodes()[live], B.nodes()[live]) and np.array_equal(A.refcounts()[live], B.refcounts()[live])
    more = np.concatenate([rs.randint(0, 1 << depth, (2000, 3)), rs.randint(0, 3, (2000, 1))], 1).astype(np.uint32)
    A.set_many(more)
    B.set_many(more)
    assert np.array_equal(A.cashes(), B.cashes()) and A.get_root() == B.get_root() and A.get_fillcnt() == B.get_fillcnt()
    assert np.array_equal(A.flatten()[0], B.flatten()[0])
    # wrong shape and garbage are refused
    C = ort.HOctree(log2cap + 1, depth, device=None)
    with pytest.raises(Exception):
        C.load(path)
    bad = tmp_path / "bad.ort"
    bad.write_bytes(b"not a table")
    with pytest.raises(Exception):
        B.load(str(bad))
    # a dump is not trusted: a root or an interior child that names no live slot of the table, a record count that
    # does not match, a truncated file -- each is refused and leaves an empty, usable table behind
    raw = bytearray(open(path, "rb").read())
    hdr = 8 + 4 * 2 + 4 * 4 + 8                      # magic, log2cap, depth, root, fillcnt, nodecnt, max_refcnt, records
    rec = 4 + 4 + 4 + 32                             # slot, refcount, tag (padded), children
    assert (len(raw) - hdr) % rec == 0, "the test's idea of the dump layout is out of date"
    cap = 1 << log2cap

    def refused(mutate, what):
        b = bytearray(raw)
        mutate(b)
        q = tmp_path / "corrupt.ort"
        q.write_bytes(bytes(b))
        D = ort.HOctree(log2cap, depth, device=None)
        with pytest.raises(Exception):
            D.load(str(q))
        assert D.get_root() == 0 and D.get_fillcnt() == 0, what
        D.set(1, 2, 3, 4)
        assert D.at(1, 2, 3) == 4, what

    import struct
    refused(lambda b: struct.pack_into("<I", b, 16, cap + 7), "root beyond the table")
    refused(lambda b: struct.pack_into("<Q", b, hdr - 8, (len(raw) - hdr) // rec + 1), "one record more than the file holds")
    refused(lambda b: b.__delitem__(slice(len(b) - rec, len(b))), "truncated")
    # the root's record: point one of its children outside the table, then at an empty slot
    root = struct.unpack_from("<I", raw, 16)[0]
    slots = np.frombuffer(bytes(raw[hdr:]), dtype=np.uint8).reshape(-1, rec)[:, :4].copy().view(np.uint32).ravel()
    k = int(np.flatnonzero(slots == root - 1)[0])
    off = hdr + k * rec + 12
    child_k = next(i for i in range(8) if struct.unpack_from("<I", raw, off + 4 * i)[0] != 0)
    refused(lambda b: struct.pack_into("<I", b, off + 4 * child_k, cap + 1), "interior child beyond the table")
    empty = int(np.setdiff1d(np.arange(cap, dtype=np.uint32), slots)[0]) + 1
    refused(lambda b: struct.pack_into("<I", b, off + 4 * child_k, empty), "interior child names an empty slot")


@pytest.mark.parametrize("depth,log2cap", [(1, 8), (2, 8), (16, 18)])
def test_extreme_depths_match_oracle(ort, oc, depth, log2cap):
    """Smallest and largest tree the interface allows (uint16 coordinates: depth 16 = 65536^3): same table as the
    restatement of the reference after random sets and removals, corners included."""
    rs = np.random.RandomState(depth)
    dim = 1 << depth
    a, p = oc.OracleTree(log2cap, depth), ort.HOctree(log2cap, depth, device=None)
    n = 24 if depth <= 2 else 2500
    pts = rs.randint(0, dim, (n, 3))
    pts[:4] = [[0, 0, 0], [dim - 1, dim - 1, dim - 1], [0, dim - 1, 0], [dim - 1, 0, dim - 1]]
    ops = np.concatenate([pts, rs.randint(1, 5, (n, 1))], 1).astype(np.uint32)
    rem = ops[rs.permutation(n)[: n // 3]].copy()
    rem[:, 3] = 0
    for batch in (ops, rem, ops[: n // 2]):
        a.set_many(batch)
        p.set_many(batch)
        assert (a.root, a.fillcnt, a.nodecnt) == (p.get_root(), p.get_fillcnt(), p.get_nodecnt())
        assert np.array_equal(a.cashes(), p.cashes())
        lm = live_mask(a.cashes())
        assert np.array_equal(a.nodes()[lm], p.nodes()[lm]) and np.array_equal(a.refcounts()[lm], p.refcounts()[lm])
    assert [a.at(*q) for q in pts[:200]] == [p.at(*q) for q in pts[:200]]
    nodes8, root, lo = p.flatten()
    # (a slot that serves both above and at the last level gets two compact ids, so >= rather than ==)
    assert p.get_fillcnt() <= nodes8.shape[0] <= p.get_fillcnt() + 8 and len(lo) == depth + 1


def test_slots_shared_across_levels_flatten_and_delta(ort, oc):
    """The table is content-addressed across levels: with small payloads a last-level node such as (2,0,0,..) has the
    same bytes as an interior node pointing at slot 1, and chains of single-child nodes can meet the same slot at
    different heights.  Compact ids therefore belong to (slot, level); both the fresh flatten and the delta stream must
    keep every interpretation apart.  (Found by the depth-16 case: isolated voxels = long single-child chains.)"""
    depth, log2cap = 16, 18
    rs = np.random.RandomState(116)
    dim = 1 << depth
    A, T = oc.OracleTree(log2cap, depth), ort.HOctree(log2cap, depth, device=None)
    n = 4000
    pts = rs.randint(0, dim, (n, 3))
    pts[: n // 2] = 32768 + rs.randint(-40, 40, (n // 2, 3))
    ops = np.concatenate([pts, rs.randint(1, 5, (n, 1))], 1).astype(np.uint32)
    m = 20000
    o = rs.uniform(1.001, 1.999, (m, 3)).astype(np.float32)
    target = (1.0 + (pts[rs.randint(0, n, m)] + rs.uniform(0, 1, (m, 3))) / dim).astype(np.float32)
    d = target - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)

    A.set_many(ops[:1000])
    T.set_many(ops[:1000])
    ids, nodes8, root, full = T.take_delta()
    assert full
    mirror = nodes8.copy()
    assert_same_hits(oc.trace_rays(mirror, root, depth, o, d), A.trace(o, d), "first flatten")
    for k in range(1000, n, 500):                      # grow through deltas, with some removals
        batch = ops[k:k + 500].copy()
        batch[::7, 3] = 0
        A.set_many(batch)
        T.set_many(batch)
        ids, nodes8, root, full = T.take_delta()
        mirror = nodes8.copy() if full else apply_delta(mirror, ids, nodes8)
        assert_same_hits(oc.trace_rays(mirror, root, depth, o, d), A.trace(o, d), f"delta after {k}")
    fresh, root2, _ = T.flatten()
    assert fresh.shape[0] > T.get_fillcnt()             # some slot really has more than one role in this scene
    assert_same_hits(oc.trace_rays(fresh, root2, depth, o, d), A.trace(o, d), "fresh flatten")


def test_host_rcp_probe_matches_builtin_table_on_this_host(ort, oc):
    """ort_host_rcp_table: the product's own probe of the host's RCPSS.  On the Intel hosts this project runs on it
    reproduces the built-in table exactly and the table model has no mismatch; it also equals the oracle's probe."""
    tab, bad = ort.host_rcp_table(11)
    otab, obad = oc.rcp_table_from_hw(11)
    assert np.array_equal(tab, otab)
    if bad == 0:                                        # Intel-style 11-bit table: must be the built-in one
        import re
        txt = open(os.path.join(ROOT, "octree_ray_tracing_b200", "csrc", "ort_rcp_table.h")).read()
        builtin = np.array([int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)], np.uint32)
        assert np.array_equal(tab, builtin)
    else:                                               # another vendor: a finer table must exist
        assert any(ort.host_rcp_table(k)[1] == 0 for k in (12, 14, 16, 20, 23))


def test_opensimplex_heightmap_option_against_the_reference_class(ort, oc, golden):
    """The demo's alternative terrain noise (OpenSimplexNoise(8789), test_och_h_octree.cpp:33, :568): the product's
    restatement (csrc/ort_opensimplex.h) equals golden values minted from the UNMODIFIED reference class bit for bit --
    sample points for two seeds and the depth-6 heightmap -- and, where oracle/_ref exists, the class itself on fresh
    points and seeds.  A terrain built from that heightmap is the heightmap again when read back through at()."""
    g = golden("opensimplex_8789")
    assert np.array_equal(ort.harness.opensimplex2(g["xy"], 8789).view(np.uint64), g["v8789"].view(np.uint64))
    assert np.array_equal(ort.harness.opensimplex2(g["xy"][:2000], -123456789).view(np.uint64), g["v_neg"].view(np.uint64))
    depth = int(g["depth"])
    h = ort.harness.heightmap(depth, noise="opensimplex")
    assert np.array_equal(h, g["heights"])
    assert not np.array_equal(h, ort.harness.heightmap(depth))              # a different terrain than the live simplex_n one
    if oc.have_ref():
        rs = np.random.RandomState(3)
        xy = rs.uniform(-100, 100, (50000, 2))
        for seed in (8789, 0, 1, 2**40 + 3):
            want = np.zeros(len(xy))
            oc.ref().ochref_opensimplex2(seed, xy.ctypes.data, len(xy), want.ctypes.data)
            assert np.array_equal(ort.harness.opensimplex2(xy, seed).view(np.uint64), want.view(np.uint64)), seed
    T = ort.HOctree(16, depth, device=None)
    ort.harness.build_terrain(T, noise="opensimplex")
    rs = np.random.RandomState(1)
    for x, y in rs.randint(0, 1 << depth, (200, 2)):
        z = int(h[y, x])
        assert T.at(int(x), int(y), z) in (2, 3) and T.at(int(x), int(y), z + 1) == 0 and T.at(int(x), int(y), max(z - 3, 0)) == 1
