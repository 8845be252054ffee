"""Incremental mirror refresh: the GPU mirror follows the CPU index's Add / Vacuum through
kdbgpu_register_nodes / kdbgpu_patch_rows / kdbgpu_remove_nodes / kdbgpu_set_entry instead of a full
kdbgpu_set_graph, and stays identical to a mirror staged from scratch."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rows(g):
    """{(node, level): tuple(neighbours)} of a Graph export."""
    out = {}
    for i in range(1, g.n + 1):
        for l in range(int(g.levels[i]) + 1):
            out[(i, l)] = tuple(int(x) for x in g.row(i, l))
    return out


def _same_graph(gi, g):
    gn, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
    assert (gn, entry, max_level) == (g.n, g.entry, g.max_level)
    assert np.array_equal(levels, g.levels) and np.array_equal(row_off, g.row_off) and np.array_equal(nbrs, g.nbrs)


def test_mirror_follows_single_adds():
    from kektordb_b200 import GpuIndex
    rng = np.random.default_rng(21)
    n0, n1, dim, m = 1500, 1900, 48, 8
    X = rng.standard_normal((n1, dim)).astype(np.float32)
    u = rng.random(n1)
    oi = O.OracleIndex(dim, O.METRIC_COSINE, m, 60, O.ARITH_KERNEL, n1)
    oi.build_batched(X[:n0], u[:n0], batch=500, threads=8)
    g0 = oi.export_graph()
    gi = GpuIndex(dim, "cosine", m, n1)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g0.n, g0.levels, g0.node_row, g0.row_off, g0.nbrs, g0.entry, g0.max_level)
    Q = rng.standard_normal((64, dim)).astype(np.float32)
    prev = _rows(g0)
    pos = n0
    for step in (1, 7, 92, 300):                      # the CPU index keeps inserting (single Add, hnsw_index.go:472)
        for i in range(pos, pos + step):
            oi.add(X[i], u[i])
        g = oi.export_graph()
        cur = _rows(g)
        changed = [key for key, row in cur.items() if prev.get(key) != row]
        gi.upload_vectors(pos + 1, oi.vectors()[pos + 1:pos + step + 1])
        gi.register_nodes(pos + 1, g.levels[pos + 1:pos + step + 1])
        gi.patch_rows([c[0] for c in changed], [c[1] for c in changed], [cur[c] for c in changed])
        gi.set_entry(g.entry, g.max_level)
        assert len(changed) < len(cur)                # a patch, not a re-upload
        _same_graph(gi, g)
        ids, sc, cnt, st = gi.SearchWithScores(Q, 10, None, 64)
        oids, osc, ocnt, ost = oi.search_batch(Q, 10, 64, threads=8)
        assert np.array_equal(ids, oids) and np.array_equal(sc, osc) and st.dist_evals == ost.dist_evals
        prev, pos = cur, pos + step
    gi.close()


def test_mirror_follows_a_vacuum():
    """Vacuum (optimizer.go:118-277): live rows lose their dead neighbours, dead nodes become nil slots, a dead
    entry point is replaced by the first live node with maxLevel = that node's own level."""
    from kektordb_b200 import GpuIndex
    rng = np.random.default_rng(22)
    n, dim, m = 2000, 32, 8
    X = rng.standard_normal((n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, O.METRIC_L2, m, 60, O.ARITH_KERNEL, n)
    oi.build_batched(X, rng.random(n), batch=500, threads=8)
    g = oi.export_graph()
    gi = GpuIndex(dim, "euclidean", m, n)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    dead = set(int(x) for x in rng.choice(np.arange(1, n + 1), 150, replace=False)) | {int(g.entry)}
    rows = _rows(g)
    # the vacuumed topology, built independently: rows without dead ids, dead nodes nil
    levels = g.levels.copy()
    new_rows = {}
    for (i, l), r in rows.items():
        if i in dead:
            continue
        new_rows[(i, l)] = tuple(x for x in r if x not in dead)
    for d in dead:
        levels[d] = -1
    entry = next(i for i in range(1, n + 1) if levels[i] >= 0)
    max_level = int(levels[entry])
    node_row, row_off, nbrs = [0], [0], []
    for i in range(1, n + 1):
        node_row.append(len(row_off) - 1)
        for l in range(int(levels[i]) + 1):
            nbrs.extend(new_rows[(i, l)])
            row_off.append(len(nbrs))
    node_row.append(len(row_off) - 1)
    g2 = O.Graph(n, levels, np.array(node_row, np.uint64), np.array(row_off, np.uint64), np.array(nbrs, np.uint32),
                 np.zeros(n + 1, np.uint8), entry, max_level)
    o2 = O.OracleIndex(dim, O.METRIC_L2, m, 60, O.ARITH_KERNEL, n)
    o2.import_graph(oi.vectors(), g2)
    # the mirror follows with patches
    changed = [key for key, r in new_rows.items() if rows[key] != r]
    gi.patch_rows([c[0] for c in changed], [c[1] for c in changed], [new_rows[c] for c in changed])
    gi.remove_nodes(sorted(dead))
    gi.set_entry(entry, max_level)
    Q = rng.standard_normal((64, dim)).astype(np.float32)
    ids, sc, cnt, st = gi.SearchWithScores(Q, 10, None, 64)
    oids, osc, ocnt, ost = o2.search_batch(Q, 10, 64, threads=8)
    assert np.array_equal(ids, oids) and np.array_equal(sc, osc) and st.dist_evals == ost.dist_evals
    assert not (set(ids.ravel().tolist()) & dead)
    # and equals a mirror staged from scratch
    gj = GpuIndex(dim, "euclidean", m, n)
    gj.upload_vectors(1, oi.vectors()[1:])
    gj.set_graph(g2.n, g2.levels, g2.node_row, g2.row_off, g2.nbrs, g2.entry, g2.max_level)
    jd = gj.SearchWithScores(Q, 10, None, 64)
    assert np.array_equal(ids, jd[0]) and np.array_equal(sc, jd[1])
    gi.close()
    gj.close()


def test_refresh_argument_checks():
    from kektordb_b200 import GpuIndex, ffi
    rng = np.random.default_rng(23)
    gi = GpuIndex(8, "euclidean", 4, 100)
    gi.AddBatch(rng.standard_normal((40, 8)).astype(np.float32), rng.random(40), 10)
    with pytest.raises(ffi.GpuError):
        gi.register_nodes(45, [0])                     # ids are registered in order
    with pytest.raises(ffi.GpuError):
        gi.patch_rows([41], [0], [[1, 2]])             # not a live node
    with pytest.raises(ffi.GpuError):
        gi.patch_rows([1], [0], [list(range(2, 12))])  # more than 2M neighbours
    with pytest.raises(ffi.GpuError):
        gi.set_entry(77, 0)
    gi.register_nodes(41, [0, -1, 1])
    assert gi.count == 43
    gi.close()


def test_searches_and_refresh_calls_from_concurrent_threads():
    """Searches (shared lock, several in flight) race with patch / delete calls (exclusive lock): every search
    must see one consistent state — its result equals the oracle's on the topology before or after the patch."""
    import threading
    from kektordb_b200 import GpuIndex
    rng = np.random.default_rng(31)
    n, dim, m = 2500, 40, 8
    X = rng.standard_normal((n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, O.METRIC_COSINE, m, 60, O.ARITH_KERNEL, n)
    oi.build_batched(X, rng.random(n), batch=500, threads=8)
    g = oi.export_graph()
    rows = _rows(g)
    # state B: 40 level-0 rows lose their last neighbour
    victims = [key for key in list(rows)[::50] if key[1] == 0 and len(rows[key]) > 2][:40]
    new_rows = dict(rows)
    for key in victims:
        new_rows[key] = rows[key][:-1]
    node_row, row_off, nbrs = [0], [0], []
    for i in range(1, n + 1):
        node_row.append(len(row_off) - 1)
        for l in range(int(g.levels[i]) + 1):
            nbrs.extend(new_rows[(i, l)])
            row_off.append(len(nbrs))
    node_row.append(len(row_off) - 1)
    gB = O.Graph(n, g.levels, np.array(node_row, np.uint64), np.array(row_off, np.uint64), np.array(nbrs, np.uint32),
                 np.zeros(n + 1, np.uint8), g.entry, g.max_level)
    oB = O.OracleIndex(dim, O.METRIC_COSINE, m, 60, O.ARITH_KERNEL, n)
    oB.import_graph(oi.vectors(), gB)
    Q = rng.standard_normal((96, dim)).astype(np.float32)
    wantA = oi.search_batch(Q, 10, 64, threads=8)
    wantB = oB.search_batch(Q, 10, 64, threads=8)
    gi = GpuIndex(dim, "cosine", m, n)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    bad, stop = [], threading.Event()

    def searcher():
        while not stop.is_set():
            ids, sc, cnt, _ = gi.SearchWithScores(Q, 10, None, 64)
            okA = np.array_equal(ids, wantA[0]) and np.array_equal(sc, wantA[1])
            okB = np.array_equal(ids, wantB[0]) and np.array_equal(sc, wantB[1])
            if not (okA or okB):
                bad.append(1)

    ts = [threading.Thread(target=searcher) for _ in range(4)]
    for t in ts:
        t.start()
    for flip in range(6):   # A -> B -> A -> ...
        src = new_rows if flip % 2 == 0 else rows
        gi.patch_rows([v[0] for v in victims], [v[1] for v in victims], [src[v] for v in victims])
        gi.set_deleted(None)
    stop.set()
    for t in ts:
        t.join()
    assert not bad
    gi.close()


# ---- the staleness policy (csrc/refresher.cpp): queued changes, one exclusive section per flush ----------------
def _replay_adds(oi, X, u, pos, step, rf, prev):
    """`step` single Adds on the CPU index; every row they rewrote goes to the refresher, change by change."""
    for i in range(pos, pos + step):
        oi.add(X[i], u[i])
        g = oi.export_graph()
        cur = _rows(g)
        nid = i + 1
        rf.add_node(nid, int(g.levels[nid]), oi.vectors()[nid])
        for key, row in cur.items():
            if prev.get(key) != row:
                rf.set_row(key[0], key[1], row)
        rf.set_entry(g.entry, g.max_level)
        prev = cur
    return prev, g


def test_refresher_batches_changes_and_ends_identical_to_a_fresh_mirror():
    from kektordb_b200 import GpuIndex, Refresher
    rng = np.random.default_rng(41)
    n0, n1, dim, m = 1200, 1320, 32, 8
    X = rng.standard_normal((n1, dim)).astype(np.float32)
    u = rng.random(n1)
    oi = O.OracleIndex(dim, O.METRIC_COSINE, m, 60, O.ARITH_KERNEL, n1)
    oi.build_batched(X[:n0], u[:n0], batch=400, threads=8)
    g0 = oi.export_graph()
    gi = GpuIndex(dim, "cosine", m, n1)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g0.n, g0.levels, g0.node_row, g0.row_off, g0.nbrs, g0.entry, g0.max_level)
    Q = rng.standard_normal((48, dim)).astype(np.float32)
    before = gi.SearchWithScores(Q, 10, None, 64)
    rf = Refresher(gi, max_pending_rows=100_000, max_lag_ms=0)       # no clock, no size trigger: flush by call only
    prev, g = _replay_adds(oi, X, u, n0, 60, rf, _rows(g0))
    st = rf.stats()
    assert st.flushes == 0 and st.pending_nodes == 60 and st.pending_rows > 60
    assert st.rows_queued > st.pending_rows                           # hub rows rewritten repeatedly: last write wins
    mid = gi.SearchWithScores(Q, 10, None, 64)                        # the mirror is still the old snapshot
    assert np.array_equal(mid[0], before[0]) and gi.count == n0
    rf.flush()
    _same_graph(gi, g)
    want = oi.search_batch(Q, 10, 64, threads=8)
    got = gi.SearchWithScores(Q, 10, None, 64)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    st = rf.stats()
    assert st.flushes == 1 and st.flushes_by_call == 1 and st.nodes_applied == 60 and st.pending_rows == 0
    # size trigger: a small max_pending_rows flushes on its own while changes stream in
    rf.close()
    rf = Refresher(gi, max_pending_rows=64, max_lag_ms=0)
    prev, g = _replay_adds(oi, X, u, n0 + 60, 60, rf, prev)
    assert rf.stats().flushes_by_rows >= 2
    rf.close()                                                        # destroy flushes the rest
    _same_graph(gi, g)
    want = oi.search_batch(Q, 10, 64, threads=8)
    got = gi.SearchWithScores(Q, 10, None, 64)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    gi.close()


def test_refresher_lag_bound_deletes_and_removals():
    import time
    from kektordb_b200 import GpuIndex, Refresher
    rng = np.random.default_rng(42)
    n, dim, m = 1500, 24, 8
    X = rng.standard_normal((n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, O.METRIC_L2, m, 60, O.ARITH_KERNEL, n)
    oi.build_batched(X, rng.random(n), batch=500, threads=8)
    g = oi.export_graph()
    gi = GpuIndex(dim, "euclidean", m, n)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    Q = rng.standard_normal((40, dim)).astype(np.float32)
    rf = Refresher(gi, max_pending_rows=1 << 20, max_lag_ms=30)
    dead = [int(x) for x in rng.choice(np.arange(1, n + 1), 40, replace=False) if int(x) != g.entry]
    for d in dead:                                                    # Delete = Node.Deleted (hnsw_index.go:2303-2336)
        oi.delete(d)
        rf.set_deleted(d)
    deadline = time.time() + 5
    while rf.stats().flushes_by_lag == 0 and time.time() < deadline:  # nobody calls flush: the clock does
        time.sleep(0.01)
    st = rf.stats()
    assert st.flushes_by_lag >= 1 and st.last_error == 0
    want = oi.search_batch(Q, 10, 64, threads=8)
    got = gi.SearchWithScores(Q, 10, None, 64)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert not np.isin(got[0], dead).any()
    # Vacuum's physical removal of the same nodes: rows naming them are patched, the nodes become nil
    rows = _rows(g)
    for (i, l), r in rows.items():
        if i not in dead and any(x in dead for x in r):
            rf.set_row(i, l, [x for x in r if x not in dead])
    for d in dead:
        rf.remove_node(d)
    rf.flush()
    got2 = gi.SearchWithScores(Q, 10, None, 64)
    assert not np.isin(got2[0], dead).any() and (got2[2] == 10).all()
    rf.close()
    gi.close()
