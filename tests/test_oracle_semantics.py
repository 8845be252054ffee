"""Cross-checks the C oracle's traversal against an independent pure-Python restatement
(tests/pyref.py) on small random graphs, including ties, deletes and allow-lists.  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O
from tests import pyref


def _build(n, dim, metric, m, efc, seed, dup_frac=0.0, batch=0):
    rng = np.random.default_rng(seed)
    X = rng.integers(-3, 4, (n, dim)).astype(np.float32) if dup_frac > 0 else rng.standard_normal((n, dim)).astype(np.float32)
    if dup_frac > 0:  # exact duplicates -> exact distance ties
        k = max(1, int(n * dup_frac))
        X[rng.integers(0, n, k)] = X[rng.integers(0, n, k)]
    idx = O.OracleIndex(dim, metric, m, efc, O.ARITH_SEQ, n + 8)
    if batch:
        idx.build_batched(X, rng.random(n), batch=batch, threads=2)
    else:
        idx.add_many(X, rng.random(n))
    return idx, X, rng


def _py_search(idx, g, q, k, ef, allow_set, metric):
    V = idx.vectors()
    qp = O.normalize(q) if metric == O.METRIC_COSINE else np.asarray(q, np.float32)
    levels = {i: int(g.levels[i]) for i in range(1, g.n + 1) if g.levels[i] >= 0}
    deleted = {i for i in range(1, g.n + 1) if g.deleted[i]}
    dist = lambda i: O.distance(metric, O.ARITH_SEQ, qp, V[i])
    rows = lambda i, l: g.row(i, l).tolist()
    return pyref.search(dist, rows, levels, deleted, g.entry, g.max_level, k, ef, allow_set)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 120), m=st.sampled_from([2, 4, 8]),
       metric=st.sampled_from([O.METRIC_L2, O.METRIC_COSINE]), k=st.integers(1, 12), ef=st.integers(0, 40),
       dup=st.sampled_from([0.0, 0.3]), use_allow=st.booleans(), n_del=st.integers(0, 10),
       batch=st.sampled_from([0, 16]))
def test_oracle_search_equals_python_restatement(seed, n, m, metric, k, ef, dup, use_allow, n_del, batch):
    idx, X, rng = _build(n, 6, metric, m, 12, seed, dup, batch)
    for d in rng.integers(1, n + 1, min(n_del, n)):
        idx.delete(int(d))
    g = idx.export_graph()
    allow_set, allow = None, None
    if use_allow:
        allow_set = set(int(i) for i in np.where(rng.random(n + 1) < 0.4)[0] if i > 0)
        allow = O.dense_bitset(sorted(allow_set), n)
    for _ in range(4):
        q = rng.integers(-3, 4, 6).astype(np.float32) if dup else rng.standard_normal(6).astype(np.float32)
        ids, sc = idx.search(q, k, ef, allow=allow)
        ref = _py_search(idx, g, q, k, ef, allow_set, metric)
        assert ids.tolist() == [r[0] for r in ref]
        assert sc.tolist() == [r[1] for r in ref]


def test_batch_build_is_thread_count_invariant():
    rng = np.random.default_rng(5)
    X = rng.standard_normal((1500, 16)).astype(np.float32)
    u = rng.random(1500)
    graphs = []
    for threads in (1, 4):
        idx = O.OracleIndex(16, O.METRIC_COSINE, 8, 40, O.ARITH_KERNEL, 1500)
        idx.build_batched(X, u, batch=256, threads=threads)
        graphs.append(idx.export_graph())
    a, b = graphs
    assert a.entry == b.entry and a.max_level == b.max_level
    assert np.array_equal(a.levels, b.levels) and np.array_equal(a.row_off, b.row_off)
    assert np.array_equal(a.nbrs, b.nbrs)


def test_graph_export_import_roundtrip():
    idx, X, rng = _build(400, 12, O.METRIC_COSINE, 6, 30, 7)
    g = idx.export_graph()
    other = O.OracleIndex(12, O.METRIC_COSINE, 6, 30, O.ARITH_SEQ, 400)
    other.import_graph(idx.vectors(), g)
    Q = rng.standard_normal((20, 12)).astype(np.float32)
    a = idx.search_batch(Q, 5, 20)
    b = other.search_batch(Q, 5, 20)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_smart_entry_point_and_upper_level_filtering():
    """hnsw_index.go:436-447: entry not in the allow-list -> smallest member becomes the entry, even
    if it only exists on level 0; the upper-level searches then return just that entry (:2521-2524)."""
    idx, X, rng = _build(300, 8, O.METRIC_L2, 4, 20, 21)
    g = idx.export_graph()
    assert g.max_level >= 1
    low = [i for i in range(1, g.n + 1) if g.levels[i] == 0 and i != g.entry][:40]
    allow = O.dense_bitset(low, g.n)
    q = X[low[3] - 1]
    ids, sc = idx.search(q, 5, 30, allow=allow)
    assert set(ids.tolist()) <= set(low) and len(ids) > 0
    ref = _py_search(idx, g, q, 5, 30, set(low), O.METRIC_L2)
    assert ids.tolist() == [r[0] for r in ref]


@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_COSINE])
def test_select_neighbors_heuristic(metric):
    """selectNeighbors hnsw_index.go:2629-2701: keep first; keep e iff no kept r is closer to e than
    e is to the base; top up from the discarded list in order."""
    idx, X, rng = _build(60, 5, metric, 4, 20, 33)
    V = idx.vectors()
    base = 7
    cand = [i for i in range(1, 61) if i != base]
    d = np.array([O.distance(metric, O.ARITH_SEQ, V[base], V[i]) for i in cand])
    order = np.argsort(d, kind="stable")
    ids = np.array(cand, dtype=np.uint32)[order]
    ds = d[order]
    got = idx.select_neighbors(ids, ds, 6).tolist()
    res, disc = [], []
    for i, e in enumerate(ids.tolist()):
        if len(res) >= 6:
            break
        if not res:
            res.append(e)
            continue
        good = all(not (O.distance(metric, O.ARITH_SEQ, V[e], V[r]) < ds[i]) for r in res)
        (res if good else disc).append(e)
    res += disc[: 6 - len(res)]
    assert got == res
    assert idx.select_neighbors(ids[:4], ds[:4], 6).tolist() == ids[:4].tolist()  # len <= m: unchanged (:2634-2636)
