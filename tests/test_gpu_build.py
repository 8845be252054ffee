"""Device-side graph construction (kdbgpu_add_batch) against the oracle's
restatement of (*Index).AddBatch / addBatchInternal (hnsw_index.go:1466-2088): the adjacency is
identical after every call — sequential single-Add fallback (:1502-1513) and batch path alike."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _compare(n, dim, metric, m, efc, batches, seed, data="normal"):
    from kektordb_b200 import GpuIndex
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, dim)) if data == "normal" else rng.integers(-2, 3, (n, dim))).astype(np.float32)
    u = rng.random(n)
    om = O.METRIC_COSINE if metric == "cosine" else O.METRIC_L2
    oi = O.OracleIndex(dim, om, m, efc, O.ARITH_KERNEL, n)
    gi = GpuIndex(dim, metric, m, n)
    pos = 0
    for b in batches:
        b = min(b, n - pos)
        if b <= 0:
            break
        gi.AddBatch(X[pos:pos + b], u[pos:pos + b], efc)
        oi.add_batch(X[pos:pos + b], u[pos:pos + b], efc, threads=8)
        pos += b
        g = oi.export_graph()
        gn, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
        assert (gn, entry, max_level) == (g.n, g.entry, g.max_level)
        assert np.array_equal(levels, g.levels)
        assert np.array_equal(row_off, g.row_off)
        assert np.array_equal(nbrs, g.nbrs)
    assert np.array_equal(gi.download_vectors(1, n), oi.vectors()[1:])  # normalize() is bit-exact too
    return gi, oi, X, rng


def test_sequential_fallback_and_batches_small():
    gi, oi, X, rng = _compare(300, 16, "euclidean", 4, 20, [10, 5, 5, 30, 50, 200], 1)
    gi.close()


def test_cosine_batches():
    gi, oi, X, rng = _compare(1500, 32, "cosine", 8, 40, [40, 60, 100, 300, 1000], 2)
    # and the graph it built is searched identically by both sides
    Q = rng.standard_normal((64, 32)).astype(np.float32)
    ids, sc, cnt, st = gi.SearchWithScores(Q, 10, None, 50)
    oids, osc, ocnt, ost = oi.search_batch(Q, 10, 50, threads=8)
    assert np.array_equal(ids, oids) and np.array_equal(sc, osc)
    gi.close()


def test_exact_ties_everywhere():
    gi, oi, X, rng = _compare(1200, 12, "euclidean", 6, 30, [30, 70, 300, 800], 3, data="grid")
    gi.close()


def test_reference_defaults_m16_efc200():
    gi, oi, X, rng = _compare(5000, 128, "cosine", 16, 200, [200, 300, 500, 1000, 3000], 4)
    gi.close()


def test_hub_rows_longer_than_the_shared_memory_list():
    """A batch much larger than the graph: some rows receive > 1024 requests and take the
    global-scratch path of the commit kernel."""
    gi, oi, X, rng = _compare(4000, 8, "euclidean", 4, 24, [24, 3976], 5)
    gi.close()


def test_add_after_import():
    """set_graph (import of a CPU-built index) followed by AddBatch keeps extending the same graph."""
    from kektordb_b200 import GpuIndex
    rng = np.random.default_rng(6)
    X = rng.standard_normal((2000, 24)).astype(np.float32)
    u = rng.random(2000)
    oi = O.OracleIndex(24, O.METRIC_COSINE, 8, 40, O.ARITH_KERNEL, 2000)
    oi.build_batched(X[:1200], u[:1200], batch=300, threads=4)
    g = oi.export_graph()
    gi = GpuIndex(24, "cosine", 8, 2000)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    gi.AddBatch(X[1200:], u[1200:], 40)
    oi.add_batch(X[1200:], u[1200:], 40, threads=4)
    g = oi.export_graph()
    gn, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
    assert (gn, entry, max_level) == (g.n, g.entry, g.max_level)
    assert np.array_equal(levels, g.levels) and np.array_equal(row_off, g.row_off) and np.array_equal(nbrs, g.nbrs)
    gi.close()


def test_capacity_is_enforced():
    from kektordb_b200 import GpuIndex, ffi
    gi = GpuIndex(8, "euclidean", 4, 50)
    rng = np.random.default_rng(7)
    gi.AddBatch(rng.standard_normal((40, 8)).astype(np.float32), rng.random(40), 10)
    with pytest.raises(ffi.GpuError):
        gi.AddBatch(rng.standard_normal((20, 8)).astype(np.float32), rng.random(20), 10)
    assert gi.count == 40
    gi.close()


# ---- float16 / int8 indexes: construction with the precision's own distances (what DB.Compress does:
# ---- TrainQuantizer, then AddBatch into a new index of that precision, pkg/core/core.go:1210-1270)
def _compare_quantized(prec, n, dim, m, efc, batches, seed, data="normal"):
    from kektordb_b200 import GpuIndex
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, dim)) if data == "normal" else rng.integers(-2, 3, (n, dim))).astype(np.float32)
    u = rng.random(n)
    oprec, metric, om = (O.PREC_F16, "euclidean", O.METRIC_L2) if prec == "float16" else (O.PREC_I8, "cosine", O.METRIC_COSINE)
    oi = O.OracleIndex(dim, om, m, efc, O.ARITH_KERNEL, n, precision=oprec)
    gi = GpuIndex(dim, metric, m, n, precision=prec)
    if prec == "int8":
        am = gi.TrainQuantizer(X)                       # on the device
        assert np.float32(am) == O.train_quantizer(X)
        oi.set_quantizer(am)
    pos = 0
    for b in batches:
        b = min(b, n - pos)
        if b <= 0:
            break
        gi.AddBatch(X[pos:pos + b], u[pos:pos + b], efc)
        oi.add_batch(X[pos:pos + b], u[pos:pos + b], efc, threads=8)
        pos += b
        g = oi.export_graph()
        gn, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
        assert (gn, entry, max_level) == (g.n, g.entry, g.max_level)
        assert np.array_equal(levels, g.levels)
        assert np.array_equal(row_off, g.row_off)
        assert np.array_equal(nbrs, g.nbrs)
    assert np.array_equal(gi.download_rows_raw(1, n), oi.rows_raw()[1:])
    if prec == "int8":
        assert np.array_equal(gi.download_norms(1, n), oi.norms()[1:])
    Q = rng.standard_normal((48, dim)).astype(np.float32)
    ids, sc, cnt, st = gi.SearchWithScores(Q, 10, None, 50)
    oids, osc, ocnt, ost = oi.search_batch(Q, 10, 50, threads=8)
    assert np.array_equal(ids, oids) and np.array_equal(sc, osc) and st.dist_evals == ost.dist_evals
    gi.close()


@pytest.mark.parametrize("prec", ["float16", "int8"])
def test_quantized_sequential_fallback_and_batches(prec):
    _compare_quantized(prec, 400, 24, 4, 20, [10, 5, 5, 30, 50, 300], 11)


@pytest.mark.parametrize("prec,dim", [("float16", 128), ("int8", 128), ("float16", 768), ("int8", 768), ("int8", 200)])
def test_quantized_batches(prec, dim):
    _compare_quantized(prec, 2000 if dim < 768 else 1200, dim, 8, 40, [40, 60, 100, 300, 2000], 12 + dim)


@pytest.mark.parametrize("prec", ["float16", "int8"])
def test_quantized_exact_ties(prec):
    _compare_quantized(prec, 1000, 12, 6, 30, [30, 70, 300, 600], 13, data="grid")


def test_int8_construction_needs_a_trained_quantizer():
    from kektordb_b200 import GpuIndex, ffi
    gi = GpuIndex(8, "cosine", 4, 50, precision="int8")
    rng = np.random.default_rng(7)
    with pytest.raises(ffi.GpuError):
        gi.AddBatch(rng.standard_normal((10, 8)).astype(np.float32), rng.random(10), 10)
    gi.close()


def test_failed_add_batch_leaves_the_mirror_as_it_was():
    """A recoverable failure inside kdbgpu_add_batch must not leave half-registered nodes behind.  At 8192-d the batch
    path's construction search (8 row slots of 32 KB) does not fit shared memory while the single-Add fallback
    (4 slots) does: the first efConstruction nodes go in, the next batch fails AFTER the library has registered its
    ids — and the id range, levels, entry point and every search answer must be those of the last successful call."""
    from kektordb_b200 import GpuIndex, ffi
    n0, dim, m, efc = 25, 8192, 4, 20
    rng = np.random.default_rng(11)
    X = rng.standard_normal((80, dim)).astype(np.float32)
    u = rng.random(80)
    oi = O.OracleIndex(dim, O.METRIC_L2, m, efc, O.ARITH_KERNEL, 80)
    gi = GpuIndex(dim, "euclidean", m, 80)
    gi.AddBatch(X[:n0], u[:n0], efc)  # n < efConstruction: the sequential single-Add path (:1502-1513)
    oi.add_batch(X[:n0], u[:n0], efc, threads=4)
    before = gi.get_graph()
    with pytest.raises(ffi.GpuError):
        gi.AddBatch(X[n0:], u[n0:], efc)
    after = gi.get_graph()
    assert before[0] == after[0] == n0 and before[5:] == after[5:]
    assert all(np.array_equal(a, b) for a, b in zip(before[1:5], after[1:5]))
    Q = rng.standard_normal((8, dim)).astype(np.float32)
    got = gi.SearchWithScores(Q, 5, None, 20)
    want = oi.search_batch(Q, 5, 20, threads=4)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    gi.close()
