"""GPU parity for the float16 / int8 precisions: the sm_100a traversal kernel on
rows kept in their stored form (2 / 1 bytes per element), through the C ABI, against the CPU oracle.

Bar: int8 is integer arithmetic end to end (exact int32 dot, float64 scaling) — bit-exact against the
oracle in EVERY arithmetic mode, i.e. against the reference's own pure-Go path.  float16 is bit-exact
against the oracle in kernel order and within 1e-5 (relative to max(1, d)) of the reference orders."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _gpu():
    from kektordb_b200 import GpuIndex, ffi
    assert ffi.lib().kdbgpu_device_count() > 0, "these tests need a CUDA device (no CPU fallback exists)"
    return GpuIndex


def _prec_name(prec):
    return {O.PREC_F16: "float16", O.PREC_I8: "int8", O.PREC_F32: "float32"}[prec]


def _metric_of(prec):
    return O.METRIC_L2 if prec == O.PREC_F16 else O.METRIC_COSINE


def _oracle_index(prec, n, dim, m, efc, seed, data="normal"):
    rng = np.random.default_rng(seed)
    if data == "normal":
        X = rng.standard_normal((n, dim)).astype(np.float32)
    else:  # small integer grid: many exact ties
        X = rng.integers(-2, 3, (n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, _metric_of(prec), m, efc, O.ARITH_KERNEL, n + 8, precision=prec)
    if prec == O.PREC_I8:
        oi.set_quantizer(O.train_quantizer(X))
    oi.build_batched(X, rng.random(n), batch=512, threads=8)
    return oi, X, rng


def _mirror(oi, prec, m, raw=True, X=None):
    GpuIndex = _gpu()
    g = oi.export_graph()
    gi = GpuIndex(oi.dim, "euclidean" if prec == O.PREC_F16 else "cosine", m, max(g.n, 1), precision=_prec_name(prec))
    if prec == O.PREC_I8:
        gi.set_quantizer(oi.abs_max)
    if raw:
        gi.upload_rows_raw(1, oi.rows_raw()[1:])      # the arena's bytes
    else:
        gi.upload_vectors(1, X)                       # float32 in, converted on the device as Add does
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    if g.deleted.any():
        gi.set_deleted(O.dense_bitset(np.where(g.deleted)[0], g.n))
    return gi, g


def _assert_same(gpu, ora):
    gids, gsc, gcnt, gst = gpu
    oids, osc, ocnt, ost = ora
    assert np.array_equal(gcnt, ocnt.astype(np.uint32))
    assert np.array_equal(gids, oids)
    assert np.array_equal(gsc, osc)
    assert gst.dist_evals == ost.dist_evals and gst.hops == ost.hops and gst.hops_l0 == ost.hops_l0


def test_unsupported_pairs_are_rejected():
    from kektordb_b200 import ffi
    GpuIndex = _gpu()
    with pytest.raises(ffi.GpuError):
        GpuIndex(8, "cosine", 4, 8, precision="float16")     # GetFloat16Func: Euclidean only
    with pytest.raises(ffi.GpuError):
        GpuIndex(8, "euclidean", 4, 8, precision="int8")     # GetInt8Func: Cosine only


def test_known_answers_through_the_distance_hook():
    GpuIndex = _gpu()
    gi = GpuIndex(2, "euclidean", 4, 8, precision="float16")
    gi.upload_vectors(1, np.array([[3, 4]], np.float32))
    assert gi.distance_batch(np.array([1, 2], np.float32), [1])[0] == 8.0      # distance_test.go:59-73
    gi.close()
    gi = GpuIndex(2, "cosine", 4, 8, precision="int8")
    gi.set_quantizer(127.0)                                                    # identity scaling: q = round(v)
    gi.upload_vectors(1, np.array([[2, 3], [0, 0]], np.float32))
    got = gi.distance_batch(np.array([10, 20], np.float32), [1, 2])
    qn, sn = np.float32(np.sqrt(500.0)), np.float32(np.sqrt(13.0))
    assert got[0] == 1.0 - 80.0 / (float(qn) * float(sn))                     # dot = 80 (distance_test.go:75-84)
    assert got[1] == 1.0                                                       # stored norm 0 (:2428-2431)
    assert np.array_equal(gi.download_norms(1, 2), np.array([sn, 0], np.float32))
    gi.close()


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
@pytest.mark.parametrize("dim", [1, 5, 16, 33, 100, 128, 257, 768, 1000, 1536])
def test_device_conversion_and_distance_hook_are_bit_exact(prec, dim):
    GpuIndex = _gpu()
    rng = np.random.default_rng(dim)
    X = (rng.standard_normal((300, dim)) * (1.5 if prec == O.PREC_I8 else 1.0)).astype(np.float32)
    X[7] = 0
    gi = GpuIndex(dim, "euclidean" if prec == O.PREC_F16 else "cosine", 4, 300, precision=_prec_name(prec))
    am = O.train_quantizer(X)
    if prec == O.PREC_I8:
        gi.set_quantizer(am)
    gi.upload_vectors(1, X)
    rows = gi.download_rows_raw(1, 300)
    want_rows = O.f32_to_f16_bits(X) if prec == O.PREC_F16 else O.quantize(am, X)
    assert np.array_equal(rows, want_rows)                 # float16.Fromfloat32 / Quantizer.Quantize on the device
    if prec == O.PREC_I8:
        assert np.array_equal(gi.download_norms(1, 300), np.array([O.int8_norm(r) for r in want_rows], np.float32))
    q = rng.standard_normal(dim).astype(np.float32)
    ids = rng.integers(1, 301, 400).astype(np.uint32)
    ids[0] = 8                                             # the all-zero row
    got = gi.distance_batch(q, ids)
    if prec == O.PREC_F16:
        qq = O.f32_to_f16_bits(q)
        want = np.array([float(O.sq_euclid_f16(O.ARITH_KERNEL, qq, want_rows[i - 1])) for i in ids])
        assert np.array_equal(got, want)
        for arith in (O.ARITH_SEQ, O.ARITH_AVX2):
            ref = np.array([float(O.sq_euclid_f16(arith, qq, want_rows[i - 1])) for i in ids[:64]])
            assert np.max(np.abs(got[:64] - ref) / np.maximum(1.0, np.abs(ref))) < 1e-5
    else:
        qq = O.quantize(am, q)
        qn = float(O.int8_norm(qq)) or 1.0
        want = np.array([O.int8_cosine_distance(O.dot_i8(qq, want_rows[i - 1]), qn, float(O.int8_norm(want_rows[i - 1])))
                         for i in ids])
        assert np.array_equal(got, want)                   # integer path: exact in every order
    gi.close()


CONFIGS = [
    # prec, n, dim, M, efC, k, ef, data
    (O.PREC_F16, 4000, 128, 16, 100, 10, 64, "normal"),
    (O.PREC_F16, 2000, 768, 32, 80, 10, 128, "normal"),
    (O.PREC_F16, 1500, 100, 8, 60, 5, 0, "grid"),
    (O.PREC_F16, 800, 1000, 8, 40, 10, 50, "normal"),
    (O.PREC_F16, 600, 1200, 8, 40, 10, 50, "normal"),       # 5 columns per lane: generic path (query in shared memory)
    (O.PREC_I8, 4000, 128, 16, 100, 10, 64, "normal"),
    (O.PREC_I8, 2000, 768, 32, 80, 10, 128, "normal"),      # 768-byte rows: half-filled last column
    (O.PREC_I8, 1500, 1536, 32, 80, 10, 128, "normal"),
    (O.PREC_I8, 1500, 60, 8, 60, 5, 0, "grid"),
    (O.PREC_I8, 800, 3000, 8, 40, 10, 50, "normal"),
    (O.PREC_I8, 600, 2500, 8, 40, 10, 50, "normal"),        # generic path
]


@pytest.mark.parametrize("prec,n,dim,m,efc,k,ef,data", CONFIGS)
def test_search_is_bit_exact(prec, n, dim, m, efc, k, ef, data):
    oi, X, rng = _oracle_index(prec, n, dim, m, efc, 1234 + dim, data)
    gi, g = _mirror(oi, prec, m)
    Q = rng.standard_normal((96, dim)).astype(np.float32) if data == "normal" else \
        rng.integers(-2, 3, (96, dim)).astype(np.float32)
    Q[5] = 0                                               # zero query: int8 qNorm 0 -> 1 (:2410-2413)
    _assert_same(gi.SearchWithScores(Q, k, None, ef), oi.search_batch(Q, k, ef, threads=8))
    if prec == O.PREC_I8:
        # integer arithmetic: the same bits as the reference's pure-Go order, not just kernel order
        oi.set_arith(O.ARITH_SEQ)
        _assert_same(gi.SearchWithScores(Q, k, None, ef), oi.search_batch(Q, k, ef, threads=8))
    else:
        oi.set_arith(O.ARITH_AVX2)
        ids, sc, cnt, _ = gi.SearchWithScores(Q, k, None, ef)
        rid, rsc, rcnt, _ = oi.search_batch(Q, k, ef, threads=8)
        if data == "normal":
            agree = np.mean([len(set(ids[i]) & set(rid[i])) / max(1, int(rcnt[i])) for i in range(len(Q))])
            assert agree >= 0.99
            same = ids == rid
            assert np.max(np.abs(sc[same] - rsc[same]) / np.maximum(1.0, np.abs(rsc[same]))) < 1e-5
    gi.close()


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
def test_search_with_allow_list_and_deletes(prec):
    n, dim, m = 3000, 96, 12
    oi, X, rng = _oracle_index(prec, n, dim, m, 80, 77)
    for d in rng.integers(1, n + 1, 200):
        oi.delete(int(d))
    gi, g = _mirror(oi, prec, m)
    Q = rng.standard_normal((64, dim)).astype(np.float32)
    for sel in (0.5, 0.1, 0.01):
        members = np.where(rng.random(n + 1) < sel)[0]
        members = members[members > 0]
        allow = O.dense_bitset(members, n)
        _assert_same(gi.SearchWithScores(Q, 10, allow, 64), oi.search_batch(Q, 10, 64, allow=allow, threads=8))
    gi.close()


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
def test_float32_upload_equals_raw_upload(prec):
    """kdbgpu_upload_vectors on a float16 / int8 handle converts on the device exactly as Add does, so the
    mirror equals one staged from the arena's bytes."""
    n, dim, m = 1500, 200, 8
    oi, X, rng = _oracle_index(prec, n, dim, m, 60, 99)
    Q = rng.standard_normal((40, dim)).astype(np.float32)
    want = oi.search_batch(Q, 10, 50, threads=8)
    for raw in (True, False):
        gi, _ = _mirror(oi, prec, m, raw=raw, X=X)
        _assert_same(gi.SearchWithScores(Q, 10, None, 50), want)
        gi.close()


def test_quantized_handles_refuse_the_flat_scan():
    from kektordb_b200 import ffi
    oi, X, rng = _oracle_index(O.PREC_I8, 300, 16, 4, 20, 5)
    gi, _ = _mirror(oi, O.PREC_I8, 4)
    with pytest.raises(ffi.GpuError):
        gi.flat_search(X[:2], 3, 0)
    gi.close()
    GpuIndex = _gpu()
    g2 = GpuIndex(16, "cosine", 4, 8, precision="int8")     # no quantizer yet
    with pytest.raises(ffi.GpuError):
        g2.upload_vectors(1, X[:2])
    g2.close()


@pytest.mark.parametrize("n,dim", [(500, 16), (10000, 32), (10001, 24), (60000, 64), (300000, 8)])
def test_train_quantizer_on_the_device_equals_quantizer_train(n, dim):
    """kdbgpu_train_quantizer == Quantizer.Train (quantizer.go:49-125): same sample, same rank, same AbsMax
    bits — including quantizer_test.go's 60 000 x 64 case."""
    GpuIndex = _gpu()
    rng = np.random.default_rng(n)
    X = (rng.random((n, dim)) * 10 - 3).astype(np.float32)
    gi = GpuIndex(dim, "cosine", 4, 16, precision="int8")
    am = gi.TrainQuantizer(X)
    assert np.float32(am) == O.train_quantizer(X) and am > 0
    gi.upload_vectors(1, X[:8])
    assert np.array_equal(gi.download_rows_raw(1, 8), O.quantize(am, X[:8]))
    gi.close()


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
def test_fast_pass_equals_heap_pass_on_quantized_rows(prec):
    """The sorted-list fast pass (default for int8) and the heap pass give the same bits, ties included."""
    oi, X, rng = _oracle_index(prec, 2500, 64, 8, 60, 555, "grid")
    gi, g = _mirror(oi, prec, 8)
    Q = rng.integers(-2, 3, (128, 64)).astype(np.float32)
    want = oi.search_batch(Q, 10, 64, threads=8)
    for mode in (2, 1, 0):
        gi.set_fast_path(mode)
        _assert_same(gi.SearchWithScores(Q, 10, None, 64), want)
    gi.close()


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
def test_wide_beams_small_indexes_and_empty_filters(prec):
    """ef > 128 (heap pass only), k larger than the index, efSearch = 0, an empty and a non-matching allow-list."""
    oi, X, rng = _oracle_index(prec, 60, 40, 4, 30, 31)
    gi, g = _mirror(oi, prec, 4)
    Q = rng.standard_normal((17, 40)).astype(np.float32)
    for k, ef in ((100, 300), (10, 200), (3, 0), (1, 1)):
        _assert_same(gi.SearchWithScores(Q, k, None, ef), oi.search_batch(Q, k, ef, threads=4))
    empty = O.dense_bitset([], g.n)                      # searchInternal returns [] (:443-445)
    ids, sc, cnt, _ = gi.SearchWithScores(Q, 5, empty, 20)
    assert not cnt.any() and not ids.any()
    one = O.dense_bitset([7], g.n)
    _assert_same(gi.SearchWithScores(Q, 5, one, 20), oi.search_batch(Q, 5, 20, allow=one, threads=4))
    ids, sc, cnt, _ = gi.SearchWithScores(Q[:0], 5, None, 20)
    assert ids.shape == (0, 5)
    gi.close()
    big, X2, rng2 = _oracle_index(prec, 3000, 64, 8, 60, 32)
    gb, _ = _mirror(big, prec, 8)
    Q2 = rng2.standard_normal((64, 64)).astype(np.float32)
    _assert_same(gb.SearchWithScores(Q2, 20, None, 250), big.search_batch(Q2, 20, 250, threads=8))
    gb.close()


@pytest.mark.parametrize("name", ["int8_cosine_d40_m8", "f16_l2_d36_m6"])
def test_quantized_golden_fixtures(name):
    """Committed fixtures replayed through the C ABI: stored rows + topology -> the expected output; and the
    float32 inputs converted on the device give the fixture's rows (and norms)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    GpuIndex = _gpu()
    prec, dim, n, m = int(z["precision"]), int(z["dim"]), int(z["n"]), int(z["m"])
    gi = GpuIndex(dim, "euclidean" if prec == O.PREC_F16 else "cosine", m, n, precision=_prec_name(prec))
    if prec == O.PREC_I8:
        assert np.float32(gi.TrainQuantizer(z["inputs"])) == np.float32(z["abs_max"])
    gi.upload_vectors(1, z["inputs"])
    assert np.array_equal(gi.download_rows_raw(1, n), z["rows"][1:])
    if prec == O.PREC_I8:
        assert np.array_equal(gi.download_norms(1, n), z["norms"][1:])
    gi.upload_rows_raw(1, z["rows"][1:])
    gi.set_graph(n, z["levels"], z["node_row"], z["row_off"], z["nbrs"], int(z["entry"]), int(z["max_level"]))
    if z["deleted"].any():
        gi.set_deleted(O.dense_bitset(np.where(z["deleted"])[0], n))
    allow = z["allow"] if z["allow"].size else None
    for mode in (1, 0, 2):
        gi.set_fast_path(mode)
        ids, sc, cnt, st = gi.SearchWithScores(z["queries"], int(z["k"]), allow, int(z["ef"]))
        assert np.array_equal(ids, z["ids"]) and np.array_equal(sc, z["scores"])
        assert np.array_equal(cnt, z["counts"].astype(np.uint32))
        assert st.dist_evals == int(z["dist_evals"]) and st.hops == int(z["hops"])
    gi.close()
