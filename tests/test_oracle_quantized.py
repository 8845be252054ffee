"""float16 / int8 precisions of the CPU oracle, pinned against the reference's
known answers and independent restatements.  CPU only.

Reference anchors: pkg/core/distance/distance_go.go:93-118 (squaredEuclideanGoFloat16,
dotProductGoInt8), pkg/core/distance/distance_test.go:59-84 and native/compute/src/lib.rs:438-458
(known answers), pkg/core/distance/quantizer.go:49-176 (+ quantizer_test.go), pkg/core/hnsw/
hnsw_index.go:2398-2449 (int8 distFn), :3371-3377 (computeInt8Norm), :417-434 (query adaptation)."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests import pyref


def _fma32(a, b, c):
    """Correctly rounded float32 fma via exact rationals (Python 3.12 has no math.fma)."""
    from fractions import Fraction
    exact = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    mid = np.float32(float(exact))
    cands = [np.nextafter(mid, np.float32(-np.inf)), mid, np.nextafter(mid, np.float32(np.inf))]
    best = min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(np.float32(v).view(np.uint32)) & 1))
    return np.float32(best)


# ---- known answers the reference's own tests hold ---------------------------------------------
def test_reference_known_answers_f16_and_int8():
    a, b = O.f32_to_f16_bits([1, 2]), O.f32_to_f16_bits([3, 4])
    for arith in (O.ARITH_SEQ, O.ARITH_AVX2, O.ARITH_KERNEL):
        assert O.sq_euclid_f16(arith, a, b) == 8.0          # distance_test.go:59-73, lib.rs:438-444
    assert O.dot_i8([10, 20], [2, 3]) == 80                 # distance_test.go:75-84, lib.rs:447-452
    assert O.dot_i8([-1, -2], [-1, -2]) == 5                # lib.rs:454-457


def test_quantizer_training_sampling():
    """quantizer_test.go:8-35: 60 000 x 64 uniform [0, 10): Train must sample and yield AbsMax > 0;
    the 99.9th percentile of U[0, 10) is 9.99."""
    rng = np.random.default_rng(0)
    X = (rng.random((60000, 64)) * 10).astype(np.float32)
    am = float(O.train_quantizer(X))
    assert am > 0 and abs(am - 9.99) < 0.01
    # restated independently: stride sample of n/10 capped at 25 000 (floor 10 000), step = n // target
    target = min(max(60000 // 10, 10000), 25000)
    step = 60000 // target
    sample = np.abs(X[::step][:target]).ravel()
    sample.sort()
    assert am == sample[min(int(sample.size * 0.999), sample.size - 1)]
    # at or below 10 000 vectors everything is used
    Y = rng.standard_normal((500, 16)).astype(np.float32)
    s = np.sort(np.abs(Y).ravel())
    assert float(O.train_quantizer(Y)) == s[int(s.size * 0.999)]


# ---- conversions -----------------------------------------------------------------------------
def test_f16_conversion_is_ieee_round_to_nearest_even():
    rng = np.random.default_rng(1)
    parts = [rng.standard_normal(20000).astype(np.float32) * s for s in (1, 1e-3, 1e-5, 1e-7, 1e3, 1e5)]
    edge = np.array([0, -0.0, 65504, 65519.99, 65520, 65536, 1e-8, 2.0 ** -25, 2.0 ** -24, 2.0 ** -25 * 1.0001,
                     5.96e-8, 6.1e-5, 6.0e-5, 1 + 2.0 ** -11, 1 + 3 * 2.0 ** -11, np.inf, -np.inf], np.float32)
    x = np.concatenate(parts + [edge])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)         # numpy's cast is IEEE RNE
    assert np.array_equal(O.f32_to_f16_bits(x), want)
    allbits = np.arange(65536, dtype=np.uint16)
    back, ref = O.f16_bits_to_f32(allbits), allbits.view(np.float16).astype(np.float32)
    assert np.array_equal(back[~np.isnan(ref)], ref[~np.isnan(ref)]) and np.isnan(back[np.isnan(ref)]).all()


def test_quantize_matches_python_restatement():
    """Quantizer.Quantize (quantizer.go:135-160): f32 divide, f32 multiply by 127, clip, round half
    away from zero."""
    rng = np.random.default_rng(2)
    v = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 2,
                        np.array([0, 0.5 / 127 * 3, -0.5 / 127 * 3, 1.5 / 127 * 3, 3, -3, 100, -100], np.float32)])
    am = np.float32(3.0)
    got = O.quantize(am, v)
    want = []
    for x in v:
        s = np.float32(np.float32(x) / am) * np.float32(127.0)
        s = min(max(s, np.float32(-127.0)), np.float32(127.0))
        f = float(s)
        want.append(int(math.floor(abs(f) + 0.5)) * (1 if f >= 0 else -1))
    assert np.array_equal(got, np.array(want, np.int8))
    assert not O.quantize(0.0, v).any()                      # AbsMax == 0 -> zeros (:140-142)
    q = O.quantize(am, v)
    assert O.int8_norm(q) == np.float32(math.sqrt(float((q.astype(np.int64) ** 2).sum())))


# ---- distances -------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 7, 8, 9, 31, 255, 256, 257, 768, 1000])
def test_f16_distance_orders(dim):
    rng = np.random.default_rng(dim)
    a = O.f32_to_f16_bits(rng.standard_normal(dim).astype(np.float32))
    b = O.f32_to_f16_bits(rng.standard_normal(dim).astype(np.float32))
    fa, fb = O.f16_bits_to_f32(a), O.f16_bits_to_f32(b)
    # pure-Go order == sequential f32 loop over the widened halves
    s = np.float32(0)
    for x, y in zip(fa, fb):
        d = np.float32(x - y)
        s = np.float32(s + np.float32(d * d))
    assert O.sq_euclid_f16(O.ARITH_SEQ, a, b) == s
    # the three orders agree within the north-star tolerance
    exact = float(((fa.astype(np.float64) - fb.astype(np.float64)) ** 2).sum())
    for arith in (O.ARITH_SEQ, O.ARITH_AVX2, O.ARITH_KERNEL):
        assert abs(float(O.sq_euclid_f16(arith, a, b)) - exact) <= 1e-5 * max(1.0, exact)
    # kernel order restated: 256 accumulators, lane tree of 8, butterfly
    acc = [np.float32(0)] * 256
    for e, (x, y) in enumerate(zip(fa, fb)):
        d = np.float32(x - y)
        acc[e & 255] = _fma32(d, d, acc[e & 255])
    lane = []
    for l in range(32):
        p = acc[8 * l:8 * l + 8]
        f = lambda u, v: np.float32(u + v)
        lane.append(f(f(f(p[0], p[1]), f(p[2], p[3])), f(f(p[4], p[5]), f(p[6], p[7]))))
    w = 16
    while w >= 1:
        for l in range(w):
            lane[l] = np.float32(lane[l] + lane[l + w])
        w >>= 1
    assert O.sq_euclid_f16(O.ARITH_KERNEL, a, b) == lane[0]


def test_int8_cosine_distance_rules():
    """hnsw_index.go:2421-2449: stored norm 0 -> 1.0; similarity clamped to [-1, 1]; float64 divide."""
    assert O.int8_cosine_distance(50, 10.0, 0.0) == 1.0
    assert O.int8_cosine_distance(200, 10.0, 10.0) == 0.0            # clamp at +1
    assert O.int8_cosine_distance(-200, 10.0, 10.0) == 2.0           # clamp at -1
    qn, sn = np.float32(math.sqrt(30.0)), np.float32(math.sqrt(77.0))
    assert O.int8_cosine_distance(47, qn, sn) == 1.0 - 47.0 / (float(qn) * float(sn))


# ---- index semantics -------------------------------------------------------------------------
def _quant_index(prec, n, dim, m, efc, seed, arith=O.ARITH_SEQ):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    metric = O.METRIC_L2 if prec == O.PREC_F16 else O.METRIC_COSINE
    oi = O.OracleIndex(dim, metric, m, efc, arith, n + 8, precision=prec)
    if prec == O.PREC_I8:
        oi.set_quantizer(O.train_quantizer(X))
    return oi, X, rng, metric


def test_unsupported_metric_precision_pairs_are_rejected():
    """GetFloat16Func / GetInt8Func (distance_go.go:159-177): float16 is Euclidean only, int8 Cosine only."""
    with pytest.raises(ValueError):
        O.OracleIndex(8, O.METRIC_COSINE, precision=O.PREC_F16)
    with pytest.raises(ValueError):
        O.OracleIndex(8, O.METRIC_L2, precision=O.PREC_I8)


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
def test_stored_rows_are_the_converted_inputs(prec):
    oi, X, rng, metric = _quant_index(prec, 300, 20, 4, 30, 5)
    oi.build_batched(X, rng.random(300), batch=100)
    rows = oi.rows_raw()[1:]
    if prec == O.PREC_F16:
        assert np.array_equal(rows, O.f32_to_f16_bits(X))             # not normalised: Euclidean
    else:
        # cosine + int8 is NOT normalised at insert (only cosine + float32 is, hnsw_index.go:485)
        assert np.array_equal(rows, O.quantize(oi.abs_max, X))
        assert np.array_equal(oi.norms()[1:], np.array([O.int8_norm(r) for r in rows], np.float32))


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
@pytest.mark.parametrize("batch", [0, 64])
def test_quantized_search_equals_python_restatement(prec, batch):
    """Traversal over a float16 / int8 index against tests/pyref.py with distances restated in numpy."""
    n, dim = 400, 12
    oi, X, rng, metric = _quant_index(prec, n, dim, 4, 24, 11 + batch)
    if batch:
        oi.build_batched(X, rng.random(n), batch=batch, threads=2)
    else:
        oi.add_many(X, rng.random(n))
    for d in rng.integers(1, n + 1, 15):
        oi.delete(int(d))
    g = oi.export_graph()
    rows = oi.rows_raw()
    levels = {i: int(g.levels[i]) for i in range(1, g.n + 1) if g.levels[i] >= 0}
    deleted = {i for i in range(1, g.n + 1) if g.deleted[i]}
    allow_set = set(int(i) for i in np.where(rng.random(n + 1) < 0.5)[0] if i > 0)
    allow = O.dense_bitset(sorted(allow_set), n)
    for t in range(6):
        q = rng.standard_normal(dim).astype(np.float32)
        if prec == O.PREC_F16:
            qq = O.f32_to_f16_bits(q)
            dist = lambda i: float(O.sq_euclid_f16(O.ARITH_SEQ, qq, rows[i]))
        else:
            qq = O.quantize(oi.abs_max, O.normalize(q))               # :406-414 then :431-434
            qn = float(O.int8_norm(qq)) or 1.0
            norms = oi.norms()
            dist = lambda i: O.int8_cosine_distance(int(qq.astype(np.int32) @ rows[i].astype(np.int32)), qn, norms[i])
        for al_set, al in ((None, None), (allow_set, allow)):
            want = pyref.search(dist, lambda i, l: g.row(i, l).tolist(), levels, deleted, g.entry, g.max_level,
                                5, 20, al_set)
            ids, sc = oi.search(q, 5, 20, allow=al)
            assert [int(i) for i in ids] == [w[0] for w in want]
            assert [float(s) for s in sc] == [w[1] for w in want]


@pytest.mark.parametrize("prec", [O.PREC_F16, O.PREC_I8])
def test_quantized_recall_against_exact_float32(prec):
    """The precisions are approximations of the float32 index: recall@10 vs the exact scan stays high
    on clustered data (the reference's own gate is recall >= 0.95 on float32, stress_test_recall.py)."""
    rng = np.random.default_rng(3)
    n, dim = 4000, 32
    W = rng.standard_normal((6, dim)).astype(np.float32)
    X = (rng.standard_normal((n, 6)).astype(np.float32) @ W + 0.05 * rng.standard_normal((n, dim))).astype(np.float32)
    metric = O.METRIC_L2 if prec == O.PREC_F16 else O.METRIC_COSINE
    oi = O.OracleIndex(dim, metric, 16, 100, O.ARITH_AVX2, n, precision=prec)
    if prec == O.PREC_I8:
        oi.set_quantizer(O.train_quantizer(X))
    oi.build_batched(X, rng.random(n), batch=1000, threads=4)
    f = O.OracleIndex(dim, metric, 16, 100, O.ARITH_AVX2, n)
    f.import_graph(np.vstack([np.zeros((1, dim), np.float32), O.normalize_rows(X) if metric == O.METRIC_COSINE else X]),
                   oi.export_graph())
    Q = (rng.standard_normal((100, 6)).astype(np.float32) @ W).astype(np.float32)
    ids, _, _, _ = oi.search_batch(Q, 10, 100, threads=4)
    gt, _, _ = f.flat_search_batch(Q, 10, 1, threads=4)
    rec = np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(len(Q))])
    assert rec >= (0.95 if prec == O.PREC_F16 else 0.85), rec


def test_flat_scan_is_float32_only():
    oi, X, rng, _ = _quant_index(O.PREC_F16, 50, 8, 4, 10, 1)
    oi.add_many(X, rng.random(50))
    ids, sc, cnt = oi.flat_search_batch(X[:2], 3, 0)
    assert not cnt.any()


# ---- committed fixtures (tests/golden/make_golden.py) ------------------------------------------------
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["int8_cosine_d40_m8", "f16_l2_d36_m6"])
def test_quantized_golden_fixtures_replay(name):
    """The committed float16 / int8 fixtures: stored rows follow from the float32 inputs by an independent
    numpy restatement of the conversions, and the oracle replays the expected output from rows + topology —
    for int8 in the reference's pure-Go order as well (integer dot: every order gives the same bits)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    prec, dim, n = int(z["precision"]), int(z["dim"]), int(z["n"])
    X = z["inputs"]
    if prec == O.PREC_F16:
        with np.errstate(over="ignore"):
            assert np.array_equal(z["rows"][1:], X.astype(np.float16).view(np.uint16))
    else:
        am = np.float32(z["abs_max"])
        s = np.sort(np.abs(X).ravel())
        assert am == s[min(int(s.size * 0.999), s.size - 1)]                       # Quantizer.Train, n <= 10 000
        scaled = np.clip((X / am).astype(np.float32) * np.float32(127.0), -127.0, 127.0).astype(np.float64)
        want = (np.sign(scaled) * np.floor(np.abs(scaled) + 0.5)).astype(np.int8)  # math.Round
        assert np.array_equal(z["rows"][1:], want)
        assert np.array_equal(z["norms"][1:], np.sqrt((want.astype(np.int64) ** 2).sum(1).astype(np.float64)).astype(np.float32))
    g = O.Graph(n, z["levels"], z["node_row"], z["row_off"], z["nbrs"], z["deleted"], int(z["entry"]), int(z["max_level"]))
    allow = z["allow"] if z["allow"].size else None
    for arith in ((O.ARITH_KERNEL, O.ARITH_SEQ, O.ARITH_AVX2) if prec == O.PREC_I8 else (O.ARITH_KERNEL,)):
        oi = O.OracleIndex(dim, int(z["metric"]), int(z["m"]), 50, arith, n, precision=prec)
        if prec == O.PREC_I8:
            oi.set_quantizer(float(z["abs_max"]))
        oi.import_graph(z["rows"], g)
        for d in np.where(z["deleted"])[0]:
            oi.delete(int(d))
        ids, sc, cnt, st = oi.search_batch(z["queries"], int(z["k"]), int(z["ef"]), allow=allow, threads=4)
        assert np.array_equal(ids, z["ids"]) and np.array_equal(sc, z["scores"]) and np.array_equal(cnt, z["counts"])
        assert st.dist_evals == int(z["dist_evals"]) and st.hops == int(z["hops"])


# ---- randomized cross-check against the pure-Python restatement (ties, deletes, filters) -----------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(2, 90), m=st.sampled_from([2, 4, 8]),
       prec=st.sampled_from([O.PREC_F16, O.PREC_I8]), k=st.integers(1, 10), ef=st.integers(0, 30),
       grid=st.booleans(), use_allow=st.booleans(), n_del=st.integers(0, 8), batch=st.sampled_from([0, 12]))
def test_quantized_oracle_equals_python_restatement(seed, n, m, prec, k, ef, grid, use_allow, n_del, batch):
    rng = np.random.default_rng(seed)
    dim = 6
    X = (rng.integers(-3, 4, (n, dim)) if grid else rng.standard_normal((n, dim))).astype(np.float32)
    metric = O.METRIC_L2 if prec == O.PREC_F16 else O.METRIC_COSINE
    oi = O.OracleIndex(dim, metric, m, 10, O.ARITH_SEQ, n + 4, precision=prec)
    if prec == O.PREC_I8:
        oi.set_quantizer(max(float(O.train_quantizer(X)), 1e-3))
    if batch:
        oi.build_batched(X, rng.random(n), batch=batch, threads=2)
    else:
        oi.add_many(X, rng.random(n))
    for d in rng.integers(1, n + 1, min(n_del, n)):
        oi.delete(int(d))
    g = oi.export_graph()
    rows = oi.rows_raw()
    levels = {i: int(g.levels[i]) for i in range(1, g.n + 1) if g.levels[i] >= 0}
    deleted = {i for i in range(1, g.n + 1) if g.deleted[i]}
    allow_set = allow = None
    if use_allow:
        allow_set = set(int(i) for i in np.where(rng.random(n + 1) < 0.4)[0] if i > 0)
        allow = O.dense_bitset(sorted(allow_set), n)
    norms = oi.norms() if prec == O.PREC_I8 else None
    for _ in range(3):
        q = (rng.integers(-3, 4, dim) if grid else rng.standard_normal(dim)).astype(np.float32)
        if prec == O.PREC_F16:
            qq = O.f32_to_f16_bits(q)
            dist = lambda i: float(O.sq_euclid_f16(O.ARITH_SEQ, qq, rows[i]))
        else:
            qq = O.quantize(oi.abs_max, O.normalize(q))
            qn = float(O.int8_norm(qq)) or 1.0
            dist = lambda i: O.int8_cosine_distance(int(qq.astype(np.int32) @ rows[i].astype(np.int32)), qn, norms[i])
        want = pyref.search(dist, lambda i, l: g.row(i, l).tolist(), levels, deleted, g.entry, g.max_level, k, ef, allow_set)
        ids, sc = oi.search(q, k, ef, allow=allow)
        assert [int(i) for i in ids] == [w[0] for w in want]
        assert [float(s) for s in sc] == [w[1] for w in want]
