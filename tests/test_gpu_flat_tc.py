"""GPU tests of the flat scan's tensor-core pre-filter (KDBGPU_FLAT_PREFILTER, csrc/flat_tc.cu).

The pre-filter only nominates rows; results must be bit-identical to the exhaustive float64 scan
(ids, order, float64 scores, counts), which tests/test_gpu_parity.py pins to the oracle's
BruteForceIndex restatement (reference pkg/core/vector_index.go:104-162).  The tcgen05 pass itself
is checked against numpy: |approximate - exact| score must stay inside the certified bound the
thresholds are built from."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _flat_index(X, metric, m=8, deleted=None):
    """A handle holding rows only (every node on level 0, no links): what the flat path needs."""
    from kektordb_b200 import GpuIndex, ffi
    assert ffi.lib().kdbgpu_device_count() > 0, "these tests need a CUDA device (no CPU fallback exists)"
    n, dim = X.shape
    gi = GpuIndex(dim, metric, m, n)
    rows = X
    if metric == "cosine":
        rows = np.stack([O.normalize(x) for x in X])
    gi.upload_vectors(1, rows)
    levels = np.zeros(n + 1, np.int32)
    levels[0] = -1
    node_row = np.concatenate([[0], np.arange(n + 1)]).astype(np.uint64)
    row_off = np.zeros(n + 1, np.uint64)
    gi.set_graph(n, levels, node_row, row_off, np.zeros(1, np.uint32), 1, 0)
    if deleted is not None:
        gi.set_deleted(O.dense_bitset(deleted, n))
    return gi, rows


def _data(n, dim, seed, lowrank=False):
    rng = np.random.default_rng(seed)
    if lowrank:
        W = rng.standard_normal((16, dim)).astype(np.float32)
        X = (rng.standard_normal((n, 16)).astype(np.float32) @ W + 0.1 * rng.standard_normal((n, dim))).astype(np.float32)
    else:
        X = rng.standard_normal((n, dim)).astype(np.float32)
    return X, rng


@pytest.mark.parametrize("metric,mode,n,dim", [("euclidean", 0, 5000, 128), ("euclidean", 1, 3001, 100),
                                                ("cosine", 1, 4097, 768), ("cosine", 0, 2500, 72)])
def test_tensor_core_scores_within_certified_bound(metric, mode, n, dim):
    X, rng = _data(n, dim, 11)
    Q = rng.standard_normal((37, dim)).astype(np.float32)
    gi, rows = _flat_index(X, metric)
    S, bound = gi.flat_prefilter_scores(Q, mode)
    gi.close()
    R = rows.astype(np.float64)
    if mode == 1 and metric == "cosine":
        Qn = np.stack([O.normalize(q) for q in Q]).astype(np.float64)
        exact = -(Qn @ R.T)                       # d = 1 + score
    else:
        Qd = Q.astype(np.float64)
        exact = (R * R).sum(1)[None, :] - 2.0 * (Qd @ R.T)   # d = score + |q|^2
    err = np.abs(S.astype(np.float64) - exact).max(axis=1)
    assert np.all(np.isfinite(S))
    assert np.all(err <= bound), (err.max(), bound.min())
    # the bound is not vacuous either: within ~50x of what bf16 rounding really does
    assert np.all(bound < 200 * np.maximum(err, 1e-6))


@pytest.mark.parametrize("nq", [256, 300, 1024])
def test_cta_pair_kernel_is_bit_identical_on_full_and_ragged_query_tile_pairs(nq, monkeypatch):
    """Batches of 256+ queries take the cta_group::2 kernel (two CTAs share every corpus tile): full tile pairs,
    a ragged last pair (300 -> 384 padded rows, the pair tile runs past them), and the benchmark's batch of 1024."""
    monkeypatch.delenv("KDBGPU_FLAT_2CTA", raising=False)  # the pair kernel is the default (the switch is read per launch)
    monkeypatch.delenv("KDBGPU_FLAT_DYNAMIC", raising=False)  # and so is its dynamic tile schedule (tile counter + ring)
    n, dim = 30000, 200
    X, rng = _data(n, dim, 77, lowrank=True)
    Q = rng.standard_normal((nq, dim)).astype(np.float32)
    Q[:10] = X[:10]
    gi, _ = _flat_index(X, "euclidean")
    S, bound = gi.flat_prefilter_scores(Q[:300] if nq > 300 else Q, 0)     # raw tensor-core scores of the pair kernel
    Qd, Xd = Q[:S.shape[0]].astype(np.float64), X.astype(np.float64)
    exact = (Xd * Xd).sum(1)[None, :] - 2.0 * (Qd @ Xd.T)
    assert np.all(np.abs(S.astype(np.float64) - exact).max(axis=1) <= bound)
    for k in (10, 100):
        a = gi.flat_search(Q[:64], k, 0)                                   # exhaustive float64 scan (first 64: it is slow)
        b = gi.flat_search(Q, k, 0, prefilter=True)
        assert np.array_equal(a[0], b[0][:64]) and np.array_equal(a[1], b[1][:64]) and np.array_equal(a[2], b[2][:64])
        c = gi.flat_search(Q[64:128], k, 0, prefilter=True)                # 64 queries: the single-CTA kernel
        assert np.array_equal(c[0], b[0][64:128]) and np.array_equal(c[1], b[1][64:128])
        assert b[3].hops <= 2
    monkeypatch.setenv("KDBGPU_FLAT_DYNAMIC", "0")                         # static tile schedule of the pair kernel
    s_ = gi.flat_search(Q, 100, 0, prefilter=True)
    assert np.array_equal(s_[0], b[0]) and np.array_equal(s_[1], b[1]) and np.array_equal(s_[2], b[2])
    monkeypatch.delenv("KDBGPU_FLAT_DYNAMIC")
    monkeypatch.setenv("KDBGPU_FLAT_2CTA", "0")
    d = gi.flat_search(Q, 10, 0, prefilter=True)                           # the single-CTA kernel, same batch
    e = gi.flat_search(Q[:64], 10, 0)
    assert np.array_equal(d[0][:64], e[0]) and np.array_equal(d[1][:64], e[1])
    gi.close()


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("k", [10, 100])
def test_prefilter_is_bit_identical_to_exhaustive_scan(metric, mode, k):
    n, dim = 20000, 96
    X, rng = _data(n, dim, 5 + k, lowrank=True)
    Q = rng.standard_normal((130, dim)).astype(np.float32)
    Q[:10] = X[:10]  # exact matches: distance 0 at rank 0
    gi, _ = _flat_index(X, metric)
    a = gi.flat_search(Q, k, mode)
    b = gi.flat_search(Q, k, mode, prefilter=True)
    gi.close()
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert b[3].hops <= 2, "the certificate should close for (almost) every query"
    assert b[3].dist_evals < 0.2 * Q.shape[0] * n


def test_prefilter_with_allow_list_deleted_rows_and_ties():
    n, dim, k = 30000, 64, 50
    rng = np.random.default_rng(3)
    X = rng.integers(-2, 3, (n, dim)).astype(np.float32)  # integer grid: many exact distance ties
    Q = rng.integers(-2, 3, (64, dim)).astype(np.float32)
    deleted = rng.choice(np.arange(1, n + 1), n // 7, replace=False)
    allow_ids = np.where(rng.random(n + 1) < 0.3)[0]
    allow_ids = allow_ids[allow_ids > 0]
    allow = O.dense_bitset(allow_ids, n)
    gi, _ = _flat_index(X, "euclidean", deleted=deleted)
    for mode in (0, 1):
        a = gi.flat_search(Q, k, mode, allow)
        b = gi.flat_search(Q, k, mode, allow, prefilter=True)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        live = set(allow_ids.tolist()) - set(deleted.tolist())
        assert set(b[0][b[0] > 0].tolist()) <= live
    gi.close()


def test_prefilter_small_corpus_and_k_larger_than_rows():
    X, rng = _data(700, 40, 9)
    Q = rng.standard_normal((5, 40)).astype(np.float32)
    gi, _ = _flat_index(X, "euclidean")
    a = gi.flat_search(Q, 1000, 0)
    b = gi.flat_search(Q, 1000, 0, prefilter=True)
    gi.close()
    assert np.array_equal(a[2], b[2]) and (b[2] == 700).all()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_prefilter_mirror_follows_row_updates():
    X, rng = _data(9000, 64, 21)
    Q = rng.standard_normal((16, 64)).astype(np.float32)
    gi, _ = _flat_index(X, "euclidean")
    b0 = gi.flat_search(Q, 10, 0, prefilter=True)
    gi.upload_vectors(1, Q)  # the queries become rows 1..16: each must now find itself at distance 0
    b1 = gi.flat_search(Q, 10, 0, prefilter=True)
    a1 = gi.flat_search(Q, 10, 0)
    gi.close()
    assert not np.array_equal(b0[0], b1[0])
    assert np.array_equal(a1[0], b1[0]) and np.array_equal(a1[1], b1[1])
    assert np.array_equal(b1[0][:, 0], np.arange(1, 17, dtype=np.uint32)) and (b1[1][:, 0] == 0).all()
