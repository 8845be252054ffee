"""Generates tests/golden/*.npz with the CPU oracle (KDBO_ARITH_KERNEL): a small graph, its stored
rows, queries and the expected SearchWithScores output.  The GPU parity tests replay these without
rebuilding anything.  These are REGRESSION fixtures written by this repo's own oracle, not vectors the reference
holds: what pins the oracle to the reference is tests/test_oracle_golden.py (the known answers of the reference's
own tests).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, n, dim, metric, m, efc, k, ef, seed, allow_frac=0.0, del_frac=0.0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    X[rng.integers(0, n, n // 20)] = X[rng.integers(0, n, n // 20)]  # exact duplicates -> ties
    Q = np.concatenate([rng.standard_normal((24, dim)).astype(np.float32), X[rng.integers(0, n, 8)]])
    idx = O.OracleIndex(dim, metric, m, efc, O.ARITH_KERNEL, n)
    idx.build_batched(X, rng.random(n), batch=256, threads=4)
    dele = rng.choice(np.arange(1, n + 1), int(n * del_frac), replace=False)
    for d in dele:
        idx.delete(int(d))
    g = idx.export_graph()
    allow = None
    if allow_frac > 0:
        allow = O.dense_bitset(np.where(rng.random(n + 1) < allow_frac)[0][1:], n)
    ids, sc, cnt, st = idx.search_batch(Q, k, ef, allow=allow, threads=4)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), dim=dim, metric=metric, m=m, k=k, ef=ef, n=g.n,
        vectors=idx.vectors(), levels=g.levels, node_row=g.node_row, row_off=g.row_off, nbrs=g.nbrs,
        deleted=g.deleted, entry=g.entry, max_level=g.max_level, queries=Q,
        allow=np.zeros(0, np.uint64) if allow is None else allow, ids=ids, scores=sc, counts=cnt,
        dist_evals=st.dist_evals, hops=st.hops)
    print(name, "rows", g.n, "evals", st.dist_evals, "hops", st.hops)


def make_quantized(name, prec, n, dim, m, efc, k, ef, seed, allow_frac=0.0, del_frac=0.0):
    """float16 (Euclidean) / int8 (Cosine) index: float32 inputs, the stored rows the oracle derives from them
    (float16.Fromfloat32 / Quantizer.Train + Quantize), the graph built with that precision's distances, and
    the expected SearchWithScores output.  int8 is integer arithmetic: the fixture holds in ANY summation order."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    X[rng.integers(0, n, n // 20)] = X[rng.integers(0, n, n // 20)]
    Q = np.concatenate([rng.standard_normal((24, dim)).astype(np.float32), X[rng.integers(0, n, 8)]])
    metric = O.METRIC_L2 if prec == O.PREC_F16 else O.METRIC_COSINE
    idx = O.OracleIndex(dim, metric, m, efc, O.ARITH_KERNEL, n, precision=prec)
    abs_max = 0.0
    if prec == O.PREC_I8:
        abs_max = float(O.train_quantizer(X))
        idx.set_quantizer(abs_max)
    idx.build_batched(X, rng.random(n), batch=256, threads=4)
    for d in rng.choice(np.arange(1, n + 1), int(n * del_frac), replace=False):
        idx.delete(int(d))
    g = idx.export_graph()
    allow = None
    if allow_frac > 0:
        allow = O.dense_bitset(np.where(rng.random(n + 1) < allow_frac)[0][1:], n)
    ids, sc, cnt, st = idx.search_batch(Q, k, ef, allow=allow, threads=4)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), dim=dim, metric=metric, precision=prec, m=m, k=k, ef=ef, n=g.n,
        inputs=X, abs_max=np.float32(abs_max), rows=idx.rows_raw(),
        norms=idx.norms() if prec == O.PREC_I8 else np.zeros(0, np.float32), levels=g.levels, node_row=g.node_row,
        row_off=g.row_off, nbrs=g.nbrs, deleted=g.deleted, entry=g.entry, max_level=g.max_level, queries=Q,
        allow=np.zeros(0, np.uint64) if allow is None else allow, ids=ids, scores=sc, counts=cnt,
        dist_evals=st.dist_evals, hops=st.hops)
    print(name, "rows", g.n, "evals", st.dist_evals, "hops", st.hops)


if __name__ == "__main__":
    make_quantized("int8_cosine_d40_m8", O.PREC_I8, 900, 40, 8, 50, 10, 40, 303, allow_frac=0.5, del_frac=0.05)
    make_quantized("f16_l2_d36_m6", O.PREC_F16, 900, 36, 6, 40, 5, 32, 404)
    make("cosine_d48_m8", 1500, 48, O.METRIC_COSINE, 8, 60, 10, 48, 101)
    make("l2_d20_m6_filtered", 1200, 20, O.METRIC_L2, 6, 40, 5, 32, 202, allow_frac=0.3, del_frac=0.1)
