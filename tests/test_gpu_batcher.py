"""The micro-batcher over a real GPU index: every caller thread gets exactly what a direct
SearchWithScores for its query alone returns (queries of a batch are independent), bit for bit."""
import threading

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_batcher_results_equal_direct_search():
    from kektordb_b200 import Batcher, GpuIndex, ffi
    assert ffi.lib().kdbgpu_device_count() > 0
    rng = np.random.default_rng(5)
    n, dim, m = 6000, 64, 12
    X = rng.standard_normal((n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, O.METRIC_COSINE, m, 80, O.ARITH_KERNEL, n)
    oi.build_batched(X, rng.random(n), batch=1000, threads=8)
    g = oi.export_graph()
    gi = GpuIndex(dim, "cosine", m, n)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    Q = rng.standard_normal((400, dim)).astype(np.float32)
    members = np.where(rng.random(n + 1) < 0.3)[0]
    allow = O.dense_bitset(members[members > 0], n)
    want = gi.SearchWithScores(Q, 10, None, 64)
    want_f = gi.SearchWithScores(Q, 10, allow, 64)
    b = Batcher(gi, max_batch=128, max_wait_us=2000)
    out = {}

    def call(i):
        out[i] = b.SearchWithScores(Q[i], 10, allow if i % 3 == 0 else None, 64)

    ts = [threading.Thread(target=call, args=(i,)) for i in range(len(Q))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(len(Q)):
        ref = want_f if i % 3 == 0 else want
        c = int(ref[2][i])
        assert np.array_equal(out[i][0], ref[0][i][:c]) and np.array_equal(out[i][1], ref[1][i][:c])
    st = b.stats()
    assert st.queries == len(Q) and st.batches < len(Q)
    # and against the CPU oracle, one query at a time (the reference's own call shape)
    for i in (1, 2, 4):
        ids, sc = oi.search(Q[i], 10, 64)
        assert np.array_equal(out[i][0], ids) and np.array_equal(out[i][1], sc)
    b.close()
    gi.close()
