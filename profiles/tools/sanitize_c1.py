"""compute-sanitizer target: the C1 shape (128-d cosine, M=16, efC=200, ef=64, k=10) at a size racecheck can
finish — graph construction on the device (build_search / request / scan / commit / seq_add kernels), then the
traversal (heap pass and sorted-list fast pass), the flat scan, the topology staging kernel, a two-shard group and
(TC=1) the tensor-core pre-filter of the flat scan in both kernels; checked against the oracle.
  compute-sanitizer --tool memcheck  python profiles/tools/sanitize_c1.py
  compute-sanitizer --tool racecheck python profiles/tools/sanitize_c1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from kektordb_b200 import GpuIndex, ShardGroup  # noqa: E402
from oracle import oracle as O  # noqa: E402

n = int(os.environ.get("N", 2400))
dim, m, efc, ef, k = 128, 16, 200, 64, 10
rng = np.random.default_rng(0)
X = rng.standard_normal((n, dim)).astype(np.float32)
Q = rng.standard_normal((96, dim)).astype(np.float32)
u = rng.random(n)
gi = GpuIndex(dim, "cosine", m, n)
oi = O.OracleIndex(dim, O.METRIC_COSINE, m, efc, O.ARITH_KERNEL, n)
for a, b in ((0, 200), (200, 600), (600, 1400), (1400, n)):
    gi.AddBatch(X[a:b], u[a:b], efc)
    oi.add_batch(X[a:b], u[a:b], efc, threads=8)
g = oi.export_graph()
gn, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
assert (gn, entry, max_level) == (g.n, g.entry, g.max_level) and np.array_equal(nbrs, g.nbrs), "build differs from the oracle"
want = oi.search_batch(Q, k, ef, threads=8)
for fast in (0, 2):
    gi.set_fast_path(fast)
    got = gi.SearchWithScores(Q, k, None, ef)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), f"search differs (fast={fast})"
allow = O.dense_bitset(np.where(rng.random(n + 1) < 0.2)[0][1:], n)
got = gi.SearchWithScores(Q, k, allow, ef)
wa = oi.search_batch(Q, k, ef, allow=allow, threads=8)
assert np.array_equal(got[0], wa[0]) and np.array_equal(got[1], wa[1]), "filtered search differs"
fi = gi.flat_search(Q[:16], k, 1)
fw = oi.flat_search_batch(Q[:16], k, mode=1, threads=8)
assert np.array_equal(fi[0], fw[0]) and np.array_equal(fi[1], fw[1]), "flat scan differs"
# the topology staged from host arrays (graph_scatter_kernel, several slices): same mirror, same answers
os.environ["KDBGPU_GRAPH_SLICE_NODES"] = "500"
g2 = GpuIndex(dim, "cosine", m, n)
g2.upload_vectors(1, oi.vectors()[1:])
g2.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
del os.environ["KDBGPU_GRAPH_SLICE_NODES"]
assert np.array_equal(g2.get_graph()[4], g.nbrs), "staged topology differs"
got = g2.SearchWithScores(Q, k, None, ef)
assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), "search over the staged topology differs"
if os.environ.get("TC", "0") == "1":  # tensor-core pre-filter of the flat scan: single-CTA kernel (64 queries), CTA pairs (256)
    Q2 = rng.standard_normal((256, dim)).astype(np.float32)
    for qq in (Q[:64], Q2):
        a = g2.flat_search(qq[:16], k, 1)
        b = g2.flat_search(qq, k, 1, prefilter=True)
        assert np.array_equal(a[0], b[0][:16]) and np.array_equal(a[1], b[1][:16]), "pre-filtered flat scan differs"
g2.close()
grp = ShardGroup.local([gi, gi], [0, n])  # the same shard twice: exercises the exchange + merge kernels
ids, sc, cnt, st = grp.SearchWithScores(Q, k, None, ef)
assert st.n_shards == 2 and (cnt == k).all()
grp.close()
gi.close()
print(f"sanitize_c1 ok: built {n} nodes on the device, {len(Q)} queries, results equal to the oracle")
