"""Config 3 probe: 1M x 768 L2 flat top-100, 1024 queries, tensor-core pre-filter vs exhaustive scan."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from kektordb_b200 import GpuIndex
N = int(os.environ.get("N", 1000000)); D = int(os.environ.get("D", 768)); K = int(os.environ.get("K", 100))
NQ = int(os.environ.get("NQ", 1024)); MODEL = os.environ.get("MODEL", "iid")
def gen(n, seed):
    g = np.random.default_rng(seed)
    if MODEL == "iid":
        return g.standard_normal((n, D), dtype=np.float32)
    W = np.random.default_rng(777).standard_normal((32, D)).astype(np.float32) / np.sqrt(32)
    return g.standard_normal((n, 32), dtype=np.float32) @ W + 0.1 * g.standard_normal((n, D), dtype=np.float32)
t = time.time(); X = gen(N, 42); Q = gen(NQ, 4242); print("gen s", time.time() - t, flush=True)
gi = GpuIndex(D, "euclidean", 8, N)
gi.upload_vectors(1, X)
lv = np.zeros(N + 1, np.int32); lv[0] = -1
gi.set_graph(N, lv, np.concatenate([[0], np.arange(N + 1)]).astype(np.uint64), np.zeros(N + 1, np.uint64), np.zeros(1, np.uint32), 1, 0)
for it in range(4):
    ids, sc, cnt, st = gi.flat_search(Q, K, 0, prefilter=True)
    fl = 2.0 * 2 * NQ * N * D
    print(f"prefilter it{it}: total {st.total_ms:.3f} ms gemm(2 passes) {st.kernel_ms:.3f} ms -> {fl/st.kernel_ms/1e9:.1f} TFLOP/s; "
          f"QPS {NQ/st.total_ms*1e3:.0f}; rescored/q {st.dist_evals/NQ:.0f}; fallbacks {st.hops}", flush=True)
nchk = int(os.environ.get("NCHK", 64))
a = gi.flat_search(Q[:nchk], K, 0)
print(f"exhaustive scan {nchk} queries: {a[3].total_ms:.1f} ms -> QPS {nchk/a[3].total_ms*1e3:.0f}")
print("bit-identical:", np.array_equal(a[0], ids[:nchk]), np.array_equal(a[1], sc[:nchk]), np.array_equal(a[2], cnt[:nchk]))
