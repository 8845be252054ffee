import sys, time, os, numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
from kektordb_b200 import GpuIndex
N = int(os.environ.get("N", 100000)); D = 768; M = 32; R = 32
rng = np.random.default_rng(42)
W = np.random.default_rng(777).standard_normal((R, D)).astype(np.float32) / np.sqrt(R)
def gen(n, seed):
    g = np.random.default_rng(seed)
    return g.standard_normal((n, R)).astype(np.float32) @ W + 0.1 * g.standard_normal((n, D)).astype(np.float32)
X = gen(N, 42); Q = gen(4096, 4242)
oi = O.OracleIndex(D, O.METRIC_COSINE, M, 200, O.ARITH_KERNEL, N)
t = time.time(); oi.build_batched(X, np.random.default_rng(1).random(N), batch=4096, ef_const=64, threads=16); print("oracle build s", time.time() - t, flush=True)
g = oi.export_graph()
gi = GpuIndex(D, "cosine", M, N)
gi.upload_vectors(1, oi.vectors()[1:]); gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
gt, _, _, fst = gi.flat_search(Q[:256], 10, 1); print("flat ms", fst.total_ms)
for shape in ((4,2,512),(4,4,512),(8,2,512),(8,1,512),(2,4,512),(4,2,128),(4,1,512)):
    gi.set_tuning(*shape, 0)
    conc = gi.search_concurrency(10, 128)
    best = None
    for it in range(3):
        ids, sc, cnt, st = gi.SearchWithScores(Q[:1024], 10, None, 128)
        if best is None or st.kernel_ms < best.kernel_ms: best = st
    byt = best.dist_evals * D * 4 + best.hops_l0 * 2*M*4 + (best.hops - best.hops_l0) * M * 4
    print(f"shape {shape} conc {conc}: kernel {best.kernel_ms:.3f} ms total {best.total_ms:.3f} ms  QPS(kernel) {1024/best.kernel_ms*1e3:.0f}  E/q {best.dist_evals/1024:.0f} H/q {best.hops/1024:.0f}  GB/s {byt/best.kernel_ms/1e6:.0f} frac {byt/best.kernel_ms/1e6/6550.4:.3f}", flush=True)
rec = np.mean([len(set(ids[i]) & set(gt[i]))/10 for i in range(256)]); print("recall@10", rec)
gi.set_tuning(4,2,512,0)
for nq in (4096,):
    ids, sc, cnt, st = gi.SearchWithScores(Q[:nq], 10, None, 128)
    ids, sc, cnt, st = gi.SearchWithScores(Q[:nq], 10, None, 128)
    byt = st.dist_evals * D * 4 + st.hops_l0 * 2*M*4
    print(f"nq {nq}: kernel {st.kernel_ms:.3f} ms QPS {nq/st.kernel_ms*1e3:.0f} GB/s {byt/st.kernel_ms/1e6:.0f}")
t=time.time(); oids, osc, ocnt, ost = oi.search_batch(Q[:1024], 10, 128, threads=16); dt=time.time()-t
print("oracle 16 threads QPS", 1024/dt, "parity", (oids==ids[:1024]).all(), (osc==sc[:1024]).all())
