import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import oracle as O
from kektordb_b200 import GpuIndex, dense_allow_list
def recall(ids, gt): return np.mean([len(set(ids[i]) & set(gt[i]))/gt.shape[1] for i in range(len(gt))])
for (N, D, metric, M, ef, k) in ((3000, 128, "cosine", 16, 64, 10), (3000, 100, "euclidean", 8, 40, 10), (2000, 768, "cosine", 32, 128, 10)):
    X = np.random.default_rng(42).standard_normal((N, D)).astype(np.float32)
    Q = np.random.default_rng(4242).standard_normal((64, D)).astype(np.float32)
    om = O.METRIC_COSINE if metric == "cosine" else O.METRIC_L2
    oi = O.OracleIndex(D, om, M, 100, O.ARITH_KERNEL, N)
    oi.build_batched(X, np.random.default_rng(1).random(N), batch=500, threads=8)
    g = oi.export_graph()
    gi = GpuIndex(D, metric, M, N)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    # distance hook
    qn = O.normalize(Q[0]) if metric == "cosine" else Q[0]
    ids = np.arange(1, 200, dtype=np.uint32)
    dg = gi.distance_batch(qn, ids)
    do = np.array([O.distance(om, O.ARITH_KERNEL, qn, oi.vectors()[i]) for i in ids])
    print("distance_batch bit-exact:", (dg == do).all(), np.abs(dg-do).max())
    t=time.time(); gids, gsc, gcnt, st = gi.SearchWithScores(Q, k, None, ef); dt=time.time()-t
    oids, osc, ocnt, ost = oi.search_batch(Q, k, ef, threads=8)
    print(f"N={N} D={D} {metric}: ids equal {(gids==oids).all()} scores equal {(gsc==osc).all()} counts {(gcnt==ocnt).all()} E {st.dist_evals}=={ost.dist_evals} H {st.hops}=={ost.hops} kernel_ms {st.kernel_ms:.3f} total {st.total_ms:.3f}")
    if not (gids==oids).all():
        bad = np.where((gids!=oids).any(axis=1))[0]; print("bad queries", bad[:10]); print(gids[bad[0]], oids[bad[0]]); print(gsc[bad[0]], osc[bad[0]])
    # allow-list 10%
    allow_ids = np.where(np.random.default_rng(7).random(N+1) < 0.1)[0]; allow_ids = allow_ids[allow_ids>0]
    al = dense_allow_list(allow_ids, N)
    gids, gsc, gcnt, st = gi.SearchWithScores(Q, k, al, ef)
    oids, osc, ocnt, ost = oi.search_batch(Q, k, ef, allow=al, threads=8)
    print("  allow-list: ids equal", (gids==oids).all(), "scores", (gsc==osc).all(), "counts", (gcnt==ocnt).all(), "mean count", gcnt.mean())
    # deleted
    dele = np.random.default_rng(9).choice(np.arange(1,N+1), N//10, replace=False)
    for d in dele: oi.delete(int(d))
    gi.set_deleted(dense_allow_list(dele, N))
    gids, gsc, gcnt, st = gi.SearchWithScores(Q, k, None, ef)
    oids, osc, ocnt, ost = oi.search_batch(Q, k, ef, threads=8)
    print("  deleted: ids equal", (gids==oids).all(), "scores", (gsc==osc).all())
    # flat
    fi, fs, fc, _ = gi.flat_search(Q, 20, mode=1)
    oi_, os_, oc_ = oi.flat_search_batch(Q, 20, mode=1, threads=8)
    print("  flat mode1: ids", (fi==oi_).all(), "scores", (fs==os_).all(), np.abs(fs-os_).max())
    fi, fs, fc, _ = gi.flat_search(Q, 20, mode=0)
    oi_, os_, oc_ = oi.flat_search_batch(Q, 20, mode=0, threads=8)
    print("  flat mode0: ids", (fi==oi_).all(), "scores", (fs==os_).all(), np.abs(fs-os_).max())
    gi.close()
