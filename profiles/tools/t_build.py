import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
from kektordb_b200 import GpuIndex
def compare(N, D, metric, M, efc, batches, seed, data="normal"):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32) if data=="normal" else rng.integers(-2,3,(N,D)).astype(np.float32)
    u = rng.random(N)
    om = O.METRIC_COSINE if metric == "cosine" else O.METRIC_L2
    oi = O.OracleIndex(D, om, M, efc, O.ARITH_KERNEL, N)
    gi = GpuIndex(D, metric, M, N)
    pos = 0; tg = to = 0
    for b in batches:
        b = min(b, N - pos)
        if b <= 0: break
        t=time.time(); gi.AddBatch(X[pos:pos+b], u[pos:pos+b], efc); tg += time.time()-t
        t=time.time(); oi.add_batch(X[pos:pos+b], u[pos:pos+b], efc, threads=16); to += time.time()-t
        pos += b
        g = oi.export_graph()
        n, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
        ok = (n == g.n and entry == g.entry and max_level == g.max_level and np.array_equal(levels, g.levels)
              and np.array_equal(row_off, g.row_off) and np.array_equal(nbrs, g.nbrs))
        if not ok:
            print(f"  MISMATCH after batch ending at {pos}: n {n}/{g.n} entry {entry}/{g.entry} maxl {max_level}/{g.max_level} levels {np.array_equal(levels, g.levels)} row_off {np.array_equal(row_off, g.row_off)}")
            if np.array_equal(levels, g.levels):
                for i in range(1, n+1):
                    for l in range(levels[i]+1):
                        r = int(node_row[i])+l
                        a = nbrs[int(row_off[r]):int(row_off[r+1])]; bb = g.row(i,l)
                        if not np.array_equal(a, bb):
                            print("   first diff node", i, "level", l, "gpu", a[:12], "len", len(a), "oracle", bb[:12], "len", len(bb)); break
                    else: continue
                    break
            return False
    vec_ok = np.array_equal(gi.download_vectors(1, N), oi.vectors()[1:])
    print(f"N={N} D={D} {metric} M={M} efc={efc} {data}: graphs identical after every batch; vectors identical {vec_ok}; gpu {tg:.2f}s oracle(16t) {to:.2f}s")
    gi.close(); return True
compare(300, 16, "euclidean", 4, 20, [10, 5, 5, 30, 50, 200], 1)
compare(1500, 32, "cosine", 8, 40, [40, 60, 100, 300, 1000], 2)
compare(1200, 12, "euclidean", 6, 30, [30, 70, 300, 800], 3, data="grid")
compare(6000, 128, "cosine", 16, 200, [200, 300, 500, 1000, 4000], 4)
compare(5000, 768, "cosine", 32, 200, [256, 256, 512, 1024, 2952], 5)
