"""Fast pass vs heap pass: share of queries re-answered after a tie, and kernel time, per precision."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import bench, bench_extra
from kektordb_b200 import GpuIndex
N = int(os.environ.get("N", 1000000)); D = 768; M = 32
PREC = os.environ.get("PREC", "float32")
metric = "euclidean" if PREC == "float16" else "cosine"
dev = torch.device("cuda", 0)
X = bench.make_data(torch, N, D, 32, 0.1, 42, dev)
if PREC != "float32" and metric == "cosine":
    X /= X.norm(dim=1, keepdim=True)
gi = GpuIndex(D, metric, M, N, precision=PREC)
if PREC == "int8": gi.train_quantizer_device(X.data_ptr(), D, N)
u = np.random.default_rng(1).random(N); pos = 0
for b in bench.build_schedule(N, 200, 16384):
    gi.add_batch_device(X[pos:pos + b].data_ptr(), b, D, u[pos:pos + b], 200); pos += b
torch.cuda.synchronize(); del X
Q = bench.make_data(torch, 8 * 1024, D, 32, 0.1, 4242, dev).cpu().numpy()
gi.prepare_search(1024, 10, 128)
ref = None
for mode in (0, 2, 0, 2):
    gi.set_fast_path(mode)
    gi.prepare_search(1024, 10, 128)
    best, redo = 1e9, 0
    for i in range(6):
        ids, sc, cnt, st = gi.SearchWithScores(Q[(i % 8) * 1024:(i % 8 + 1) * 1024], 10, None, 128)
        if i == 0:
            if ref is None: ref = (ids.copy(), sc.copy())
            else: assert np.array_equal(ref[0], ids) and np.array_equal(ref[1], sc), "fast != heap"
        best = min(best, st.kernel_ms); redo += st.heap_pass_queries
    print(f"{PREC} fast={mode}: best kernel {best:.3f} ms ({1024/best*1e3:.0f} QPS isolated), re-answered {redo/6/1024*100:.1f} % of queries", flush=True)
