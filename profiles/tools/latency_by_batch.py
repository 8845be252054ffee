"""Latency of ONE blocking call through the host-buffer entry point (kdbgpu_search_batch: H2D, preparation, traversal,
D2H, one stream synchronisation) by batch size, on the benchmark workload (1 M x 768 cosine, M=32, efC=200, ef=128,
k=10).  The reference answers one query per call (SearchWithScores): nq = 1 is that call served by the GPU alone,
with no batcher in front.  Prints one JSON line: wall-clock p50 / p99 / mean in ms over CALLS calls per batch size.
  python profiles/tools/latency_by_batch.py [1,8,64,256,1024]        (IDLE=4,8,16: repeat per lone-batch launch shape)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kektordb_b200 import GpuIndex  # noqa: E402

N = int(os.environ.get("N", 1_000_000))
CALLS = int(os.environ.get("CALLS", 200))
D, M, EFC, EF, K = 768, 32, 200, 128, 10
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,8,64,256,1024").split(",")]
dev = torch.device("cuda", 0)
X = bench.make_data(torch, N, D, 32, 0.1, 42, dev)
gi = GpuIndex(D, "cosine", M, N, device=0)
u = np.random.default_rng(1).random(N)
pos = 0
for b in bench.build_schedule(N, EFC, 16384):
    gi.add_batch_device(X[pos:pos + b].data_ptr(), b, D, u[pos:pos + b], EFC)
    pos += b
torch.cuda.synchronize()
del X
Q = bench.make_data(torch, 4096, D, 32, 0.1, 4242, dev).cpu().numpy()
out = {"workload": f"{N}x{D} cosine HNSW M={M} efC={EFC} efSearch={EF} top-{K}; one blocking kdbgpu_search_batch call at a time, pageable host buffers",
       "calls_per_size": CALLS, "by_batch": []}
shapes = [int(x) for x in os.environ.get("IDLE", "0").split(",")]  # IDLE=4,8,16: the lone-batch launch shape (row slots)
for idle, nq in [(a, b) for a in shapes for b in sizes]:
    if idle:
        gi.set_idle_slots(idle)
    for i in range(8):
        gi.SearchWithScores(Q[i * nq % 2048:i * nq % 2048 + nq], K, None, EF)
    t = []
    for i in range(CALLS):
        o = (i * nq) % (4096 - nq + 1)
        t0 = time.perf_counter()
        gi.SearchWithScores(Q[o:o + nq], K, None, EF)
        t.append((time.perf_counter() - t0) * 1e3)
    t = np.array(t)
    out["by_batch"].append({"idle_slots": idle or "default", "nq": nq, "p50_ms": round(float(np.percentile(t, 50)), 3), "p99_ms": round(float(np.percentile(t, 99)), 3),
                            "mean_ms": round(float(t.mean()), 3), "queries_per_s": round(nq / (t.mean() / 1e3), 1)})
print(json.dumps(out))
gi.close()
