"""int8 / float16 traversal: throughput vs batches in flight, slots, batch size (one build, many configs)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import bench
from kektordb_b200 import GpuIndex
import bench_extra
N = int(os.environ.get("N", 1000000)); D = 768; M = 32
PREC = os.environ.get("PREC", "int8")
metric = "cosine" if PREC == "int8" else "euclidean"
dev = torch.device("cuda", 0)
X = bench.make_data(torch, N, D, 32, 0.1, 42, dev)
gf, bs = bench.build_index(torch, GpuIndex, X, M, 200, 16384, 1, 0, metric)
print("build s", bs, flush=True)
del X
Qd = bench.make_data(torch, 64 * 1024, D, 32, 0.1, 4242, dev)
V = bench_extra._download_rows(gf, N); graph = gf.get_graph(); gf.close()
gi = GpuIndex(D, metric, M, N, precision=PREC)
if PREC == "int8": gi.TrainQuantizer(V[1:])
for i in range(1, N + 1, 1 << 17): gi.upload_vectors(i, V[i:i + (1 << 17)])
gi.set_graph(*graph); del V
k, ef = 10, 128
def run(B, n_ov, steps=24):
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_ov)]
    d_ids = [torch.zeros((B, k), dtype=torch.int32, device=dev) for _ in range(n_ov)]
    d_sc = [torch.zeros((B, k), dtype=torch.float64, device=dev) for _ in range(n_ov)]
    d_cnt = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(n_ov)]
    nb = Qd.shape[0] // B
    def step(i):
        j = i % n_ov; q = Qd[(i % nb) * B:((i % nb) + 1) * B]
        gi.search_device(q.data_ptr(), B, k, ef, d_ids[j].data_ptr(), d_sc[j].data_ptr(), d_cnt[j].data_ptr(), streams[j].cuda_stream)
    for i in range(4): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]: s.wait_event(e0)
    for i in range(steps): step(i)
    for s in streams[1:]: streams[0].wait_stream(s)
    e1.record(streams[0]); torch.cuda.synchronize()
    return B * steps / (e0.elapsed_time(e1) / 1e3)
if os.environ.get("ONE"):
    print("one:", run(1024, 1, steps=6), flush=True)
    sys.exit(0)
for slots, cand in ((4, 192), (8, 192), (2, 192), (4, 64), (8, 64), (16, 64)):
    try:
        gi.set_tuning(slots, cand, 0)
    except Exception as ex:
        print("tuning", slots, cand, ex); continue
    conc = gi.search_concurrency(k, ef)
    for B, n_ov in ((1024, 1), (1024, 3), (1024, 4), (2048, 2), (4096, 1), (4096, 2)):
        print(f"{PREC} slots {slots} cand_smem {cand} conc/SM {conc} batch {B} in-flight {n_ov}: {run(B, n_ov):.0f} QPS", flush=True)
