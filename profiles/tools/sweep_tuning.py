"""Traversal-kernel shape sweep on the benchmark workload (1M x 768 cosine, M=32, ef=128, k=10, 1024 queries):
isolated launch time and pipelined step interval for every (slots, cand_smem, max CTAs/SM) asked for.
  gpurun -- 'python profiles/tools/sweep_tuning.py 4,192,0 8,192,0 8,192,7 16,128,0'
Env: PREC=float32|int8|float16, N, OV (batches in flight, default 3), FAST (kdbgpu_set_fast_path).
A fourth field is the idle shape (kdbgpu_set_idle_slots): 4,192,0,8 = slots 4 with batches in flight, 8 alone."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kektordb_b200 import GpuIndex  # noqa: E402

N = int(os.environ.get("N", 1_000_000))
D, M, EFC, EF, K, B = 768, 32, 200, 128, 10, int(os.environ.get("B", 1024))
prec = os.environ.get("PREC", "float32")
ov = int(os.environ.get("OV", 3))
dev = torch.device("cuda", 0)
metric = "euclidean" if prec == "float16" else "cosine"
X = bench.make_data(torch, N, D, 32, 0.1, 42, dev)
if prec != "float32" and metric == "cosine":
    X /= X.norm(dim=1, keepdim=True).clamp_min(1e-30)
gi = GpuIndex(D, metric, M, N, device=0, precision=prec)
if prec == "int8":
    gi.train_quantizer_device(X.data_ptr(), D, N)
u = np.random.default_rng(1).random(N)
pos = 0
for b in bench.build_schedule(N, EFC, 16384):
    gi.add_batch_device(X[pos:pos + b].data_ptr(), b, D, u[pos:pos + b], EFC)
    pos += b
torch.cuda.synchronize()
del X
if "FAST" in os.environ:
    gi.set_fast_path(int(os.environ["FAST"]))
nb = max(4, 24 * 1024 // B)
Qd = bench.make_data(torch, nb * B, D, 32, 0.1, 4242, dev)
esz = {"float32": 4, "float16": 2, "int8": 1}[prec]
row_bytes = (D * esz + 127) // 128 * 128 + (4 if prec == "int8" else 0)
shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(4, 192, 0)]
for shape in shapes:
    slots, cs, cap = shape[:3]
    gi.set_tuning(slots, cs, cap)
    gi.set_idle_slots(shape[3] if len(shape) > 3 else slots)
    gi.prepare_search(B, K, EF)
    conc = gi.search_concurrency(K, EF)
    for n_ov in (1, ov):
        run = bench.DeviceRunner(torch, dev, gi.search_device, Qd, B, K, EF, n_ov)
        for i in range(4):
            run.step(i)
        torch.cuda.synchronize()
        ms = run.timed(4, 20, torch.cuda.synchronize) / 20
        st = gi.last_search_stats()
        byts = st.dist_evals * row_bytes + st.hops_l0 * 2 * M * 4 + (st.hops - st.hops_l0) * M * 4
        print(f"{prec} slots={slots}{'/idle ' + str(shape[3]) if len(shape) > 3 else ''} cand_smem={cs} cap={cap} resident={conc} in_flight={n_ov}: {ms:.3f} ms/step "
              f"{B / ms * 1e3:,.0f} q/s  hbm_frac={byts / (ms / 1e3) / 1e9 / 6550.4:.3f}", flush=True)
