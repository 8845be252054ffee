"""A/B of flat-scan launch switches in ONE process on one box (the switches are read per launch):
    python profiles/tools/flat_ab.py  [N] [steps]
Variants: KDBGPU_FLAT_DYNAMIC=0 (static tile schedule of the pair kernel) vs default (tile counter + ring), interleaved
A B A B ... so clock drift under the power cap hits both alike.  Prints the device time of the two tensor passes and of
the whole call per 1024-query step (CUDA events inside the library), and checks that both return the same bits."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import bench_extra  # noqa: E402
from kektordb_b200 import GpuIndex  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 24
D, B, K = 768, 1024, 100
dev = torch.device("cuda", 0)
X = bench.make_data(torch, N, D, 32, 0.1, 42, dev)
gi = GpuIndex(D, "euclidean", 8, N, device=0)
ffi = bench_extra.bench_ffi()
ffi.check(ffi.lib().kdbgpu_upload_vectors_device(gi._h, 1, N, X.data_ptr(), D))
bench_extra._rows_only_graph(gi, N)
Q = bench.make_data(torch, (STEPS + 3) * B, D, 32, 0.1, 4242, dev).cpu().numpy()
variants = {"static": {"KDBGPU_FLAT_DYNAMIC": "0"}, "dynamic": {}}
if os.environ.get("AB") == "rescore":   # re-score row chunk (CTAs per SM) and the sampling stride of the threshold pass
    variants = {"chunk64": {"KDBGPU_FLAT_RS_CHUNK": "64"}, "chunk32": {"KDBGPU_FLAT_RS_CHUNK": "32"},
                "chunk16": {"KDBGPU_FLAT_RS_CHUNK": "16"},
                "chunk32+sample12": {"KDBGPU_FLAT_RS_CHUNK": "32", "KDBGPU_FLAT_SAMPLE": "12"},
                "chunk32+sample16": {"KDBGPU_FLAT_RS_CHUNK": "32", "KDBGPU_FLAT_SAMPLE": "16"}}
keys = {k for v in variants.values() for k in v}
acc = {name: [0.0, 0.0, 0] for name in variants}
ref = {}
for i in range(STEPS + 3):
    for name, env in variants.items():
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        ids, sc, cnt, st = gi.flat_search(Q[i * B:(i + 1) * B], K, 0, prefilter=True)
        if i == 0:
            ref[name] = (ids.copy(), sc.copy(), cnt.copy())
        if i >= 3:
            acc[name][0] += st.hops_l0 / 1e6
            acc[name][1] += st.kernel_ms
            acc[name][2] += 1
names = list(variants)
same = all(np.array_equal(ref[names[0]][j], ref[n][j]) for n in names[1:] for j in range(3))
for name, (t, c, n) in acc.items():
    flops = 2.0 * B * ((N + 255) // 256 * 256) * D * (1 + 1 / int(variants[name].get("KDBGPU_FLAT_SAMPLE", 12)))
    print(f"{name:18s} tensor passes {t / n:.4f} ms/step = {flops / (t / n * 1e-3) / 1e12:7.1f} TFLOP/s   whole call {c / n:.4f} ms/step "
          f"= {B / (c / n * 1e-3) / 1e3:6.1f} k queries/s   ({n} steps)")
print("bit-identical across variants:", same)
gi.close()
