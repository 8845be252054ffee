"""Cold start of a GPU mirror from the reference's on-disk state:
    python profiles/tools/cold_start.py [N] [precision: float32|float16|int8] [dir]
 1. writes <dir>/arena_%04d.bin in the reference's vector-arena format (pkg/storage/mmap/arena.go:14-19, :307-376;
    sequential slots, 64 MiB chunks) and a graph sidecar (include/kektordb_gpu.h: kdbgpu_graph_file_write) for a
    synthetic topology of the benchmark's shape (M = 32: 64 level-0 neighbours per node),
 2. times kdbgpu_arena_load_dir (chunk files mapped, registered with cudaHostRegister, scattered into row order on the
    device) and kdbgpu_set_graph_file on a fresh handle — once with the files in the page cache, and once after
    dropping the page cache when /proc/sys/vm/drop_caches is writable,
 3. checks a sample of rows read back from the device against what was written.
Prints ONE JSON line."""
import ctypes as C
import json
import os
import shutil
import struct
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from kektordb_b200 import GpuIndex, ffi  # noqa: E402
from oracle import arena as A  # noqa: E402  (the format restatement: constants only)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
PREC = sys.argv[2] if len(sys.argv) > 2 else "float32"
DIR = sys.argv[3] if len(sys.argv) > 3 else "/tmp/kdb_cold_start"
D, M = 768, 32
prec_id = {"float32": 0, "float16": 1, "int8": 2}[PREC]
dt = {"float32": np.float32, "float16": np.uint16, "int8": np.int8}[PREC]
vsize = D * np.dtype(dt).itemsize
vpc = (A.CHUNK_SIZE - A.HEADER) // vsize
shutil.rmtree(DIR, ignore_errors=True)
os.makedirs(DIR)
rng = np.random.default_rng(7)
n_chunks = (N + vpc - 1) // vpc
sample = {}
t0 = time.perf_counter()
for c in range(n_chunks):  # slot p of chunk c holds internal id c * vpc + p + 1 (sequential allocation)
    cnt = min(vpc, N - c * vpc)
    if PREC == "float32":
        rows = rng.standard_normal((cnt, D), dtype=np.float32)
    elif PREC == "float16":
        rows = rng.standard_normal((cnt, D), dtype=np.float32).astype(np.float16).view(np.uint16)
    else:
        rows = rng.integers(-127, 128, (cnt, D), dtype=np.int8)
    with open(os.path.join(DIR, f"arena_{c:04d}.bin"), "wb") as f:
        f.write(struct.pack("<IIIB", A.MAGIC, A.VERSION, D, prec_id).ljust(A.HEADER, b"\0"))
        f.write(rows.tobytes())
        f.truncate(A.CHUNK_SIZE)
    for p in (0, cnt // 2, cnt - 1):
        sample[c * vpc + p + 1] = rows[p].copy()
write_s = time.perf_counter() - t0
# topology: every node on level 0 with 2M neighbours (what the sidecar of a built M = 32 graph weighs)
lv = np.zeros(N + 1, np.int32)
lv[0] = -1
node_row = np.concatenate([[0], np.arange(N + 1)]).astype(np.uint64)
row_off = (np.arange(N + 1, dtype=np.uint64) * np.uint64(2 * M))
nbrs = rng.integers(1, N + 1, N * 2 * M, dtype=np.uint32)
gpath = os.path.join(DIR, "graph.kdbg")
p = lambda a: a.ctypes.data_as(C.c_void_p)
ffi.check(ffi.lib().kdbgpu_graph_file_write(gpath.encode(), N, M, p(lv), p(node_row), p(row_off), p(nbrs), 1, 0))
arena_bytes = N * vsize
graph_bytes = os.path.getsize(gpath)


def load(tag):
    gi = GpuIndex(D, "cosine" if PREC == "int8" else "euclidean", M, N, precision=PREC)
    t0 = time.perf_counter()
    staged = gi.load_arena(DIR, None, N)
    t1 = time.perf_counter()
    gi.set_graph_file(gpath)
    t2 = time.perf_counter()
    ok = staged == N
    for i, row in sample.items():
        ok = ok and np.array_equal(gi.download_rows_raw(i, 1)[0].view(dt), row.view(dt))
    reg = int(ffi.lib().kdbgpu_arena_chunks_registered(gi._handle()))
    gi.close()
    return {"page_cache": tag, "arena_s": round(t1 - t0, 3), "arena_GBps": round(arena_bytes / (t1 - t0) / 1e9, 2),
            "graph_s": round(t2 - t1, 3), "graph_GBps": round(graph_bytes / (t2 - t1) / 1e9, 2),
            "total_s": round(t2 - t0, 3), "chunks_registered": reg, "rows_match": bool(ok)}


GpuIndex(D, "euclidean", M, 16).close()  # CUDA context creation stays out of the timings
runs = [load("warm"), load("warm, second handle")]
try:
    os.sync()
    with open("/proc/sys/vm/drop_caches", "w") as f:
        f.write("3\n")
    runs.append(load("dropped"))
except OSError as e:
    runs.append({"page_cache": "dropped", "skipped": str(e)})
print(json.dumps({"workload": f"cold start: {N} x {D} {PREC} arena ({n_chunks} chunks of 64 MiB) + graph sidecar (M={M})",
                  "arena_bytes": arena_bytes, "graph_file_bytes": graph_bytes, "write_s": round(write_s, 1), "runs": runs}))
shutil.rmtree(DIR, ignore_errors=True)
