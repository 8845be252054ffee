"""The topology sidecar file (include/kektordb_gpu.h "graph file", csrc/graphfile.cpp): written and probed
without a device; staged on the GPU it gives the same mirror as kdbgpu_set_graph with the same arrays.
Stands in for the graph part of the reference's gob-encoded .kdb snapshot (pkg/core/core.go:177-306,
hnsw_index.go:3064-3150), which only Go can decode."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from oracle import oracle as O


def _small_graph(n=600, dim=16, m=6, seed=3, metric=O.METRIC_COSINE):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, metric, m, 40, O.ARITH_KERNEL, n)
    oi.build_batched(X, rng.random(n), batch=200, threads=4)
    return oi, oi.export_graph(), rng


def _write(path, g, m):
    from kektordb_b200 import ffi
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lv = np.ascontiguousarray(g.levels, np.int32)
    nr = np.ascontiguousarray(g.node_row, np.uint64)
    ro = np.ascontiguousarray(g.row_off, np.uint64)
    nb = np.ascontiguousarray(g.nbrs, np.uint32)
    ffi.check(ffi.lib().kdbgpu_graph_file_write(path.encode(), g.n, m, p(lv), p(nr), p(ro), p(nb), g.entry, g.max_level))


def _probe(path):
    from kektordb_b200 import ffi
    n, m, rows, edges = C.c_uint32(), C.c_int(), C.c_uint64(), C.c_uint64()
    entry, ml = C.c_uint32(), C.c_int()
    rc = ffi.lib().kdbgpu_graph_file_probe(path.encode(), C.byref(n), C.byref(m), C.byref(rows), C.byref(edges),
                                           C.byref(entry), C.byref(ml))
    return rc, (n.value, m.value, rows.value, edges.value, entry.value, ml.value)


def test_write_probe_and_layout_without_a_device(tmp_path):
    oi, g, _ = _small_graph()
    path = str(tmp_path / "graph.kdbg")
    _write(path, g, 6)
    rc, hdr = _probe(path)
    assert rc == 0 and hdr == (g.n, 6, len(g.row_off) - 1, len(g.nbrs), g.entry, g.max_level)
    raw = open(path, "rb").read()
    magic, ver, n, entry, ml, m, n_rows, n_edges = struct.unpack_from("<IIIIiIQQ", raw, 0)
    assert (magic, ver, n, entry, ml, m) == (0x4742444B, 1, g.n, g.entry, g.max_level, 6)
    off = 64
    assert np.array_equal(np.frombuffer(raw, np.int32, g.n + 1, off), g.levels)
    off += (4 * (g.n + 1) + 7) // 8 * 8
    assert np.array_equal(np.frombuffer(raw, np.uint64, g.n + 2, off), g.node_row)
    off += 8 * (g.n + 2)
    assert np.array_equal(np.frombuffer(raw, np.uint64, n_rows + 1, off), g.row_off)
    off += 8 * (n_rows + 1)
    assert np.array_equal(np.frombuffer(raw, np.uint32, n_edges, off), g.nbrs)
    assert not os.path.exists(path + ".tmp")                  # written atomically
    # damaged files are refused with a message
    from kektordb_b200 import ffi
    bad = str(tmp_path / "bad.kdbg")
    open(bad, "wb").write(b"\0" * 64 + raw[64:])
    assert _probe(bad)[0] == ffi.ERR_INVALID and b"magic" in ffi.lib().kdbgpu_last_error()
    open(bad, "wb").write(raw[:len(raw) // 2])
    assert _probe(bad)[0] == ffi.ERR_INVALID and b"truncated" in ffi.lib().kdbgpu_last_error()
    assert _probe(str(tmp_path / "missing.kdbg"))[0] == ffi.ERR_INVALID


@pytest.mark.gpu
def test_set_graph_file_gives_the_same_mirror_as_set_graph(tmp_path):
    from kektordb_b200 import GpuIndex, ffi
    assert ffi.lib().kdbgpu_device_count() > 0
    oi, g, rng = _small_graph(n=3000, dim=32, m=8, seed=9)
    path = str(tmp_path / "graph.kdbg")
    _write(path, g, 8)
    Q = rng.standard_normal((64, 32)).astype(np.float32)
    want = oi.search_batch(Q, 10, 64, threads=4)
    gi = GpuIndex(32, "cosine", 8, g.n)
    gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph_file(path)
    got = gi.SearchWithScores(Q, 10, None, 64)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[3].dist_evals == want[3].dist_evals
    # the mirror writes the same file back (kdbgpu_save_graph_file), e.g. after a build on the device
    path2 = str(tmp_path / "graph2.kdbg")
    gi.save_graph_file(path2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    # a file written for another M is refused
    gj = GpuIndex(32, "cosine", 16, g.n)
    with pytest.raises(ffi.GpuError):
        gj.set_graph_file(path)
    gi.close()
    gj.close()
