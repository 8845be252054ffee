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


def _random_csr(rng, n, m, nil_frac=0.05, upper_frac=0.06):
    """A synthetic reference-shaped topology: levels (nil slots = -1, a few nodes on upper levels), rows with 0 .. 2M / M
    neighbours that may name nil nodes, id 0 and ids beyond n (all of which the mirror must drop, keeping the order)."""
    levels = np.zeros(n + 1, np.int32)
    levels[0] = -1
    up = rng.random(n + 1) < upper_frac
    levels[up] = rng.integers(1, 4, up.sum())
    levels[rng.random(n + 1) < nil_frac] = -1
    levels[1] = 3
    node_row, row_off, nbrs = [0], [0], []
    for i in range(n + 1):
        for l in range(levels[i] + 1 if levels[i] >= 0 else 0):
            cap = 2 * m if l == 0 else m
            row = rng.integers(0, n + 3, rng.integers(0, cap + 1))
            nbrs.extend(row.tolist())
            row_off.append(len(nbrs))
        node_row.append(len(row_off) - 1)
    return levels, np.array(node_row, np.uint64), np.array(row_off, np.uint64), np.array(nbrs, np.uint32)


def _live_rows(levels, node_row, row_off, nbrs, n):
    out = {}
    for i in range(1, n + 1):
        for l in range(levels[i] + 1 if levels[i] >= 0 else 0):
            r = int(node_row[i]) + l
            row = nbrs[int(row_off[r]):int(row_off[r + 1])]
            out[(i, l)] = [int(x) for x in row if 0 < x <= n and levels[x] >= 0]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("slice_nodes,slice_edges", [(0, 0), (97, 0), (0, 700), (1, 0)])
def test_topology_is_staged_slice_by_slice_on_the_device(slice_nodes, slice_edges, monkeypatch):
    """kdbgpu_set_graph pads the rows and drops nil neighbours on the device, one slice of whole nodes at a time; with
    the slice limits shrunk the same small graph travels as one slice, as dozens, and as one slice per node — the
    mirror read back (kdbgpu_get_graph) must hold exactly the live neighbours of every row, in order."""
    from kektordb_b200 import GpuIndex, ffi
    assert ffi.lib().kdbgpu_device_count() > 0
    for k, v in (("KDBGPU_GRAPH_SLICE_NODES", slice_nodes), ("KDBGPU_GRAPH_SLICE_EDGES", slice_edges)):
        if v:
            monkeypatch.setenv(k, str(v))
        else:
            monkeypatch.delenv(k, raising=False)
    rng = np.random.default_rng(5)
    n, m = 1500, 6
    levels, node_row, row_off, nbrs = _random_csr(rng, n, m)
    gi = GpuIndex(8, "euclidean", m, n + 10)
    gi.upload_vectors(1, rng.standard_normal((n, 8)).astype(np.float32))
    gi.set_graph(n, levels, node_row, row_off, nbrs, 1, 3)
    n2, lv2, nr2, ro2, nb2, entry, ml = gi.get_graph()
    assert (n2, entry, ml) == (n, 1, 3) and np.array_equal(lv2, levels)
    assert _live_rows(lv2, nr2, ro2, nb2, n) == _live_rows(levels, node_row, row_off, nbrs, n)
    assert sum(len(v) for v in _live_rows(lv2, nr2, ro2, nb2, n).values()) == nb2.size   # nothing but live ids is kept
    gi.close()


@pytest.mark.gpu
def test_a_row_with_too_many_live_neighbours_leaves_the_mirror_without_a_graph():
    from kektordb_b200 import GpuIndex, ffi
    rng = np.random.default_rng(6)
    n, m = 300, 4
    levels, node_row, row_off, nbrs = _random_csr(rng, n, m, nil_frac=0.0)
    gi = GpuIndex(8, "euclidean", m, n)
    gi.upload_vectors(1, rng.standard_normal((n, 8)).astype(np.float32))
    gi.set_graph(n, levels, node_row, row_off, nbrs, 1, 3)
    Q = rng.standard_normal((4, 8)).astype(np.float32)
    gi.SearchWithScores(Q, 3, None, 16)
    # node 7, level 0: 2M + 1 live neighbours
    bad_row = np.arange(10, 10 + 2 * m + 1, dtype=np.uint32)
    r = int(node_row[7])
    b, e = int(row_off[r]), int(row_off[r + 1])
    nbrs_bad = np.concatenate([nbrs[:b], bad_row, nbrs[e:]])
    row_off_bad = row_off.copy()
    row_off_bad[r + 1:] += np.uint64(bad_row.size - (e - b))     # the row held at most 2M entries before
    with pytest.raises(ffi.GpuError, match="node 7 level 0 has more than 8 neighbours"):
        gi.set_graph(n, levels, node_row, row_off_bad, nbrs_bad, 1, 3)
    with pytest.raises(ffi.GpuError):                      # no graph until a set_graph succeeds
        gi.SearchWithScores(Q, 3, None, 16)
    lv = levels.copy()
    lv[9] = 121                                            # per-node errors are found before the mirror is touched
    gi.set_graph(n, levels, node_row, row_off, nbrs, 1, 3)
    with pytest.raises(ffi.GpuError, match="node 9"):
        gi.set_graph(n, lv, node_row, row_off, nbrs, 1, 3)
    assert gi.SearchWithScores(Q, 3, None, 16)[2].min() >= 1
    gi.close()
