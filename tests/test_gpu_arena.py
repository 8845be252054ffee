"""Arena -> HBM staging: kdbgpu_arena_load_dir / kdbgpu_arena_stage_chunk against
arena files laid out as pkg/storage/mmap/arena.go does (oracle/arena.py): permuted and freed slots,
several 64 MiB chunks, all three precisions, header validation; then a search over the staged rows is
bit-identical to one over rows uploaded directly."""
import os

import numpy as np
import pytest

from oracle import arena as A
from oracle import oracle as O

pytestmark = pytest.mark.gpu

PREC_NAME = {0: "float32", 1: "float16", 2: "int8"}


def _gpu():
    from kektordb_b200 import GpuIndex, ffi
    assert ffi.lib().kdbgpu_device_count() > 0
    return GpuIndex


def _rows(rng, n, dim, precision):
    if precision == 0:
        r = rng.standard_normal((n + 1, dim)).astype(np.float32)
    elif precision == 1:
        r = rng.standard_normal((n + 1, dim)).astype(np.float16).view(np.uint16)
    else:
        r = rng.integers(-127, 128, (n + 1, dim)).astype(np.int8)
    r[0] = 0
    return r


@pytest.mark.parametrize("precision,dim,n", [(0, 768, 50000), (1, 768, 60000), (2, 768, 100000), (0, 5, 1000),
                                             (2, 33, 2000), (1, 7, 500)])
def test_load_dir_places_every_row_at_its_logical_id(tmp_path, precision, dim, n):
    GpuIndex = _gpu()
    rng = np.random.default_rng(n)
    rows = _rows(rng, n, dim, precision)
    st = np.full(n + 1, A.UNALLOCATED, np.uint32)
    st[1:] = rng.permutation(n).astype(np.uint32)
    freed = rng.integers(1, n + 1, n // 50)
    st[freed] = A.UNALLOCATED
    d = str(tmp_path / "arena")
    n_chunks = A.write_arena(d, rows, precision, st)
    if dim == 768:
        assert n_chunks >= 2                                   # more than one 64 MiB chunk
    gi = GpuIndex(dim, "cosine" if precision == 2 else "euclidean", 4, n, precision=PREC_NAME[precision])
    staged = gi.load_arena(d, st)
    ok = st != A.UNALLOCATED
    assert staged == int(ok.sum())
    got = gi.download_rows_raw(1, n)
    want = rows.copy()
    want[~ok] = 0
    assert np.array_equal(got, want[1:])
    assert np.array_equal(got, A.read_arena(d, dim, precision, st)[1:])
    if precision == 2:
        nr = gi.download_norms(1, n)
        assert np.array_equal(nr[:200], np.array([O.int8_norm(r) for r in want[1:201]], np.float32))
    gi.close()


def test_sequential_slots_and_stage_chunk_from_memory(tmp_path):
    GpuIndex = _gpu()
    rng = np.random.default_rng(3)
    n, dim = 30000, 768
    rows = _rows(rng, n, dim, 0)
    d = str(tmp_path / "arena")
    n_chunks = A.write_arena(d, rows, 0, A.sequential_slot_table(n))
    gi = GpuIndex(dim, "euclidean", 4, n)
    assert gi.load_arena(d, None, n) == n                      # NULL slot table: id i in slot i - 1
    assert np.array_equal(gi.download_vectors(1, n), rows[1:])
    gi.close()
    g2 = GpuIndex(dim, "euclidean", 4, n)                      # the host already has the chunks mapped
    total = 0
    for c in range(n_chunks):
        chunk = np.fromfile(os.path.join(d, f"arena_{c:04d}.bin"), dtype=np.uint8)
        total += g2.stage_arena_chunk(c, chunk, None, n + 1)
    assert total == n and np.array_equal(g2.download_vectors(1, n), rows[1:])
    g2.close()


def test_header_validation_and_missing_files(tmp_path):
    from kektordb_b200 import ffi
    GpuIndex = _gpu()
    rows = _rows(np.random.default_rng(1), 10, 16, 0)
    st = A.sequential_slot_table(10)
    gi = GpuIndex(16, "euclidean", 4, 10)
    for bad, msg in (({"magic": 7}, "magic mismatch"), ({"version": 3}, "unsupported version"),
                     ({"dim": 17}, "dimension mismatch"), ({"precision": 2}, "precision mismatch")):
        d = str(tmp_path / ("a_" + next(iter(bad))))
        A.write_arena(d, rows, 0, st, truncate=True, header_override=bad)
        with pytest.raises(ffi.GpuError, match=msg):           # addChunk's messages (arena.go:352-363)
            gi.load_arena(d, st)
    with pytest.raises(ffi.GpuError):
        gi.load_arena(str(tmp_path / "nothing_here"), st)
    far = st.copy()
    far[3] = 10_000_000                                        # a slot in a chunk that does not exist
    d = str(tmp_path / "ok")
    A.write_arena(d, rows, 0, st, truncate=True)
    with pytest.raises(ffi.GpuError, match="refers to chunk"):
        gi.load_arena(d, far)
    assert gi.load_arena(d, st) == 10                          # truncated file: short payload reads as zero pages
    assert np.array_equal(gi.download_vectors(1, 10), rows[1:])
    gi.close()


def test_search_over_arena_staged_rows_is_bit_identical(tmp_path):
    GpuIndex = _gpu()
    rng = np.random.default_rng(9)
    n, dim, m = 3000, 96, 8
    X = rng.standard_normal((n, dim)).astype(np.float32)
    oi = O.OracleIndex(dim, O.METRIC_COSINE, m, 60, O.ARITH_KERNEL, n)
    oi.build_batched(X, rng.random(n), batch=500, threads=8)
    g = oi.export_graph()
    st = np.full(n + 1, A.UNALLOCATED, np.uint32)
    st[1:] = rng.permutation(n).astype(np.uint32)
    d = str(tmp_path / "arena")
    A.write_arena(d, oi.vectors(), 0, st, truncate=True)
    gi = GpuIndex(dim, "cosine", m, n)
    assert gi.load_arena(d, st) == n
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    Q = rng.standard_normal((64, dim)).astype(np.float32)
    ids, sc, cnt, stt = gi.SearchWithScores(Q, 10, None, 64)
    oids, osc, ocnt, ost = oi.search_batch(Q, 10, 64, threads=8)
    assert np.array_equal(ids, oids) and np.array_equal(sc, osc) and stt.dist_evals == ost.dist_evals
    gi.close()


def test_stage_chunk_from_the_hosts_own_mmap_page_locks_it_in_place(tmp_path):
    """The Go side keeps every chunk mmap'ed (arena.go:307-376): kdbgpu_arena_stage_chunk reads that mapping
    directly (cudaHostRegister for the duration of the copy) — no bounce through another pinned buffer."""
    import mmap
    from kektordb_b200 import ffi
    GpuIndex = _gpu()
    rng = np.random.default_rng(7)
    n, dim = 30000, 768                                       # 92 MB of float32 rows: two chunks
    rows = _rows(rng, n, dim, 0)
    d = str(tmp_path / "arena")
    n_chunks = A.write_arena(d, rows, 0, A.sequential_slot_table(n))
    assert n_chunks == 2
    gi = GpuIndex(dim, "euclidean", 4, n)
    maps, staged = [], 0
    for c in range(n_chunks):
        f = open(os.path.join(d, "arena_%04d.bin" % c), "r+b")
        # PROT_READ | PROT_WRITE, MAP_SHARED: the mapping the reference holds (pkg/storage/mmap/mmap_unix.go:15)
        mm = mmap.mmap(f.fileno(), 0, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
        maps.append((f, mm))
        staged += gi.stage_arena_chunk(c, np.frombuffer(mm, dtype=np.uint8), None, n + 1)
    assert staged == n
    assert np.array_equal(gi.download_rows_raw(1, n), rows[1:])
    # whether a file-backed mapping can be page-locked is the driver's call (it refuses on some kernels): the counter
    # says what happened, the rows are right either way
    reg = ffi.lib().kdbgpu_arena_chunks_registered(gi._h)
    assert reg in (0, n_chunks)
    print(f"file-backed arena mappings page-locked in place: {reg} of {n_chunks}")
    # a malloc'ed, unaligned copy of the same bytes takes the ordinary copy path and stages the same rows
    gj = GpuIndex(dim, "euclidean", 4, n)
    for c in range(n_chunks):
        buf = np.empty(len(maps[c][1]) + 1, np.uint8)[1:]
        buf[:] = np.frombuffer(maps[c][1], dtype=np.uint8)
        gj.stage_arena_chunk(c, buf, None, n + 1)
    assert np.array_equal(gj.download_rows_raw(1, n), rows[1:])
    assert ffi.lib().kdbgpu_arena_chunks_registered(gj._h) == 0
    gi.close()
    gj.close()
    for f, mm in maps:
        del mm
        f.close()
