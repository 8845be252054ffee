"""The arena file format restatement (oracle/arena.py) against the constants and layout rules of
pkg/storage/mmap/arena.go; CPU only.  The GPU loader is checked against these files in
tests/test_gpu_arena.py."""
import os
import struct

import numpy as np
import pytest

from oracle import arena as A


def test_constants_and_capacity_rule():
    assert (A.CHUNK_SIZE, A.MAGIC, A.VERSION, A.HEADER) == (64 * 1024 * 1024, 0x4B414F4E, 1, 64)   # arena.go:14-19
    assert A.MAGIC.to_bytes(4, "big") == b"KAON"
    assert A.vecs_per_chunk(768, 0) == (64 * 1024 * 1024 - 64) // 3072 == 21845                     # :93-95
    assert A.vecs_per_chunk(768, 1) == 43690 and A.vecs_per_chunk(768, 2) == 87381


@pytest.mark.parametrize("precision,dim", [(0, 5), (1, 7), (2, 9), (0, 128)])
def test_write_read_round_trip_with_permuted_slots(tmp_path, precision, dim):
    rng = np.random.default_rng(dim)
    n = 300
    dt = A.PREC_DTYPE[precision]
    rows = (rng.standard_normal((n + 1, dim)) * 50).astype(dt)
    st = np.full(n + 1, A.UNALLOCATED, np.uint32)
    st[1:] = rng.permutation(n).astype(np.uint32)          # AddBatch allocates slots from racing goroutines
    st[rng.integers(1, n + 1, 20)] = A.UNALLOCATED         # FreeSlot'ed ids
    d = str(tmp_path / "arena")
    assert A.write_arena(d, rows, precision, st, truncate=True) == 1
    raw = open(os.path.join(d, "arena_0000.bin"), "rb").read()
    assert struct.unpack("<IIIB", raw[:13]) == (A.MAGIC, 1, dim, precision) and raw[13:64] == b"\0" * 51
    got = A.read_arena(d, dim, precision, st)
    ok = st != A.UNALLOCATED
    assert np.array_equal(got[ok], rows[ok]) and not got[~ok].any()
    # slot p sits at byte 64 + p * vectorSize (single chunk)
    i = int(np.where(ok)[0][3])
    vs = dim * np.dtype(dt).itemsize
    off = 64 + int(st[i]) * vs
    assert raw[off:off + vs] == rows[i].tobytes()


def test_header_validation(tmp_path):
    rows = np.ones((3, 4), np.float32)
    st = A.sequential_slot_table(2)
    for bad in ({"magic": 1}, {"version": 2}, {"dim": 5}, {"precision": 1}):
        d = str(tmp_path / ("a_" + next(iter(bad))))
        A.write_arena(d, rows, 0, st, truncate=True, header_override=bad)
        with pytest.raises(ValueError):
            A.read_arena(d, 4, 0, st)


def test_library_probe_reads_the_same_headers_without_a_device(tmp_path):
    """kdbgpu_arena_probe (host-only): dim / precision / chunk count / rows per chunk, and addChunk's header checks."""
    from kektordb_b200 import arena_probe, ffi
    rng = np.random.default_rng(5)
    n, dim = 90000, 768
    rows = rng.integers(-100, 100, (n + 1, dim)).astype(np.int8)
    d = str(tmp_path / "arena")
    n_chunks = A.write_arena(d, rows, 2, A.sequential_slot_table(n), truncate=True)
    assert n_chunks == 2
    assert arena_probe(d) == (dim, "int8", 2, A.vecs_per_chunk(dim, 2))
    for bad, msg in (({"magic": 9}, "magic mismatch"), ({"version": 4}, "unsupported version")):
        b = str(tmp_path / next(iter(bad)))
        A.write_arena(b, rows[:10], 2, A.sequential_slot_table(9), truncate=True, header_override=bad)
        with pytest.raises(ffi.GpuError, match=msg):
            arena_probe(b)
    with pytest.raises(ffi.GpuError):
        arena_probe(str(tmp_path / "missing"))
