"""CPU restatement of the reference's vector-arena file format (TEST INFRASTRUCTURE ONLY, like the rest
of oracle/): writes and reads <dir>/arena_%04d.bin exactly as pkg/storage/mmap/arena.go lays them out,
so the GPU loader (kdbgpu_arena_load_dir / kdbgpu_arena_stage_chunk) can be checked against files a
reference process would have produced.

Format (arena.go:14-19 constants, :79-118 NewVectorArena, :307-376 addChunk, :378-444 GetBytes,
:121-152 AllocSlot):
  * chunk files of DefaultChunkSize = 64 MiB, header of ArenaHeaderSize = 64 bytes:
    u32 LE ArenaMagic 0x4B414F4E, u32 ArenaVersion 1, u32 dim, u8 precision (0 f32, 1 f16, 2 int8), zeros;
  * vecsPerChk = (64 MiB - 64) // vectorSize, vectorSize = dim * element bytes;
  * logical internal id -> physical slot via the slot table (0xFFFFFFFF = unallocated);
    slot p is in chunk p // vecsPerChk at byte 64 + (p % vecsPerChk) * vectorSize.
"""
from __future__ import annotations

import os
import struct

import numpy as np

CHUNK_SIZE = 64 * 1024 * 1024
MAGIC = 0x4B414F4E
VERSION = 1
HEADER = 64
UNALLOCATED = 0xFFFFFFFF
PREC_DTYPE = {0: np.float32, 1: np.uint16, 2: np.int8}


def vecs_per_chunk(dim: int, precision: int) -> int:
    return (CHUNK_SIZE - HEADER) // (dim * np.dtype(PREC_DTYPE[precision]).itemsize)


def sequential_slot_table(n: int) -> np.ndarray:
    """AllocSlot over ids 1..n in order: id i gets physical slot i - 1; entry 0 (the nil id) stays unallocated."""
    t = np.full(n + 1, UNALLOCATED, dtype=np.uint32)
    t[1:] = np.arange(n, dtype=np.uint32)
    return t


def write_arena(arena_dir: str, rows: np.ndarray, precision: int, slot_table: np.ndarray, truncate: bool = False,
                header_override: dict | None = None) -> int:
    """rows: [n+1, dim] stored rows indexed by internal id (row 0 unused).  Returns the number of chunks.
    truncate=False writes full 64 MiB files as the reference's Truncate does (sparse on disk)."""
    os.makedirs(arena_dir, exist_ok=True)
    dt = np.dtype(PREC_DTYPE[precision])
    rows = np.ascontiguousarray(rows, dtype=dt)
    dim = rows.shape[1]
    vsize = dim * dt.itemsize
    vpc = (CHUNK_SIZE - HEADER) // vsize
    alloc = np.where(slot_table != UNALLOCATED)[0]
    n_chunks = int(slot_table[alloc].max()) // vpc + 1 if alloc.size else 1
    files = []
    for c in range(n_chunks):
        f = open(os.path.join(arena_dir, f"arena_{c:04d}.bin"), "wb+")
        hd = {"magic": MAGIC, "version": VERSION, "dim": dim, "precision": precision}
        hd.update(header_override or {})
        f.write(struct.pack("<IIIB", hd["magic"], hd["version"], hd["dim"], hd["precision"]).ljust(HEADER, b"\0"))
        if not truncate:
            f.truncate(CHUNK_SIZE)
        files.append(f)
    for i in alloc:
        p = int(slot_table[i])
        f = files[p // vpc]
        f.seek(HEADER + (p % vpc) * vsize)
        f.write(rows[i].tobytes())
    for f in files:
        f.close()
    return n_chunks


def read_arena(arena_dir: str, dim: int, precision: int, slot_table: np.ndarray) -> np.ndarray:
    """VectorArena.GetBytes for every allocated id: [len(slot_table), dim] (unallocated rows are zero)."""
    dt = np.dtype(PREC_DTYPE[precision])
    vsize = dim * dt.itemsize
    vpc = (CHUNK_SIZE - HEADER) // vsize
    out = np.zeros((len(slot_table), dim), dtype=dt)
    cache = {}
    for i in np.where(slot_table != UNALLOCATED)[0]:
        p = int(slot_table[i])
        c = p // vpc
        if c not in cache:
            buf = np.fromfile(os.path.join(arena_dir, f"arena_{c:04d}.bin"), dtype=np.uint8)
            magic, version, fdim, fprec = struct.unpack("<IIIB", buf[:13].tobytes())
            if magic != MAGIC:
                raise ValueError("not a valid arena (magic mismatch)")
            if version != VERSION:
                raise ValueError(f"unsupported version {version}")
            if fdim != dim or fprec != precision:
                raise ValueError("dimension / precision mismatch")
            cache[c] = buf
        off = HEADER + (p % vpc) * vsize
        chunk = cache[c]
        raw = np.zeros(vsize, np.uint8)
        avail = chunk[off:off + vsize]
        raw[:avail.size] = avail
        out[i] = raw.view(dt)
    return out
