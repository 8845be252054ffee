"""Id-range shards: which ids a shard owns, and `ShardGroup` — the host-side mirror of the
sharded SearchWithScores the C ABI offers (kdbgpu_shard_*).  The per-shard traversal, the exchange
(NCCL all-gather between ranks, peer copies inside one process) and the merge all run inside
libkektordb_gpu; this module only moves numpy buffers across the boundary."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import ffi
from .index import GpuIndex, _ptr, effective_ef


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous id range [base, base + count) of 0-based corpus rows owned by `rank`;
    the remainder goes to the last ranks one row each, so sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    q, r = divmod(n, world)
    count = q + (1 if rank >= world - r else 0)
    base = rank * q + max(0, rank - (world - r))
    return base, count


def globalize_ids(local_ids: np.ndarray, counts: np.ndarray, base: int) -> np.ndarray:
    """Local internal ids (1-based, 0 = empty slot) -> global internal ids (1-based)."""
    out = local_ids.astype(np.int64, copy=True)
    k = out.shape[1]
    valid = np.arange(k)[None, :] < np.asarray(counts)[:, None]
    out[valid] += base
    out[~valid] = 0
    return out.astype(np.uint32)


@dataclass
class ShardStats:
    dist_evals: int = 0
    hops: int = 0
    hops_l0: int = 0
    traversal_ms: float = 0.0
    exchange_ms: float = 0.0
    merge_ms: float = 0.0
    total_ms: float = 0.0
    n_shards: int = 0


def _stats(st: ffi.ShardStats) -> ShardStats:
    return ShardStats(st.dist_evals, st.hops, st.hops_l0, st.traversal_ms, st.exchange_ms, st.merge_ms, st.total_ms,
                      st.n_shards)


def unique_id() -> bytes:
    """ncclGetUniqueId through the library: rank 0 calls it and hands the bytes to every rank."""
    buf = (C.c_ubyte * ffi.SHARD_ID_BYTES)()
    ffi.check(ffi.lib().kdbgpu_shard_unique_id(buf))
    return bytes(buf)


class ShardGroup:
    """G id-range shards answering SearchWithScores together.

    ShardGroup.local(indexes, id_bases)            every shard in this process (peer copies, no NCCL)
    ShardGroup.rank(index, rank, world, uid, base)  one shard per process; calls are collective (NCCL)
    """

    def __init__(self, handle, members):
        self._lib = ffi.lib()
        self._g = handle
        self._members = members  # keeps the indexes alive
        self.needs_refine = False

    @classmethod
    def local(cls, indexes: list[GpuIndex], id_bases) -> "ShardGroup":
        n = len(indexes)
        arr = (C.c_void_p * n)(*[ix._handle() for ix in indexes])
        bases = np.ascontiguousarray(id_bases, dtype=np.uint32)
        if bases.shape != (n,):
            raise ValueError("one id base per shard")
        g = C.c_void_p()
        ffi.check(ffi.lib().kdbgpu_shard_group_create_local(arr, n, _ptr(bases), C.byref(g)))
        return cls(g, list(indexes))

    @classmethod
    def rank(cls, index: GpuIndex, rank: int, world: int, uid: bytes, id_base: int) -> "ShardGroup":
        if len(uid) != ffi.SHARD_ID_BYTES:
            raise ValueError("uid must be the 128 bytes of unique_id()")
        buf = (C.c_ubyte * ffi.SHARD_ID_BYTES).from_buffer_copy(uid)
        g = C.c_void_p()
        ffi.check(ffi.lib().kdbgpu_shard_group_create_rank(index._handle(), rank, world, buf, id_base, C.byref(g)))
        return cls(g, [index])

    def close(self) -> None:
        if getattr(self, "_g", None):
            self._lib.kdbgpu_shard_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self) -> int:
        return int(self._lib.kdbgpu_shard_group_size(self._g))

    @property
    def dim(self) -> int:
        return self._members[0].dim

    def _out(self, nq, k):
        return (np.zeros((nq, k), dtype=np.uint32), np.zeros((nq, k), dtype=np.float64), np.zeros(nq, dtype=np.uint32))

    def SearchWithScores(self, query, k: int, allowList: np.ndarray | None = None, efSearch: int = 0):
        """Batched SearchWithScores over the sharded corpus; allowList is a dense bitset over GLOBAL ids.
        Returns (global ids [nq,k], scores [nq,k], counts [nq], ShardStats)."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}]")
        ids, scores, counts = self._out(q.shape[0], k)
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        st = ffi.ShardStats()
        ffi.check(self._lib.kdbgpu_shard_search_batch(self._g, _ptr(q), q.shape[0], k,
                                                      effective_ef(int(efSearch), self.needs_refine), _ptr(allow),
                                                      0 if allow is None else allow.size, _ptr(ids), _ptr(scores),
                                                      _ptr(counts), C.byref(st)))
        return ids, scores, counts, _stats(st)

    def submit(self, query: np.ndarray, k: int, efSearch: int = 0, allowList: np.ndarray | None = None):
        """kdbgpu_shard_search_submit: queue one batch, return a ticket for wait().  `query` must be a
        C-contiguous float32 [nq, dim] array that stays alive until wait()."""
        if query.dtype != np.float32 or not query.flags.c_contiguous or query.ndim != 2 or query.shape[1] != self.dim:
            raise ValueError("query must be C-contiguous float32 [nq, dim]")
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        t = C.c_void_p()
        ffi.check(self._lib.kdbgpu_shard_search_submit(self._g, _ptr(query), query.shape[0], k,
                                                       effective_ef(int(efSearch), self.needs_refine), _ptr(allow),
                                                       0 if allow is None else allow.size, C.byref(t)))
        return (t, query, allow, query.shape[0], k)

    def wait(self, ticket):
        t, _q, _a, nq, k = ticket
        ids, scores, counts = self._out(nq, k)
        st = ffi.ShardStats()
        ffi.check(self._lib.kdbgpu_shard_search_wait(t, _ptr(ids), _ptr(scores), _ptr(counts), C.byref(st)))
        return ids, scores, counts, _stats(st)

    def search_device(self, d_queries_ptr: int, nq: int, k: int, ef_search: int, d_ids_ptr: int, d_scores_ptr: int,
                      d_counts_ptr: int, stream_ptr: int = 0) -> None:
        ffi.check(self._lib.kdbgpu_shard_search_batch_device(
            self._g, C.c_void_p(d_queries_ptr), nq, k, effective_ef(int(ef_search), self.needs_refine),
            C.c_void_p(d_ids_ptr), C.c_void_p(d_scores_ptr), C.c_void_p(d_counts_ptr),
            C.c_void_p(stream_ptr) if stream_ptr else None))

    def sync(self) -> ShardStats:
        st = ffi.ShardStats()
        ffi.check(self._lib.kdbgpu_shard_sync(self._g, C.byref(st)))
        return _stats(st)

    def flat_search(self, query, k: int, mode: int = 0, allowList: np.ndarray | None = None, prefilter: bool = False):
        if prefilter:
            mode |= ffi.FLAT_PREFILTER
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        ids, scores, counts = self._out(q.shape[0], k)
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        st = ffi.ShardStats()
        ffi.check(self._lib.kdbgpu_shard_flat_search_batch(self._g, _ptr(q), q.shape[0], k, mode, _ptr(allow),
                                                           0 if allow is None else allow.size, _ptr(ids), _ptr(scores),
                                                           _ptr(counts), C.byref(st)))
        return ids, scores, counts, _stats(st)
