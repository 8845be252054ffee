"""Host-side mirror of the micro-batcher (include/kektordb_gpu.h, csrc/batcher.cpp): the reference's
call shape — ONE query per blocking call, any number of caller threads, as every request goroutine
does through idx.SearchWithScores (reference pkg/engine/ops.go:1006) — on top of the batched device
entry point.  ctypes releases the GIL inside the C call, so Python threads batch for real."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import ffi
from .index import GpuIndex, _ptr, effective_ef


@dataclass
class BatcherStats:
    queries: int
    batches: int
    max_batch_seen: int
    dispatched_idle: int
    dispatched_full: int
    dispatched_deadline: int

    @property
    def mean_batch(self) -> float:
        return self.queries / self.batches if self.batches else 0.0


class Batcher:
    """kdbgpu_batcher over one GpuIndex, or over any batch executor `fn(queries[nq,dim], k, ef, allow)
    -> (ids[nq,k], scores[nq,k], counts[nq])` (create_fn: sharded searches, test doubles)."""

    def __init__(self, index: GpuIndex | None = None, max_batch: int = 1024, max_wait_us: int = 200,
                 fn=None, dim: int | None = None, group=None):
        self._lib = ffi.lib()
        self._index = index
        self._group = group
        self._cb = None
        self.dim = int(index.dim if index is not None else (group.dim if group is not None else dim))
        h = C.c_void_p()
        if index is not None:
            ffi.check(self._lib.kdbgpu_batcher_create(index._handle(), max_batch, max_wait_us, C.byref(h)))
        elif group is not None:  # a local shard group as the executor: one call per query in front of G GPUs
            ffi.check(self._lib.kdbgpu_batcher_create_group(group._g, self.dim, max_batch, max_wait_us, C.byref(h)))
        else:
            if fn is None or dim is None:
                raise ValueError("either an index or (fn, dim)")

            def tramp(_ctx, q, nq, k, ef, allow, allow_words, out_ids, out_scores, out_counts):
                try:
                    qa = np.ctypeslib.as_array(q, shape=(nq, self.dim))
                    al = np.ctypeslib.as_array(allow, shape=(allow_words,)) if allow and allow_words else None
                    ids, sc, cnt = fn(qa, k, ef, al)
                    np.ctypeslib.as_array(out_ids, shape=(nq, k))[:] = ids
                    np.ctypeslib.as_array(out_scores, shape=(nq, k))[:] = sc
                    np.ctypeslib.as_array(out_counts, shape=(nq,))[:] = cnt
                    return ffi.OK
                except ffi.GpuError as ex:
                    return ex.code
                except Exception:
                    return ffi.ERR_STATE

            self._cb = ffi.BATCH_FN(tramp)  # keep the callback alive
            ffi.check(self._lib.kdbgpu_batcher_create_fn(self._cb, None, self.dim, max_batch, max_wait_us, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.kdbgpu_batcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SearchWithScores(self, query, k: int, allowList: np.ndarray | None = None, efSearch: int = 0):
        """(*Index).SearchWithScores (hnsw_index.go:343) for ONE query; blocks until its batch is done.
        Returns (ids[count], scores[count]); a failed search yields empty arrays (:355-359)."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.shape != (self.dim,):
            raise ValueError(f"query must be [{self.dim}]")
        ids = np.zeros(k, dtype=np.uint32)
        sc = np.zeros(k, dtype=np.float64)
        cnt = C.c_uint32(0)
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        needs_refine = bool(self._index.needs_refine) if self._index is not None else False
        rc = self._lib.kdbgpu_batcher_search(self._h, _ptr(q), k, effective_ef(int(efSearch), needs_refine), _ptr(allow),
                                             0 if allow is None else allow.size, _ptr(ids), _ptr(sc), C.byref(cnt))
        self.last_rc = rc
        n = cnt.value if rc == ffi.OK else 0
        return ids[:n], sc[:n]

    # -- asynchronous form: submit / poll / take (what a Go dispatcher goroutine drives) ------------------
    def register_filter(self, allowList: np.ndarray) -> int:
        allow = np.ascontiguousarray(allowList, dtype=np.uint64)
        fid = C.c_uint64(0)
        ffi.check(self._lib.kdbgpu_batcher_register_filter(self._h, _ptr(allow), allow.size, C.byref(fid)))
        return int(fid.value)

    def release_filter(self, filter_id: int) -> None:
        ffi.check(self._lib.kdbgpu_batcher_release_filter(self._h, filter_id))

    def submit(self, query, k: int, efSearch: int = 0, allowList: np.ndarray | None = None, filter_id: int = 0) -> int:
        """Queues ONE query and returns its ticket at once (the query is copied; nothing of the caller's is kept)."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.shape != (self.dim,):
            raise ValueError(f"query must be [{self.dim}]")
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        needs_refine = bool(self._index.needs_refine) if self._index is not None else False
        t = C.c_uint64(0)
        ffi.check(self._lib.kdbgpu_batcher_submit(self._h, _ptr(q), k, effective_ef(int(efSearch), needs_refine), _ptr(allow),
                                                  0 if allow is None else allow.size, filter_id, C.byref(t)))
        return int(t.value)

    def poll(self, max_tickets: int = 1024, timeout_us: int = 1000) -> list[int]:
        """Tickets of queries that finished since the last poll (waits up to timeout_us for the first)."""
        buf = (C.c_uint64 * max_tickets)()
        n = C.c_uint32(0)
        ffi.check(self._lib.kdbgpu_batcher_poll(self._h, buf, max_tickets, timeout_us, C.byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def take(self, ticket: int, k: int):
        """Result of a ticket: (ids[count], scores[count], rc).  Every ticket is taken exactly once."""
        ids = np.zeros(k, dtype=np.uint32)
        sc = np.zeros(k, dtype=np.float64)
        cnt = C.c_uint32(0)
        rc = self._lib.kdbgpu_batcher_take(self._h, ticket, _ptr(ids), _ptr(sc), C.byref(cnt))
        n = cnt.value if rc == ffi.OK else 0
        return ids[:n], sc[:n], rc

    def stats(self) -> BatcherStats:
        st = ffi.BatcherStats()
        ffi.check(self._lib.kdbgpu_batcher_stats(self._h, C.byref(st)))
        return BatcherStats(st.queries, st.batches, st.max_batch_seen, st.dispatched_idle, st.dispatched_full,
                            st.dispatched_deadline)
