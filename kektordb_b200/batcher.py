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
                 fn=None, dim: int | None = None):
        self._lib = ffi.lib()
        self._index = index
        self._cb = None
        self.dim = int(index.dim if index is not None else dim)
        h = C.c_void_p()
        if index is not None:
            ffi.check(self._lib.kdbgpu_batcher_create(index._handle(), max_batch, max_wait_us, C.byref(h)))
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

    def stats(self) -> BatcherStats:
        st = ffi.BatcherStats()
        ffi.check(self._lib.kdbgpu_batcher_stats(self._h, C.byref(st)))
        return BatcherStats(st.queries, st.batches, st.max_batch_seen, st.dispatched_idle, st.dispatched_full,
                            st.dispatched_deadline)
