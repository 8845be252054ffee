"""Host-side mirror of the refresher (include/kektordb_gpu.h, csrc/refresher.cpp): the staleness policy that
keeps the GPU mirror behind the CPU index by at most max_lag_ms / max_pending_rows while Add / Delete / Vacuum /
Refine (reference hnsw_index.go:472-809, :2303-2336, optimizer.go:118-468) keep changing it."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import ffi
from .index import GpuIndex, _ptr, _RAW_DTYPES


@dataclass
class RefresherStats:
    pending_rows: int
    pending_nodes: int
    flushes: int
    flushes_by_rows: int
    flushes_by_lag: int
    flushes_by_call: int
    rows_queued: int
    rows_applied: int
    nodes_applied: int
    oldest_pending_ms: float
    last_flush_ms: float
    last_error: int


class Refresher:
    def __init__(self, index: GpuIndex, max_pending_rows: int = 4096, max_lag_ms: int = 50):
        self._lib = ffi.lib()
        self._index = index
        h = C.c_void_p()
        ffi.check(self._lib.kdbgpu_refresher_create(index._handle(), max_pending_rows, max_lag_ms, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            rc = self._lib.kdbgpu_refresher_destroy(self._h)
            self._h = None
            ffi.check(rc)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_node(self, node_id: int, level: int, row_raw) -> None:
        row = np.ascontiguousarray(row_raw, dtype=_RAW_DTYPES[self._index.precision])
        if row.shape != (self._index.dim,):
            raise ValueError(f"row must be [{self._index.dim}] in stored form")
        ffi.check(self._lib.kdbgpu_refresher_add_node(self._h, node_id, level, _ptr(row)))

    def set_row(self, node_id: int, level: int, nbrs) -> None:
        nb = np.ascontiguousarray(nbrs, dtype=np.uint32)
        ffi.check(self._lib.kdbgpu_refresher_set_row(self._h, node_id, level, _ptr(nb) if nb.size else None, nb.size))

    def remove_node(self, node_id: int) -> None:
        ffi.check(self._lib.kdbgpu_refresher_remove_node(self._h, node_id))

    def set_deleted(self, node_id: int, deleted: bool = True) -> None:
        ffi.check(self._lib.kdbgpu_refresher_set_deleted(self._h, node_id, 1 if deleted else 0))

    def set_entry(self, entry: int, max_level: int) -> None:
        ffi.check(self._lib.kdbgpu_refresher_set_entry(self._h, entry, max_level))

    def flush(self) -> None:
        ffi.check(self._lib.kdbgpu_refresher_flush(self._h))

    def stats(self) -> RefresherStats:
        st = ffi.RefresherStats()
        ffi.check(self._lib.kdbgpu_refresher_stats(self._h, C.byref(st)))
        return RefresherStats(*(getattr(st, f[0]) for f in ffi.RefresherStats._fields_))
