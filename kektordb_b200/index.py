"""Host-side mirror of the reference's hnsw.Index search surface over the C ABI.

Same names and argument meaning as the reference (pkg/core/hnsw/hnsw_index.go):
`SearchWithScores(query, k, allowList, efSearch)` (:343), metric strings "euclidean" /
"cosine" (pkg/core/distance/distance_go.go:34-39), internal uint32 ids from 1, scores = raw
float64 distances, failures yield empty results (:355-359).  The only addition is that a call
takes a batch of queries — the batch a Go-side micro-batcher would form.

All compute happens in libkektordb_gpu.so (hand-written sm_100a CUDA); this module only moves
numpy buffers across the boundary.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import ffi

Euclidean = "euclidean"
Cosine = "cosine"
_METRICS = {Euclidean: ffi.METRIC_L2, Cosine: ffi.METRIC_COSINE, "l2": ffi.METRIC_L2,
            ffi.METRIC_L2: ffi.METRIC_L2, ffi.METRIC_COSINE: ffi.METRIC_COSINE}


Float32, Float16, Int8 = "float32", "float16", "int8"  # distance.PrecisionType (distance_go.go:41-46)
_PRECISIONS = {Float32: ffi.PRECISION_F32, Float16: ffi.PRECISION_F16, Int8: ffi.PRECISION_INT8,
               ffi.PRECISION_F32: ffi.PRECISION_F32, ffi.PRECISION_F16: ffi.PRECISION_F16,
               ffi.PRECISION_INT8: ffi.PRECISION_INT8}
_PRECISION_NAMES = {ffi.PRECISION_F32: Float32, ffi.PRECISION_F16: Float16, ffi.PRECISION_INT8: Int8}
_RAW_DTYPES = {ffi.PRECISION_F32: np.float32, ffi.PRECISION_F16: np.uint16, ffi.PRECISION_INT8: np.int8}


@dataclass
class SearchStats:
    dist_evals: int = 0
    hops: int = 0
    hops_l0: int = 0
    kernel_ms: float = 0.0
    total_ms: float = 0.0
    heap_pass_queries: int = 0


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def effective_ef(ef_search: int, needs_refine: bool) -> int:
    """The needsRefine boost of searchInternal (hnsw_index.go:387-399), applied by the caller."""
    actual = ef_search
    if needs_refine:
        boosted = int(float(ef_search) * 2)
        boosted = max(boosted, 80)
        boosted = min(boosted, 200)
        if boosted > actual:
            actual = boosted
    return actual


def dense_allow_list(ids, n: int) -> np.ndarray:
    """Dense uint64 bitset over internal ids 0..n with the membership of a roaring bitmap."""
    words = np.zeros((n >> 6) + 1, dtype=np.uint64)
    ids = np.asarray(ids, dtype=np.uint64)
    if ids.size:
        np.bitwise_or.at(words, (ids >> np.uint64(6)).astype(np.int64), np.uint64(1) << (ids & np.uint64(63)))
    return words


def arena_probe(arena_dir: str):
    """(dim, precision name, n_chunks, vecs_per_chunk) of a reference vector arena directory; host-only."""
    dim, prec, nch, vpc = C.c_uint32(0), C.c_int(0), C.c_uint32(0), C.c_uint32(0)
    ffi.check(ffi.lib().kdbgpu_arena_probe(arena_dir.encode(), C.byref(dim), C.byref(prec), C.byref(nch), C.byref(vpc)))
    return dim.value, _PRECISION_NAMES[prec.value], nch.value, vpc.value


class GpuIndex:
    """GPU mirror of one hnsw.Index: corpus rows + adjacency staged in HBM, searched on device."""

    def __init__(self, dim: int, metric, m: int = 16, capacity: int = 1 << 20, device: int = 0,
                 precision=Float32):
        if precision not in _PRECISIONS:
            raise ValueError(f"unknown precision '{precision}'")
        if metric not in _METRICS:  # distance_go.go:155-157
            raise ValueError(f"metric '{metric}' not supported for {_PRECISION_NAMES[_PRECISIONS[precision]]} precision")
        self._lib = ffi.lib()
        self.dim, self.m, self.capacity, self.device = int(dim), int(m) if m > 0 else 16, int(capacity), int(device)
        self.metric = _METRICS[metric]
        self.precision = _PRECISIONS[precision]
        self.needs_refine = False
        h = C.c_void_p()
        ffi.check(self._lib.kdbgpu_index_create_ex(device, dim, self.metric, self.precision, m, capacity, C.byref(h)))
        self._h = h

    # -- lifecycle -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.kdbgpu_index_destroy(self._h)
            self._h = None

    Close = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise ffi.GpuError(ffi.ERR_STATE, "index is closed")
        return self._h

    # -- staging ---------------------------------------------------------------------------
    def upload_vectors(self, first_id: int, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [count, {self.dim}]")
        ffi.check(self._lib.kdbgpu_upload_vectors(self._handle(), first_id, rows.shape[0], _ptr(rows)))

    def upload_rows_raw(self, first_id: int, rows: np.ndarray) -> None:
        """Rows already in stored form (float32 / float16 bits as uint16 / int8), as the arena holds them."""
        rows = np.ascontiguousarray(rows, dtype=_RAW_DTYPES[self.precision])
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [count, {self.dim}]")
        ffi.check(self._lib.kdbgpu_upload_rows_raw(self._handle(), first_id, rows.shape[0], _ptr(rows)))

    def download_rows_raw(self, first_id: int, count: int) -> np.ndarray:
        out = np.zeros((count, self.dim), dtype=_RAW_DTYPES[self.precision])
        ffi.check(self._lib.kdbgpu_download_rows_raw(self._handle(), first_id, count, _ptr(out)))
        return out

    def download_norms(self, first_id: int, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.float32)
        ffi.check(self._lib.kdbgpu_download_norms(self._handle(), first_id, count, _ptr(out)))
        return out

    def load_arena(self, arena_dir: str, slot_table: np.ndarray | None = None, n: int | None = None) -> int:
        """Stage the rows of a reference vector arena (pkg/storage/mmap/arena.go) from its chunk files.
        slot_table = ArenaState.SlotTable (uint32, indexed by internal id) or None for sequential slots
        of ids 1..n.  Returns the number of rows staged."""
        staged = C.c_uint64(0)
        if slot_table is None:
            if n is None:
                raise ValueError("n is needed without a slot table")
            ffi.check(self._lib.kdbgpu_arena_load_dir(self._handle(), arena_dir.encode(), None, n + 1, C.byref(staged)))
        else:
            st = np.ascontiguousarray(slot_table, dtype=np.uint32)
            ffi.check(self._lib.kdbgpu_arena_load_dir(self._handle(), arena_dir.encode(), _ptr(st), st.size, C.byref(staged)))
        return int(staged.value)

    def stage_arena_chunk(self, chunk_id: int, chunk: np.ndarray, slot_table: np.ndarray | None, table_len: int) -> int:
        """One chunk the host already holds (header included), e.g. the Go side's mmap of arena_%04d.bin."""
        buf = np.ascontiguousarray(chunk, dtype=np.uint8)
        st = None if slot_table is None else np.ascontiguousarray(slot_table, dtype=np.uint32)
        staged = C.c_uint32(0)
        ffi.check(self._lib.kdbgpu_arena_stage_chunk(self._handle(), chunk_id, _ptr(buf), buf.size, _ptr(st), table_len,
                                                     C.byref(staged)))
        return int(staged.value)

    def set_quantizer(self, abs_max: float) -> None:
        """Quantizer.AbsMax of an int8 index (pkg/core/distance/quantizer.go:19-22)."""
        ffi.check(self._lib.kdbgpu_set_quantizer(self._handle(), float(abs_max)))

    def TrainQuantizer(self, vectors) -> float:
        """(*Index).TrainQuantizer (hnsw_index.go:2962-2966) -> Quantizer.Train; returns AbsMax."""
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"vectors must be [n, {self.dim}]")
        am = C.c_float(0.0)
        ffi.check(self._lib.kdbgpu_train_quantizer(self._handle(), _ptr(v), v.shape[0], C.byref(am)))
        return float(am.value)

    def train_quantizer_device(self, d_rows_ptr: int, row_stride: int, n: int) -> float:
        am = C.c_float(0.0)
        ffi.check(self._lib.kdbgpu_train_quantizer_device(self._handle(), C.c_void_p(d_rows_ptr), row_stride, n,
                                                          C.byref(am)))
        return float(am.value)

    def upload_vectors_device(self, first_id: int, d_rows_ptr: int, count: int, row_stride: int) -> None:
        ffi.check(self._lib.kdbgpu_upload_vectors_device(self._handle(), first_id, count, C.c_void_p(d_rows_ptr),
                                                         row_stride))

    def set_graph(self, n: int, levels, node_row, row_off, nbrs, entry: int, max_level: int) -> None:
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        node_row = np.ascontiguousarray(node_row, dtype=np.uint64)
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
        nbrs = np.ascontiguousarray(nbrs, dtype=np.uint32)
        if levels.shape != (n + 1,) or node_row.shape != (n + 2,):
            raise ValueError("levels must be [n+1] and node_row [n+2]")
        if nbrs.size == 0:
            nbrs = np.zeros(1, dtype=np.uint32)
        ffi.check(self._lib.kdbgpu_set_graph(self._handle(), n, _ptr(levels), _ptr(node_row), _ptr(row_off),
                                             _ptr(nbrs), entry, max_level))

    def save_graph_file(self, path: str) -> None:
        """The mirror's topology as a flat sidecar file (layout in include/kektordb_gpu.h)."""
        ffi.check(self._lib.kdbgpu_save_graph_file(self._handle(), path.encode()))

    def set_graph_file(self, path: str) -> None:
        """kdbgpu_set_graph from a sidecar file: mapped and staged inside the library."""
        ffi.check(self._lib.kdbgpu_set_graph_file(self._handle(), path.encode()))

    # -- incremental refresh (follow the CPU index's Add / Vacuum / Refine) ----------------------------
    def register_nodes(self, first_id: int, levels) -> None:
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        ffi.check(self._lib.kdbgpu_register_nodes(self._handle(), first_id, lv.size, _ptr(lv)))

    def patch_rows(self, ids, row_levels, rows) -> None:
        """rows: list of neighbour-id sequences, one per (ids[i], row_levels[i])."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        lv = np.ascontiguousarray(row_levels, dtype=np.int32)
        off = np.zeros(len(rows) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in rows])
        flat = np.concatenate([np.asarray(r, dtype=np.uint32) for r in rows]) if len(rows) and off[-1] else np.zeros(1, np.uint32)
        flat = np.ascontiguousarray(flat, dtype=np.uint32)
        ffi.check(self._lib.kdbgpu_patch_rows(self._handle(), ids.size, _ptr(ids), _ptr(lv), _ptr(off), _ptr(flat)))

    def remove_nodes(self, ids) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        ffi.check(self._lib.kdbgpu_remove_nodes(self._handle(), ids.size, _ptr(ids)))

    def set_entry(self, entry: int, max_level: int) -> None:
        ffi.check(self._lib.kdbgpu_set_entry(self._handle(), entry, max_level))

    def set_deleted(self, bitset: np.ndarray | None) -> None:
        if bitset is None:
            ffi.check(self._lib.kdbgpu_set_deleted(self._handle(), None, 0))
            return
        b = np.ascontiguousarray(bitset, dtype=np.uint64)
        ffi.check(self._lib.kdbgpu_set_deleted(self._handle(), _ptr(b), b.size))

    # -- construction on the device (extension of the reference surface) ---------------------------
    def AddBatch(self, vectors, level_draws, ef_construction: int = 0) -> None:
        """(*Index).AddBatch (hnsw_index.go:1466): ids are assigned in order from count+1;
        level_draws are the rand.Float64() values of randomLevel."""
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        u = np.ascontiguousarray(level_draws, dtype=np.float64)
        if v.ndim != 2 or v.shape[1] != self.dim or u.shape != (v.shape[0],):
            raise ValueError(f"vectors must be [count, {self.dim}] with one level draw each")
        ffi.check(self._lib.kdbgpu_add_batch(self._handle(), v.shape[0], _ptr(v), _ptr(u), ef_construction))

    def add_batch_device(self, d_rows_ptr: int, count: int, row_stride: int, level_draws,
                         ef_construction: int = 0) -> None:
        u = np.ascontiguousarray(level_draws, dtype=np.float64)
        assert u.shape == (count,)
        ffi.check(self._lib.kdbgpu_add_batch_device(self._handle(), count, C.c_void_p(d_rows_ptr), row_stride,
                                                    _ptr(u), ef_construction))

    def get_graph(self):
        """Topology read-back: (n, levels, node_row, row_off, nbrs, entry, max_level)."""
        n, rows, edges = C.c_uint32(), C.c_uint64(), C.c_uint64()
        entry, max_level = C.c_uint32(), C.c_int()
        ffi.check(self._lib.kdbgpu_get_graph_sizes(self._handle(), C.byref(n), C.byref(rows), C.byref(edges),
                                                   C.byref(entry), C.byref(max_level)))
        levels = np.zeros(n.value + 1, dtype=np.int32)
        node_row = np.zeros(n.value + 2, dtype=np.uint64)
        row_off = np.zeros(rows.value + 1, dtype=np.uint64)
        nbrs = np.zeros(max(edges.value, 1), dtype=np.uint32)
        ffi.check(self._lib.kdbgpu_get_graph(self._handle(), _ptr(levels), _ptr(node_row), _ptr(row_off), _ptr(nbrs)))
        return n.value, levels, node_row, row_off, nbrs[: edges.value], entry.value, max_level.value

    def download_vectors(self, first_id: int, count: int) -> np.ndarray:
        out = np.zeros((count, self.dim), dtype=np.float32)
        ffi.check(self._lib.kdbgpu_download_vectors(self._handle(), first_id, count, _ptr(out)))
        return out

    def set_tuning(self, slots: int = 0, cand_smem: int = 0, max_ctas_per_sm: int = -1) -> None:
        ffi.check(self._lib.kdbgpu_set_tuning(self._handle(), slots, cand_smem, max_ctas_per_sm))

    def set_idle_slots(self, slots_idle: int) -> None:
        """Slot count of a launch that finds no other batch in flight (0 = same as the throughput shape)."""
        ffi.check(self._lib.kdbgpu_set_idle_slots(self._handle(), slots_idle))

    # -- query -----------------------------------------------------------------------------
    def SearchWithScores(self, query, k: int, allowList: np.ndarray | None = None, efSearch: int = 0):
        """Batched (*Index).SearchWithScores.  `query` is [nq, dim] (or [dim]); allowList a dense
        uint64 bitset or None.  Returns (ids [nq,k] uint32, scores [nq,k] float64, counts [nq], stats)."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        single = q.ndim == 1
        if single:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}]")
        nq = q.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        scores = np.zeros((nq, k), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.uint32)
        st = ffi.Stats()
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        ef = effective_ef(int(efSearch), self.needs_refine)
        ffi.check(self._lib.kdbgpu_search_batch(self._handle(), _ptr(q), nq, k, ef, _ptr(allow),
                                                0 if allow is None else allow.size, _ptr(ids), _ptr(scores),
                                                _ptr(counts), C.byref(st)))
        stats = SearchStats(st.dist_evals, st.hops, st.hops_l0, st.kernel_ms, st.total_ms, st.heap_pass_queries)
        return ids, scores, counts, stats

    search_with_scores = SearchWithScores

    def search_device(self, d_queries_ptr: int, nq: int, k: int, ef_search: int, d_ids_ptr: int, d_scores_ptr: int,
                      d_counts_ptr: int, stream_ptr: int = 0, d_allow_ptr: int = 0, allow_words: int = 0,
                      allow_first_id: int = 0) -> None:
        """kdbgpu_search_batch_device: every buffer already resident on the handle's device; the
        launch is queued on `stream_ptr` (a cudaStream_t) without host synchronisation."""
        ffi.check(self._lib.kdbgpu_search_batch_device(
            self._handle(), C.c_void_p(d_queries_ptr), nq, k, effective_ef(int(ef_search), self.needs_refine),
            C.c_void_p(d_allow_ptr) if d_allow_ptr else None, allow_words, allow_first_id, C.c_void_p(d_ids_ptr),
            C.c_void_p(d_scores_ptr), C.c_void_p(d_counts_ptr), C.c_void_p(stream_ptr) if stream_ptr else None))

    def last_search_stats(self) -> SearchStats:
        st = ffi.Stats()
        ffi.check(self._lib.kdbgpu_last_search_stats(self._handle(), C.byref(st)))
        return SearchStats(st.dist_evals, st.hops, st.hops_l0, 0.0, 0.0)

    def distance_batch(self, prepared_query, ids) -> np.ndarray:
        q = np.ascontiguousarray(prepared_query, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.zeros(ids.size, dtype=np.float64)
        ffi.check(self._lib.kdbgpu_distance_batch(self._handle(), _ptr(q), _ptr(ids), ids.size, _ptr(out)))
        return out

    def flat_search(self, query, k: int, mode: int = 0, allowList: np.ndarray | None = None,
                    prefilter: bool = False):
        """BruteForceIndex.SearchWithScores (mode 0) / exact f64 ground truth (mode 1).  With
        `prefilter` the same answer is found through the tensor-core nomination pass
        (KDBGPU_FLAT_PREFILTER); stats.hops then counts queries the exhaustive scan answered."""
        if prefilter:
            mode |= ffi.FLAT_PREFILTER
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        scores = np.zeros((nq, k), dtype=np.float64)
        counts = np.zeros(nq, dtype=np.uint32)
        st = ffi.Stats()
        allow = None if allowList is None else np.ascontiguousarray(allowList, dtype=np.uint64)
        ffi.check(self._lib.kdbgpu_flat_search_batch(self._handle(), _ptr(q), nq, k, mode, _ptr(allow),
                                                     0 if allow is None else allow.size, _ptr(ids), _ptr(scores),
                                                     _ptr(counts), C.byref(st)))
        return ids, scores, counts, SearchStats(st.dist_evals, st.hops, st.hops_l0, st.kernel_ms, st.total_ms)

    def flat_prefilter_scores(self, query, mode: int = 0):
        """Validation hook: approximate scores [nq, n] of the tensor-core pass and the certified
        per-query bound on their error."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        n = self.count
        out = np.zeros((q.shape[0], n), dtype=np.float32)
        bound = np.zeros(q.shape[0], dtype=np.float32)
        ffi.check(self._lib.kdbgpu_flat_prefilter_scores(self._handle(), _ptr(q), q.shape[0], mode, _ptr(out),
                                                         _ptr(bound)))
        return out, bound

    # -- introspection ---------------------------------------------------------------------
    def Metric(self) -> str:
        return Cosine if self.metric == ffi.METRIC_COSINE else Euclidean

    def Precision(self) -> str:
        return _PRECISION_NAMES[self.precision]

    def GetDimension(self) -> int:
        return self.dim

    @property
    def count(self) -> int:
        return int(self._lib.kdbgpu_index_count(self._handle()))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.kdbgpu_index_device_bytes(self._handle()))

    def set_fast_path(self, mode) -> None:
        """0 / False: heap pass only; 1 / True: default (fast pass for int8); 2: fast pass for every precision."""
        ffi.check(self._lib.kdbgpu_set_fast_path(self._handle(), int(mode)))

    def prepare_search(self, nq: int, k: int, ef_search: int) -> None:
        ffi.check(self._lib.kdbgpu_prepare_search(self._handle(), nq, k, effective_ef(int(ef_search), self.needs_refine)))

    def search_concurrency(self, k: int, ef_search: int) -> int:
        return int(self._lib.kdbgpu_search_concurrency(self._handle(), k, ef_search))
