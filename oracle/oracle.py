"""ctypes binding of the CPU ORACLE (oracle/kdb_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (kektordb_b200) must never
import this module.

The oracle restates, in plain C, the reference's HNSW search/insert/flat semantics
(/root/reference pkg/core/hnsw/hnsw_index.go, hnsw_heap.go, bitset.go,
pkg/core/distance/distance_go.go, native/compute/src/lib.rs, pkg/core/vector_index.go).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkdb_oracle.so")

METRIC_L2 = 0
METRIC_COSINE = 1
ARITH_SEQ = 0      # pure-Go sequential f32 loops (distance_go.go:57-89)
ARITH_AVX2 = 1     # Rust AVX2/FMA order (native/compute/src/lib.rs:22-99)
ARITH_KERNEL = 2   # summation order of the sm_100a kernel
PREC_F32 = 0       # distance.Float32 (distance_go.go:41-46)
PREC_F16 = 1       # distance.Float16 — Euclidean only
PREC_I8 = 2        # distance.Int8 — Cosine only
_RAW_DTYPE = {PREC_F32: np.float32, PREC_F16: np.uint16, PREC_I8: np.int8}


def build(force: bool = False) -> str:
    """Compile oracle/libkdb_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "kdb_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("kdb_oracle.c", "kdb_oracle.h", "Makefile")
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    assert os.path.exists(src)
    return _LIB_PATH


class _Stats(C.Structure):
    _fields_ = [("dist_evals", C.c_uint64), ("hops", C.c_uint64), ("hops_l0", C.c_uint64)]


@dataclass
class Stats:
    dist_evals: int = 0
    hops: int = 0
    hops_l0: int = 0


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, u32p, f32p, f64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.kdbo_new.restype = vp
    L.kdbo_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
    L.kdbo_new_ex.restype = vp
    L.kdbo_new_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
    L.kdbo_free.argtypes = [vp]
    L.kdbo_precision.restype = C.c_int
    L.kdbo_precision.argtypes = [vp]
    L.kdbo_set_quantizer.argtypes = [vp, C.c_float]
    L.kdbo_abs_max.restype = C.c_float
    L.kdbo_abs_max.argtypes = [vp]
    L.kdbo_f32_to_f16.restype = C.c_uint16
    L.kdbo_f32_to_f16.argtypes = [C.c_float]
    L.kdbo_f16_to_f32.restype = C.c_float
    L.kdbo_f16_to_f32.argtypes = [C.c_uint16]
    L.kdbo_sq_euclid_f16.restype = C.c_float
    L.kdbo_sq_euclid_f16.argtypes = [C.c_int, vp, vp, C.c_size_t]
    L.kdbo_dot_i8.restype = C.c_int32
    L.kdbo_dot_i8.argtypes = [vp, vp, C.c_size_t]
    L.kdbo_int8_norm.restype = C.c_float
    L.kdbo_int8_norm.argtypes = [vp, C.c_size_t]
    L.kdbo_int8_cosine_distance.restype = C.c_double
    L.kdbo_int8_cosine_distance.argtypes = [C.c_int32, C.c_float, C.c_float]
    L.kdbo_quantize.argtypes = [C.c_float, vp, vp, C.c_size_t]
    L.kdbo_train_quantizer.restype = C.c_float
    L.kdbo_train_quantizer.argtypes = [vp, C.c_size_t, C.c_size_t]
    L.kdbo_row_raw.restype = vp
    L.kdbo_row_raw.argtypes = [vp, C.c_uint32]
    L.kdbo_row_raw_stride.restype = C.c_size_t
    L.kdbo_row_raw_stride.argtypes = [vp]
    L.kdbo_norms.restype = C.POINTER(C.c_float)
    L.kdbo_norms.argtypes = [vp]
    L.kdbo_set_arith.argtypes = [vp, C.c_int]
    L.kdbo_add.restype = C.c_uint32
    L.kdbo_add.argtypes = [vp, vp, C.c_double]
    L.kdbo_add_many.restype = C.c_int
    L.kdbo_add_many.argtypes = [vp, vp, C.c_size_t, vp, C.c_int]
    L.kdbo_add_batch.restype = C.c_int
    L.kdbo_add_batch.argtypes = [vp, vp, C.c_size_t, vp, C.c_int, C.c_int]
    L.kdbo_delete.argtypes = [vp, C.c_uint32]
    L.kdbo_search.restype = C.c_int
    L.kdbo_search.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, C.POINTER(_Stats)]
    L.kdbo_search_batch.restype = C.c_int
    L.kdbo_search_batch.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, vp,
                                    C.POINTER(_Stats), C.c_int]
    L.kdbo_search_layer.restype = C.c_int
    L.kdbo_search_layer.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_size_t, C.c_int, vp, vp,
                                    C.POINTER(_Stats)]
    L.kdbo_flat_search_batch.restype = C.c_int
    L.kdbo_flat_search_batch.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, vp, C.c_int]
    for name in ("kdbo_distance", "kdbo_distance_generic"):
        f = getattr(L, name)
        f.restype = C.c_double
        f.argtypes = [C.c_int, C.c_int, vp, vp, C.c_size_t]
    for name in ("kdbo_sq_euclid_f32", "kdbo_dot_f32"):
        f = getattr(L, name)
        f.restype = C.c_float
        f.argtypes = [C.c_int, vp, vp, C.c_size_t]
    L.kdbo_normalize.argtypes = [vp, C.c_size_t]
    L.kdbo_random_level.restype = C.c_int
    L.kdbo_random_level.argtypes = [C.c_double, C.c_int, C.c_int]
    L.kdbo_score_from_distance.restype = C.c_double
    L.kdbo_score_from_distance.argtypes = [C.c_double]
    L.kdbo_effective_ef.restype = C.c_int
    L.kdbo_effective_ef.argtypes = [C.c_int, C.c_int]
    L.kdbo_heap_roundtrip.argtypes = [C.c_int, vp, vp, C.c_size_t, vp, vp]
    L.kdbo_select_neighbors.restype = C.c_int
    L.kdbo_select_neighbors.argtypes = [vp, vp, vp, C.c_size_t, C.c_int, vp]
    L.kdbo_count.restype = C.c_uint32
    L.kdbo_count.argtypes = [vp]
    L.kdbo_entry.restype = C.c_uint32
    L.kdbo_entry.argtypes = [vp]
    L.kdbo_max_level.restype = C.c_int
    L.kdbo_max_level.argtypes = [vp]
    L.kdbo_vector.restype = C.POINTER(C.c_float)
    L.kdbo_vector.argtypes = [vp, C.c_uint32]
    L.kdbo_row_stride.restype = C.c_size_t
    L.kdbo_row_stride.argtypes = [vp]
    L.kdbo_export_sizes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kdbo_export_graph.argtypes = [vp, vp, vp, vp, vp, vp]
    L.kdbo_import_graph.restype = C.c_int
    L.kdbo_import_graph.argtypes = [vp, C.c_uint32, vp, C.c_size_t, vp, vp, vp, vp, vp, C.c_uint32, C.c_int]
    _ = (u32p, f32p, f64p)
    _lib = L
    return L


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def distance(metric: int, arith: int, a, b, generic: bool = False) -> float:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape and a.ndim == 1
    fn = lib().kdbo_distance_generic if generic else lib().kdbo_distance
    return float(fn(metric, arith, _p(a), _p(b), a.size))


def sq_euclid_f32(arith: int, a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().kdbo_sq_euclid_f32(arith, _p(a), _p(b), a.size))


def dot_f32(arith: int, a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().kdbo_dot_f32(arith, _p(a), _p(b), a.size))


def normalize(v) -> np.ndarray:
    out = _f32(v).copy()
    lib().kdbo_normalize(_p(out), out.size)
    return out


def normalize_rows(m) -> np.ndarray:
    out = _f32(m).copy()
    L = lib()
    for i in range(out.shape[0]):
        L.kdbo_normalize(C.c_void_p(out.ctypes.data + i * out.strides[0]), out.shape[1])
    return out


def f32_to_f16_bits(v) -> np.ndarray:
    """float16.Fromfloat32(x).Bits() elementwise (IEEE round-to-nearest-even)."""
    a = _f32(v)
    L = lib()
    return np.array([L.kdbo_f32_to_f16(float(x)) for x in a.ravel()], dtype=np.uint16).reshape(a.shape)


def f16_bits_to_f32(v) -> np.ndarray:
    a = np.ascontiguousarray(v, dtype=np.uint16)
    L = lib()
    return np.array([L.kdbo_f16_to_f32(int(x)) for x in a.ravel()], dtype=np.float32).reshape(a.shape)


def sq_euclid_f16(arith: int, a, b) -> np.float32:
    a, b = np.ascontiguousarray(a, np.uint16), np.ascontiguousarray(b, np.uint16)
    return np.float32(lib().kdbo_sq_euclid_f16(arith, _p(a), _p(b), a.size))


def dot_i8(a, b) -> int:
    a, b = np.ascontiguousarray(a, np.int8), np.ascontiguousarray(b, np.int8)
    return int(lib().kdbo_dot_i8(_p(a), _p(b), a.size))


def int8_norm(v) -> np.float32:
    a = np.ascontiguousarray(v, np.int8)
    return np.float32(lib().kdbo_int8_norm(_p(a), a.size))


def int8_cosine_distance(dot: int, qnorm: float, stored_norm: float) -> float:
    return float(lib().kdbo_int8_cosine_distance(int(dot), float(qnorm), float(stored_norm)))


def quantize(abs_max: float, v) -> np.ndarray:
    """Quantizer.Quantize (pkg/core/distance/quantizer.go:135-160), row-wise for 2-D input."""
    a = _f32(v)
    out = np.zeros(a.shape, dtype=np.int8)
    lib().kdbo_quantize(float(abs_max), _p(a), _p(out), a.size)
    return out


def train_quantizer(vecs) -> np.float32:
    """Quantizer.Train (quantizer.go:49-125): AbsMax = 99.9th percentile of |value| over a stride sample."""
    a = _f32(vecs)
    assert a.ndim == 2
    return np.float32(lib().kdbo_train_quantizer(_p(a), a.shape[0], a.shape[1]))


def random_level(u: float, m: int, current_max: int) -> int:
    return int(lib().kdbo_random_level(u, m, current_max))


def score_from_distance(d: float) -> float:
    return float(lib().kdbo_score_from_distance(d))


def effective_ef(ef_search: int, needs_refine: bool) -> int:
    return int(lib().kdbo_effective_ef(ef_search, int(needs_refine)))


def heap_roundtrip(kind: str, ids, dists):
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    d = np.ascontiguousarray(dists, dtype=np.float64)
    oi, od = np.empty_like(ids), np.empty_like(d)
    lib().kdbo_heap_roundtrip(0 if kind == "min" else 1, _p(ids), _p(d), ids.size, _p(oi), _p(od))
    return oi, od


def dense_bitset(ids, n_bits: int) -> np.ndarray:
    """Dense uint64 bitset (bit i of word i//64) over internal ids — the form the allow-list
    takes at the C boundary (same membership as the reference's roaring bitmap)."""
    words = np.zeros((n_bits >> 6) + 1, dtype=np.uint64)
    ids = np.asarray(ids, dtype=np.uint64)
    if ids.size:
        np.bitwise_or.at(words, (ids >> np.uint64(6)).astype(np.int64), np.uint64(1) << (ids & np.uint64(63)))
    return words


@dataclass
class Graph:
    """Flattened graph exchange format (see kdb_oracle.h)."""
    n: int
    levels: np.ndarray     # int32 [n+1], -1 = nil
    node_row: np.ndarray   # uint64 [n+2]
    row_off: np.ndarray    # uint64 [rows+1]
    nbrs: np.ndarray       # uint32 [edges]
    deleted: np.ndarray    # uint8 [n+1]
    entry: int
    max_level: int

    def save(self, path: str) -> None:
        np.savez(path, n=self.n, levels=self.levels, node_row=self.node_row, row_off=self.row_off,
                 nbrs=self.nbrs, deleted=self.deleted, entry=self.entry, max_level=self.max_level)

    @staticmethod
    def load(path: str) -> "Graph":
        z = np.load(path)
        return Graph(int(z["n"]), z["levels"], z["node_row"], z["row_off"], z["nbrs"], z["deleted"],
                     int(z["entry"]), int(z["max_level"]))

    def row(self, node: int, level: int) -> np.ndarray:
        r = int(self.node_row[node]) + level
        return self.nbrs[int(self.row_off[r]):int(self.row_off[r + 1])]


class OracleIndex:
    """Mirror of hnsw.Index (reference pkg/core/hnsw/hnsw_index.go) restricted to the hot path."""

    def __init__(self, dim: int, metric: int, m: int = 16, ef_construction: int = 200,
                 arith: int = ARITH_SEQ, capacity: int = 1 << 16, precision: int = PREC_F32):
        self._L = lib()
        self._h = self._L.kdbo_new_ex(dim, metric, precision, m, ef_construction, arith, capacity)
        if not self._h:
            raise ValueError("kdbo_new_ex failed (unsupported metric/precision pair, or out of memory)")
        self.dim, self.metric, self.capacity, self.precision = dim, metric, capacity, precision
        self.m = m if m > 0 else 16
        self.ef_construction = ef_construction if ef_construction > 0 else 200
        self.needs_refine = False

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.kdbo_free(self._h)
            self._h = None

    # -- build ---------------------------------------------------------------------------
    def set_arith(self, arith: int) -> None:
        self._L.kdbo_set_arith(self._h, arith)

    def set_quantizer(self, abs_max: float) -> None:
        self._L.kdbo_set_quantizer(self._h, float(abs_max))

    @property
    def abs_max(self) -> float:
        return float(self._L.kdbo_abs_max(self._h))

    def add(self, vec, u: float) -> int:
        v = _f32(vec)
        assert v.shape == (self.dim,)
        return int(self._L.kdbo_add(self._h, _p(v), float(u)))

    def add_many(self, vecs, u, threads: int = 1) -> None:
        v = _f32(vecs)
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert v.ndim == 2 and v.shape[1] == self.dim and u.shape == (v.shape[0],)
        rc = self._L.kdbo_add_many(self._h, _p(v), v.shape[0], _p(u), threads)
        if rc != 0:
            raise RuntimeError("kdbo_add_many failed (capacity?)")

    def add_batch(self, vecs, u, ef_const: int = 0, threads: int = 1) -> None:
        """AddBatch (hnsw_index.go:1466) — ef_const=0 uses efConstruction; AddBatchFast = max(2M, 40)."""
        v = _f32(vecs)
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert v.ndim == 2 and v.shape[1] == self.dim and u.shape == (v.shape[0],)
        rc = self._L.kdbo_add_batch(self._h, _p(v), v.shape[0], _p(u), ef_const, threads)
        if rc != 0:
            raise RuntimeError("kdbo_add_batch failed (capacity?)")

    def build_batched(self, vecs, u, batch: int = 10000, ef_const: int = 0, threads: int = 1) -> None:
        """The reference benchmark's ingestion: vadd_batch in chunks (clients/python/benchmark2.py:199-205)."""
        for i in range(0, len(vecs), batch):
            self.add_batch(vecs[i:i + batch], u[i:i + batch], ef_const, threads)

    def delete(self, internal_id: int) -> None:
        self._L.kdbo_delete(self._h, internal_id)

    # -- search --------------------------------------------------------------------------
    def search(self, query, k: int, ef_search: int = 0, allow: np.ndarray | None = None, with_stats=False):
        q = _f32(query)
        ids = np.zeros(max(k, 1), dtype=np.uint32)
        sc = np.zeros(max(k, 1), dtype=np.float64)
        st = _Stats()
        n = self._L.kdbo_search(self._h, _p(q), k, ef_search, int(self.needs_refine), _p(allow),
                                0 if allow is None else allow.size, _p(ids), _p(sc), C.byref(st))
        res = (ids[:n].copy(), sc[:n].copy())
        return (*res, Stats(st.dist_evals, st.hops, st.hops_l0)) if with_stats else res

    def search_batch(self, queries, k: int, ef_search: int = 0, allow: np.ndarray | None = None,
                     threads: int = 1):
        q = _f32(queries)
        nq = q.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        sc = np.zeros((nq, k), dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        st = _Stats()
        self._L.kdbo_search_batch(self._h, _p(q), nq, k, ef_search, int(self.needs_refine), _p(allow),
                                  0 if allow is None else allow.size, _p(ids), _p(sc), _p(cnt), C.byref(st),
                                  threads)
        return ids, sc, cnt, Stats(st.dist_evals, st.hops, st.hops_l0)

    def search_layer(self, prepared_query, entry: int, k: int, level: int, ef_search: int,
                     allow: np.ndarray | None = None):
        q = np.ascontiguousarray(prepared_query, dtype=_RAW_DTYPE[self.precision])
        cap = max(k, ef_search, 1)
        ids = np.zeros(cap, dtype=np.uint32)
        sc = np.zeros(cap, dtype=np.float64)
        st = _Stats()
        n = self._L.kdbo_search_layer(self._h, _p(q), entry, k, level, _p(allow),
                                      0 if allow is None else allow.size, ef_search, _p(ids), _p(sc),
                                      C.byref(st))
        if n < 0:
            raise KeyError(f"entry point node {entry} not found")
        return ids[:n].copy(), sc[:n].copy()

    def flat_search_batch(self, queries, k: int, mode: int = 1, allow: np.ndarray | None = None,
                          threads: int = 1):
        q = _f32(queries)
        nq = q.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        sc = np.zeros((nq, k), dtype=np.float64)
        cnt = np.zeros(nq, dtype=np.int32)
        self._L.kdbo_flat_search_batch(self._h, _p(q), nq, k, mode, _p(allow),
                                       0 if allow is None else allow.size, _p(ids), _p(sc), _p(cnt), threads)
        return ids, sc, cnt

    def select_neighbors(self, ids, dists, m: int) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        d = np.ascontiguousarray(dists, dtype=np.float64)
        out = np.zeros(max(ids.size, 1), dtype=np.uint32)
        n = self._L.kdbo_select_neighbors(self._h, _p(ids), _p(d), ids.size, m, _p(out))
        return out[:n].copy()

    # -- introspection -------------------------------------------------------------------
    @property
    def count(self) -> int:
        return int(self._L.kdbo_count(self._h))

    @property
    def entry(self) -> int:
        return int(self._L.kdbo_entry(self._h))

    @property
    def max_level(self) -> int:
        return int(self._L.kdbo_max_level(self._h))

    def vectors(self) -> np.ndarray:
        """Stored rows [count+1, dim] (row 0 is the nil slot), exactly as the reference keeps them
        (unit-normalised for cosine)."""
        n = self.count
        stride = int(self._L.kdbo_row_stride(self._h))
        base = self._L.kdbo_vector(self._h, 0)
        arr = np.ctypeslib.as_array(base, shape=((n + 1) * stride,)).reshape(n + 1, stride)
        return np.ascontiguousarray(arr[:, : self.dim])

    def rows_raw(self) -> np.ndarray:
        """Stored rows [count+1, dim] in the index precision: float32, float16 bits (uint16) or int8."""
        n = self.count
        stride = int(self._L.kdbo_row_raw_stride(self._h))
        dt = np.dtype(_RAW_DTYPE[self.precision])
        base = self._L.kdbo_row_raw(self._h, 0)
        buf = (C.c_char * ((n + 1) * stride * dt.itemsize)).from_address(base)
        arr = np.frombuffer(buf, dtype=dt).reshape(n + 1, stride)
        return np.ascontiguousarray(arr[:, : self.dim])

    def norms(self) -> np.ndarray:
        """int8 indexes: quantizedNorms[0..count] (computeInt8Norm of every stored row)."""
        assert self.precision == PREC_I8
        return np.ctypeslib.as_array(self._L.kdbo_norms(self._h), shape=(self.count + 1,)).copy()

    def export_graph(self) -> Graph:
        n = self.count
        rows, edges = C.c_uint64(), C.c_uint64()
        self._L.kdbo_export_sizes(self._h, C.byref(rows), C.byref(edges))
        levels = np.zeros(n + 1, dtype=np.int32)
        node_row = np.zeros(n + 2, dtype=np.uint64)
        row_off = np.zeros(rows.value + 1, dtype=np.uint64)
        nbrs = np.zeros(max(edges.value, 1), dtype=np.uint32)
        deleted = np.zeros(n + 1, dtype=np.uint8)
        self._L.kdbo_export_graph(self._h, _p(levels), _p(node_row), _p(row_off), _p(nbrs), _p(deleted))
        return Graph(n, levels, node_row, row_off, nbrs[: edges.value], deleted, self.entry, self.max_level)

    def import_graph(self, vectors: np.ndarray, g: Graph) -> None:
        """vectors: [n+1, dim] stored rows (row 0 unused) in the index precision."""
        v = np.ascontiguousarray(vectors, dtype=_RAW_DTYPE[self.precision])
        assert v.shape == (g.n + 1, self.dim)
        nbrs = g.nbrs if g.nbrs.size else np.zeros(1, dtype=np.uint32)
        rc = self._L.kdbo_import_graph(self._h, g.n, _p(v), self.dim, _p(np.ascontiguousarray(g.levels, np.int32)),
                                       _p(np.ascontiguousarray(g.node_row, np.uint64)),
                                       _p(np.ascontiguousarray(g.row_off, np.uint64)),
                                       _p(np.ascontiguousarray(nbrs, np.uint32)),
                                       _p(np.ascontiguousarray(g.deleted, np.uint8)), g.entry, g.max_level)
        if rc != 0:
            raise RuntimeError(f"kdbo_import_graph failed rc={rc}")
