"""ctypes binding of include/kektordb_gpu.h — the same C ABI the Go host binds through cgo.

Loading fails loudly if the library has not been built; calling fails loudly (GpuError) when
there is no CUDA device.  There is no CPU fallback anywhere behind this module.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

LIB_PATH = _build.LIB

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_OVERFLOW = -1, -2, -3, -4, -5
METRIC_L2, METRIC_COSINE = 0, 1
PRECISION_F32, PRECISION_F16, PRECISION_INT8 = 0, 1, 2
FLAT_PREFILTER = 0x10


class GpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kektordb_gpu error {code}: {msg}")
        self.code = code


class BatcherStats(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("batches", C.c_uint64), ("max_batch_seen", C.c_uint64),
                ("dispatched_idle", C.c_uint64), ("dispatched_full", C.c_uint64), ("dispatched_deadline", C.c_uint64)]


# kdbgpu_batch_fn: the batch executor a micro-batcher fronts
BATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint64),
                       C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_uint32))


class Stats(C.Structure):
    _fields_ = [("dist_evals", C.c_uint64), ("hops", C.c_uint64), ("hops_l0", C.c_uint64),
                ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("heap_pass_queries", C.c_uint32)]


class RefresherStats(C.Structure):
    _fields_ = [("pending_rows", C.c_uint64), ("pending_nodes", C.c_uint64), ("flushes", C.c_uint64),
                ("flushes_by_rows", C.c_uint64), ("flushes_by_lag", C.c_uint64), ("flushes_by_call", C.c_uint64),
                ("rows_queued", C.c_uint64), ("rows_applied", C.c_uint64), ("nodes_applied", C.c_uint64),
                ("oldest_pending_ms", C.c_float), ("last_flush_ms", C.c_float), ("last_error", C.c_int)]


class ShardStats(C.Structure):
    _fields_ = [("dist_evals", C.c_uint64), ("hops", C.c_uint64), ("hops_l0", C.c_uint64),
                ("traversal_ms", C.c_float), ("exchange_ms", C.c_float), ("merge_ms", C.c_float),
                ("total_ms", C.c_float), ("n_shards", C.c_uint32)]


SHARD_ID_BYTES = 128

# every symbol include/kektordb_gpu.h declares: name -> (restype, argtypes)
_vp, _u32, _i32, _sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
SIGNATURES = {
    "kdbgpu_device_count": (_i32, []),
    "kdbgpu_last_error": (C.c_char_p, []),
    "kdbgpu_version": (C.c_char_p, []),
    "kdbgpu_index_create": (_i32, [_i32, _i32, _i32, _i32, _u32, C.POINTER(_vp)]),
    "kdbgpu_index_create_ex": (_i32, [_i32, _i32, _i32, _i32, _i32, _u32, C.POINTER(_vp)]),
    "kdbgpu_index_destroy": (_i32, [_vp]),
    "kdbgpu_index_precision": (_i32, [_vp]),
    "kdbgpu_set_quantizer": (_i32, [_vp, C.c_float]),
    "kdbgpu_train_quantizer": (_i32, [_vp, _vp, _u32, C.POINTER(C.c_float)]),
    "kdbgpu_train_quantizer_device": (_i32, [_vp, _vp, _sz, _u32, C.POINTER(C.c_float)]),
    "kdbgpu_arena_probe": (_i32, [C.c_char_p, C.POINTER(_u32), C.POINTER(_i32), C.POINTER(_u32), C.POINTER(_u32)]),
    "kdbgpu_arena_load_dir": (_i32, [_vp, C.c_char_p, _vp, _u32, C.POINTER(C.c_uint64)]),
    "kdbgpu_arena_stage_chunk": (_i32, [_vp, _u32, _vp, _sz, _vp, _u32, C.POINTER(_u32)]),
    "kdbgpu_arena_chunks_registered": (C.c_uint64, [_vp]),
    "kdbgpu_upload_rows_raw": (_i32, [_vp, _u32, _u32, _vp]),
    "kdbgpu_download_rows_raw": (_i32, [_vp, _u32, _u32, _vp]),
    "kdbgpu_download_norms": (_i32, [_vp, _u32, _u32, _vp]),
    "kdbgpu_upload_vectors": (_i32, [_vp, _u32, _u32, _vp]),
    "kdbgpu_upload_vectors_device": (_i32, [_vp, _u32, _u32, _vp, _sz]),
    "kdbgpu_set_graph": (_i32, [_vp, _u32, _vp, _vp, _vp, _vp, _u32, _i32]),
    "kdbgpu_graph_file_write": (_i32, [C.c_char_p, _u32, _i32, _vp, _vp, _vp, _vp, _u32, _i32]),
    "kdbgpu_save_graph_file": (_i32, [_vp, C.c_char_p]),
    "kdbgpu_graph_file_probe": (_i32, [C.c_char_p, C.POINTER(_u32), C.POINTER(_i32), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.POINTER(_u32), C.POINTER(_i32)]),
    "kdbgpu_set_graph_file": (_i32, [_vp, C.c_char_p]),
    "kdbgpu_index_m": (_i32, [_vp]),
    "kdbgpu_refresher_create": (_i32, [_vp, _u32, _u32, C.POINTER(_vp)]),
    "kdbgpu_refresher_destroy": (_i32, [_vp]),
    "kdbgpu_refresher_add_node": (_i32, [_vp, _u32, _i32, _vp]),
    "kdbgpu_refresher_set_row": (_i32, [_vp, _u32, _i32, _vp, _u32]),
    "kdbgpu_refresher_remove_node": (_i32, [_vp, _u32]),
    "kdbgpu_refresher_set_deleted": (_i32, [_vp, _u32, _i32]),
    "kdbgpu_refresher_set_entry": (_i32, [_vp, _u32, _i32]),
    "kdbgpu_refresher_flush": (_i32, [_vp]),
    "kdbgpu_refresher_stats": (_i32, [_vp, C.POINTER(RefresherStats)]),
    "kdbgpu_register_nodes": (_i32, [_vp, _u32, _u32, _vp]),
    "kdbgpu_patch_rows": (_i32, [_vp, _u32, _vp, _vp, _vp, _vp]),
    "kdbgpu_remove_nodes": (_i32, [_vp, _u32, _vp]),
    "kdbgpu_set_entry": (_i32, [_vp, _u32, _i32]),
    "kdbgpu_set_deleted": (_i32, [_vp, _vp, _sz]),
    "kdbgpu_search_batch": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _sz, _vp, _vp, _vp, C.POINTER(Stats)]),
    "kdbgpu_search_batch_device": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _sz, _u32, _vp, _vp, _vp, _vp]),
    "kdbgpu_distance_batch": (_i32, [_vp, _vp, _vp, _u32, _vp]),
    "kdbgpu_flat_search_batch": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _sz, _vp, _vp, _vp, C.POINTER(Stats)]),
    "kdbgpu_flat_prefilter_scores": (_i32, [_vp, _vp, _u32, _i32, _vp, _vp]),
    "kdbgpu_merge_topk_device": (_i32, [_vp, _i32, _u32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kdbgpu_add_batch": (_i32, [_vp, _u32, _vp, _vp, _i32]),
    "kdbgpu_add_batch_device": (_i32, [_vp, _u32, _vp, _sz, _vp, _i32]),
    "kdbgpu_get_graph_sizes": (_i32, [_vp, C.POINTER(_u32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(_u32), C.POINTER(_i32)]),
    "kdbgpu_get_graph": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "kdbgpu_download_vectors": (_i32, [_vp, _u32, _u32, _vp]),
    "kdbgpu_last_search_stats": (_i32, [_vp, C.POINTER(Stats)]),
    "kdbgpu_index_device": (_i32, [_vp]),
    "kdbgpu_index_dim": (_i32, [_vp]),
    "kdbgpu_batcher_create": (_i32, [_vp, _u32, _u32, C.POINTER(_vp)]),
    "kdbgpu_batcher_create_fn": (_i32, [BATCH_FN, _vp, _i32, _u32, _u32, C.POINTER(_vp)]),
    "kdbgpu_batcher_destroy": (_i32, [_vp]),
    "kdbgpu_batcher_search": (_i32, [_vp, _vp, _i32, _i32, _vp, _sz, _vp, _vp, C.POINTER(_u32)]),
    "kdbgpu_batcher_stats": (_i32, [_vp, C.POINTER(BatcherStats)]),
    "kdbgpu_batcher_create_group": (_i32, [_vp, _i32, _u32, _u32, C.POINTER(_vp)]),
    "kdbgpu_batcher_submit": (_i32, [_vp, _vp, _i32, _i32, _vp, _sz, C.c_uint64, C.POINTER(C.c_uint64)]),
    "kdbgpu_batcher_poll": (_i32, [_vp, _vp, _u32, _u32, C.POINTER(_u32)]),
    "kdbgpu_batcher_take": (_i32, [_vp, C.c_uint64, _vp, _vp, C.POINTER(_u32)]),
    "kdbgpu_batcher_register_filter": (_i32, [_vp, _vp, _sz, C.POINTER(C.c_uint64)]),
    "kdbgpu_batcher_release_filter": (_i32, [_vp, C.c_uint64]),
    "kdbgpu_host_alloc": (_i32, [C.POINTER(_vp), _sz]),
    "kdbgpu_host_free": (None, [_vp]),
    "kdbgpu_index_count": (_u32, [_vp]),
    "kdbgpu_index_device_bytes": (C.c_uint64, [_vp]),
    "kdbgpu_search_concurrency": (_i32, [_vp, _i32, _i32]),
    "kdbgpu_prepare_search": (_i32, [_vp, _u32, _i32, _i32]),
    "kdbgpu_set_fast_path": (_i32, [_vp, _i32]),
    "kdbgpu_set_tuning": (_i32, [_vp, _i32, _i32, _i32]),
    "kdbgpu_set_idle_slots": (_i32, [_vp, _i32]),
    "kdbgpu_set_candidate_bound": (_i32, [_vp, _u32]),
    "kdbgpu_shard_unique_id": (_i32, [_vp]),
    "kdbgpu_shard_group_create_rank": (_i32, [_vp, _i32, _i32, _vp, _u32, C.POINTER(_vp)]),
    "kdbgpu_shard_group_create_local": (_i32, [C.POINTER(_vp), _i32, _vp, C.POINTER(_vp)]),
    "kdbgpu_shard_group_destroy": (_i32, [_vp]),
    "kdbgpu_shard_group_size": (_i32, [_vp]),
    "kdbgpu_shard_search_batch": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _sz, _vp, _vp, _vp, C.POINTER(ShardStats)]),
    "kdbgpu_shard_search_submit": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _sz, C.POINTER(_vp)]),
    "kdbgpu_shard_search_wait": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(ShardStats)]),
    "kdbgpu_shard_search_batch_device": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "kdbgpu_shard_sync": (_i32, [_vp, C.POINTER(ShardStats)]),
    "kdbgpu_shard_flat_search_batch": (_i32, [_vp, _vp, _u32, _i32, _i32, _vp, _sz, _vp, _vp, _vp, C.POINTER(ShardStats)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libkektordb_gpu.so (built in-tree by kektordb_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -m kektordb_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise GpuError(rc, lib().kdbgpu_last_error().decode("utf-8", "replace"))
