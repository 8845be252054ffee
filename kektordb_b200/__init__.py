"""kektordb_b200 — B200 (sm_100a) vector-search hot path for KektorDB behind a C ABI.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/kektordb_gpu.h), `ffi.py` (ctypes binding of that ABI), `index.py` (host-side mirror
of hnsw.Index's search surface), `batcher.py` (one-query-per-call micro-batcher over the same ABI), `sharding.py` (id-range shard groups) and `build.py` (nvcc build, in-tree).
"""
from .index import Cosine, Euclidean, Float16, Float32, Int8, GpuIndex, SearchStats, arena_probe, dense_allow_list, effective_ef  # noqa: F401

from .batcher import Batcher, BatcherStats  # noqa: E402,F401
from .sharding import ShardGroup, ShardStats, shard_range  # noqa: E402,F401
from .refresher import Refresher, RefresherStats  # noqa: E402,F401

__all__ = ["Batcher", "BatcherStats", "ShardGroup", "ShardStats", "shard_range", "Refresher", "RefresherStats", "GpuIndex", "SearchStats", "arena_probe", "Cosine", "Euclidean", "Float32", "Float16", "Int8", "dense_allow_list", "effective_ef"]
