"""Host-side plumbing for id-range shards: which ids a rank owns, how local ids map
to global ids, and the [shards][Q][k] layout the merge kernel consumes.  Pure functions — the
compute (per-shard search, merge) stays in libkektordb_gpu."""
from __future__ import annotations

import numpy as np


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


def merge_reference(ids: np.ndarray, scores: np.ndarray, counts: np.ndarray, k: int):
    """What kdbgpu_merge_topk_device computes, in numpy: per query the k smallest of the union by
    (distance, id).  ids/scores [S][Q][k], counts [S][Q]."""
    S, Q, kk = ids.shape
    out_ids = np.zeros((Q, k), dtype=np.uint32)
    out_sc = np.zeros((Q, k), dtype=np.float64)
    out_cnt = np.zeros(Q, dtype=np.uint32)
    for q in range(Q):
        pool = [(float(scores[s, q, i]), int(ids[s, q, i])) for s in range(S) for i in range(min(int(counts[s, q]), kk))]
        pool.sort()
        pool = pool[:k]
        out_cnt[q] = len(pool)
        for i, (d, idx) in enumerate(pool):
            out_ids[q, i], out_sc[q, i] = idx, d
    return out_ids, out_sc, out_cnt
