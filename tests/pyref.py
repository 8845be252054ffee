"""A second, independent restatement (pure Python loops) of the reference's layer search, used
only to cross-check the C oracle on small cases.  Follows /root/reference
pkg/core/hnsw/hnsw_index.go:2351-2611 (searchLayerUnlocked), :369-468 (searchInternal) and
hnsw_heap.go:18-156.  Distances are supplied by the caller (a dict id -> float64) so the two
implementations are compared on the traversal logic alone."""
from __future__ import annotations


class _Heap:
    """hnsw_heap.go: swap-based up()/down() with strict comparisons."""

    def __init__(self, is_min: bool):
        self.a: list[tuple[int, float]] = []
        self.is_min = is_min

    def _before(self, x, y) -> bool:
        return x[1] < y[1] if self.is_min else x[1] > y[1]

    def __len__(self):
        return len(self.a)

    def peek(self):
        return self.a[0]

    def push(self, x):
        a = self.a
        a.append(x)
        j = len(a) - 1
        while True:
            i = (j - 1) // 2 if j > 0 else 0
            if i == j or not self._before(a[j], a[i]):
                break
            a[i], a[j] = a[j], a[i]
            j = i

    def pop(self):
        a = self.a
        x = a[0]
        a[0] = a[-1]
        a.pop()
        n = len(a)
        i = 0
        while True:
            j1 = 2 * i + 1
            if j1 >= n:
                break
            j = j1
            if j1 + 1 < n and self._before(a[j1 + 1], a[j1]):
                j = j1 + 1
            if not self._before(a[j], a[i]):
                break
            a[i], a[j] = a[j], a[i]
            i = j
        return x


def search_layer(dist, rows, node_levels, deleted, entry, k, level, allow, ef_search):
    """dist: callable id -> float; rows: callable (id, level) -> list of neighbour ids;
    node_levels: dict id -> level (absent = nil); allow: None or a set (empty set = inactive)."""
    if entry not in node_levels:
        raise KeyError(entry)
    ef = max(ef_search, k)
    allow_active = allow is not None and len(allow) > 0
    cands, results = _Heap(True), _Heap(False)
    visited = set()
    evals = 1
    ep = (entry, dist(entry))
    cands.push(ep)
    visited.add(entry)
    if (not allow_active or entry in allow) and entry not in deleted:
        results.push(ep)
    while len(cands) > 0:
        cur = cands.pop()
        if len(results) >= ef and cur[1] > results.peek()[1]:
            break
        if cur[0] not in node_levels or level > node_levels[cur[0]]:
            continue
        for nb in list(rows(cur[0], level)):
            if nb in visited:
                continue
            visited.add(nb)
            if allow_active and nb not in allow:
                continue
            if nb not in node_levels:
                continue
            d = dist(nb)
            evals += 1
            if len(results) < ef or d < results.peek()[1]:
                cands.push((nb, d))
                if nb not in deleted:
                    results.push((nb, d))
                    if len(results) > ef:
                        results.pop()
    out = [None] * len(results)
    for i in range(len(out) - 1, -1, -1):
        out[i] = results.pop()
    return out[:k], evals


def search(dist, rows, node_levels, deleted, entry, max_level, k, ef_search, allow):
    """searchInternal minus query preparation.  allow: None or a set of ids."""
    if max_level == -1:
        return []
    if allow is not None and entry not in allow:
        if not allow:
            return []
        entry = min(allow)
    for l in range(max_level, 0, -1):
        try:
            nearest, _ = search_layer(dist, rows, node_levels, deleted, entry, 1, l, allow, 0)
        except KeyError:
            return []
        if not nearest:
            return []
        entry = nearest[0][0]
    try:
        res, _ = search_layer(dist, rows, node_levels, deleted, entry, k, 0, allow, ef_search)
    except KeyError:
        return []
    return res


def merge_reference(ids, scores, counts, k):
    """What the merge kernels compute (kdbgpu_merge_topk_device, the shard group's merge), in numpy: per
    query the k smallest of the union by (distance, id).  ids/scores [S][Q][k], counts [S][Q]."""
    import numpy as np
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
