"""N>1 host logic on CPU: two gloo ranks, each owning an id-range shard with its own
reference-semantic index (the oracle stands in for the per-GPU search), one all_gather of the
per-shard top-k and a merge by (distance, id) — compared with the single-process answer
"G independent indexes + exact merge" (SURVEY.md §8e)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_corpus():
    from kektordb_b200.sharding import shard_range
    for n in (1, 7, 10, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_globalize_ids_and_reference_merge():
    from kektordb_b200.sharding import globalize_ids
    from tests.pyref import merge_reference
    ids = np.array([[3, 1, 0], [2, 0, 0]], dtype=np.uint32)
    g = globalize_ids(ids, np.array([2, 1]), 100)
    assert g.tolist() == [[103, 101, 0], [102, 0, 0]]
    S = np.array([[[0.1, 0.5], [0.2, 0.2]], [[0.1, 0.3], [0.2, 0.9]]])
    I = np.array([[[5, 6], [9, 7]], [[4, 8], [3, 2]]], dtype=np.uint32)
    C = np.array([[2, 2], [2, 1]])
    oi, os_, oc = merge_reference(I, S, C, 3)
    assert oi.tolist() == [[4, 5, 8], [3, 7, 9]]  # ties on distance resolved by id
    assert oc.tolist() == [3, 3]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from kektordb_b200.sharding import globalize_ids, shard_range
    from tests.pyref import merge_reference
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, dim, k, ef = 3000, 24, 10, 48
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    X[rng.integers(0, n, 200)] = X[rng.integers(0, n, 200)]  # cross-shard duplicates: distance ties
    Q = rng.standard_normal((40, dim)).astype(np.float32)
    base, count = shard_range(n, world, rank)
    idx = O.OracleIndex(dim, O.METRIC_COSINE, 8, 60, O.ARITH_KERNEL, count)
    idx.build_batched(X[base:base + count], np.random.default_rng(100 + rank).random(count), batch=256)
    ids, sc, cnt, _ = idx.search_batch(Q, k, ef)
    gids = globalize_ids(ids, cnt, base)
    g_ids = torch.zeros((world, len(Q), k), dtype=torch.int32)
    g_sc = torch.zeros((world, len(Q), k), dtype=torch.float64)
    g_cnt = torch.zeros((world, len(Q)), dtype=torch.int32)
    # concatenated-along-dim-0 form: accepted by gloo and NCCL alike; [S*Q, k] is the same memory as [S][Q][k]
    dist.all_gather_into_tensor(g_ids.view(world * len(Q), k), torch.from_numpy(gids.astype(np.int32)))
    dist.all_gather_into_tensor(g_sc.view(world * len(Q), k), torch.from_numpy(sc))
    dist.all_gather_into_tensor(g_cnt.view(world * len(Q)), torch.from_numpy(cnt.astype(np.int32)))
    m_ids, m_sc, m_cnt = merge_reference(g_ids.numpy().astype(np.uint32), g_sc.numpy(), g_cnt.numpy(), k)
    if rank == 0:
        np.savez(out_path, ids=m_ids, sc=m_sc, cnt=m_cnt)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from kektordb_b200.sharding import globalize_ids, shard_range
    from tests.pyref import merge_reference
    from oracle import oracle as O
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    # single-process restatement: G reference-semantic indexes + exact merge by (distance, id)
    n, dim, k, ef, world = 3000, 24, 10, 48, 2
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    X[rng.integers(0, n, 200)] = X[rng.integers(0, n, 200)]
    Q = rng.standard_normal((40, dim)).astype(np.float32)
    all_ids, all_sc, all_cnt = [], [], []
    for r in range(world):
        base, count = shard_range(n, world, r)
        idx = O.OracleIndex(dim, O.METRIC_COSINE, 8, 60, O.ARITH_KERNEL, count)
        idx.build_batched(X[base:base + count], np.random.default_rng(100 + r).random(count), batch=256)
        ids, sc, cnt, _ = idx.search_batch(Q, k, ef)
        all_ids.append(globalize_ids(ids, cnt, base))
        all_sc.append(sc)
        all_cnt.append(cnt)
    w_ids, w_sc, w_cnt = merge_reference(np.stack(all_ids), np.stack(all_sc), np.stack(all_cnt), k)
    assert np.array_equal(got["ids"], w_ids) and np.array_equal(got["sc"], w_sc) and np.array_equal(got["cnt"], w_cnt)
    # and the merged answer is a good one: recall against the exact scan of the whole corpus
    full = O.OracleIndex(dim, O.METRIC_COSINE, 8, 60, O.ARITH_KERNEL, n)
    full.build_batched(X, np.random.default_rng(1).random(n), batch=256)
    gt, _, _ = full.flat_search_batch(Q, k, mode=1)
    rec = np.mean([len(set(w_ids[i]) & set(gt[i])) / k for i in range(len(Q))])
    assert rec >= 0.9
