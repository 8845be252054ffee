"""GPU parity of the SHARDED path through the C ABI's shard group.

The reference has no sharded mode, so the oracle is "G reference-semantic indexes + exact merge by
(distance, id)": every shard is an oracle index over its id range (KDBO_ARITH_KERNEL), searched on the
CPU, ids shifted to global, merged in numpy.  The bar is bit-exact: ids, float64 scores, counts and the
summed E / H counters.

One GPU is enough for the exchange + merge logic (several shards on one device, local group: same-device
copies instead of NVLink peer copies); the two-GPU variants — peer copies between devices and the NCCL
rank group, one process per GPU — run when the box has two devices."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tests.pyref import merge_reference

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    from kektordb_b200 import ffi
    n = ffi.lib().kdbgpu_device_count()
    assert n > 0, "these tests need a CUDA device (no CPU fallback exists)"
    return n


def _make_corpus(n, dim, seed, dup=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    if dup:  # duplicates across shards: exact distance ties that the merge must order by id
        X[rng.integers(0, n, n // 8)] = X[rng.integers(0, n, n // 8)]
    Q = rng.standard_normal((48, dim)).astype(np.float32)
    Q[:8] = X[rng.integers(0, n, 8)]
    return X, Q


def _oracle_shards(X, world, metric, m, efc, seed):
    from kektordb_b200.sharding import shard_range
    out = []
    for r in range(world):
        base, count = shard_range(len(X), world, r)
        oi = O.OracleIndex(X.shape[1], metric, m, efc, O.ARITH_KERNEL, count + 4)
        if count:
            oi.build_batched(X[base:base + count], np.random.default_rng(seed + r).random(count), batch=256, threads=8)
        out.append((oi, base, count))
    return out


def _mirror(oi, metric, m, device=0):
    from kektordb_b200 import GpuIndex
    g = oi.export_graph()
    gi = GpuIndex(oi.dim, "cosine" if metric == O.METRIC_COSINE else "euclidean", m, max(g.n, 1), device=device)
    if g.n:
        gi.upload_vectors(1, oi.vectors()[1:])
    gi.set_graph(g.n, g.levels, g.node_row, g.row_off, g.nbrs, g.entry, g.max_level)
    return gi


def _local_allow(allow, base, count):
    """The slice of a global allow bitset a shard sees: local bit i = global bit base + i."""
    if allow is None:
        return None
    bits = np.unpackbits(allow.view(np.uint8), bitorder="little")
    loc = np.zeros(((count + 1 + 63) // 64) * 64, dtype=np.uint8)
    src = bits[base + 1: base + 1 + count]
    loc[1:1 + len(src)] = src
    return np.packbits(loc, bitorder="little").view(np.uint64)


def _oracle_answer(shards, Q, k, ef, allow=None):
    from kektordb_b200.sharding import globalize_ids
    ids, sc, cnt = [], [], []
    E = H = H0 = 0
    for oi, base, count in shards:
        la = _local_allow(allow, base, count)
        if count == 0 or (la is not None and not la.any()):
            i, s, c = np.zeros((len(Q), k), np.uint32), np.zeros((len(Q), k)), np.zeros(len(Q), np.uint32)
        else:
            i, s, c, st = oi.search_batch(Q, k, ef, allow=la, threads=8)
            E, H, H0 = E + st.dist_evals, H + st.hops, H0 + st.hops_l0
        ids.append(globalize_ids(i, c, base))
        sc.append(s)
        cnt.append(c)
    return merge_reference(np.stack(ids), np.stack(sc), np.stack(cnt), k) + ((E, H, H0),)


def _assert_same(got, want):
    gi, gs, gc = got[:3]
    wi, ws, wc = want[:3]
    assert np.array_equal(gc, wc.astype(np.uint32))
    assert np.array_equal(gi, wi)
    assert np.array_equal(gs, ws)  # float64 bit patterns


@pytest.mark.parametrize("world,metric,n,dim,k,ef", [
    (2, O.METRIC_COSINE, 3000, 24, 10, 48),
    (3, O.METRIC_L2, 2500, 40, 5, 32),
    (4, O.METRIC_COSINE, 1203, 128, 10, 64),
])
def test_sharded_search_is_bit_exact_vs_g_oracle_indexes_and_exact_merge(world, metric, n, dim, k, ef):
    from kektordb_b200.sharding import ShardGroup
    _gpu_count()
    X, Q = _make_corpus(n, dim, 5)
    shards = _oracle_shards(X, world, metric, 8, 60, 100)
    gis = [_mirror(oi, metric, 8) for oi, _, _ in shards]
    grp = ShardGroup.local(gis, [b for _, b, _ in shards])
    assert grp.size == world
    want = _oracle_answer(shards, Q, k, ef)
    got = grp.SearchWithScores(Q, k, None, ef)
    _assert_same(got, want)
    st = got[3]
    assert (st.dist_evals, st.hops, st.hops_l0) == want[3] and st.n_shards == world
    assert st.total_ms > 0 and st.traversal_ms > 0
    # every id lies in the corpus and is global
    assert got[0].max() <= n
    grp.close()
    for g in gis:
        g.close()


def test_sharded_allow_list_is_sliced_per_shard_and_an_empty_slice_contributes_nothing():
    from kektordb_b200.sharding import ShardGroup
    _gpu_count()
    n, dim, k, ef, world = 2400, 32, 10, 48, 3
    X, Q = _make_corpus(n, dim, 11)
    shards = _oracle_shards(X, world, O.METRIC_COSINE, 8, 60, 200)
    gis = [_mirror(oi, O.METRIC_COSINE, 8) for oi, _, _ in shards]
    grp = ShardGroup.local(gis, [b for _, b, _ in shards])
    rng = np.random.default_rng(3)
    # (a) 10 % of all ids; (b) members only in shards 0 and 2 (shard 1's slice is empty); (c) a single member
    lists = [np.where(rng.random(n) < 0.1)[0] + 1,
             np.concatenate([np.arange(1, 300), np.arange(1700, 2401)]),
             np.array([1234])]
    for members in lists:
        allow = O.dense_bitset(members, n)
        want = _oracle_answer(shards, Q, k, ef, allow)
        got = grp.SearchWithScores(Q, k, allow, ef)
        _assert_same(got, want)
        ms = set(members.tolist())
        assert all(int(i) in ms for i in got[0][got[0] > 0])
    # an allow-list with no member at all: every shard returns [] (hnsw_index.go:443-445)
    got = grp.SearchWithScores(Q, k, np.zeros(n // 64 + 1, np.uint64), ef)
    assert not got[2].any() and not got[0].any()
    grp.close()
    for g in gis:
        g.close()


def test_submit_wait_keeps_four_batches_in_flight_and_device_form_agrees():
    import torch
    from kektordb_b200.sharding import ShardGroup
    _gpu_count()
    n, dim, k, ef, world = 3000, 64, 10, 64, 2
    X, Q = _make_corpus(n, dim, 21)
    shards = _oracle_shards(X, world, O.METRIC_COSINE, 8, 60, 300)
    gis = [_mirror(oi, O.METRIC_COSINE, 8) for oi, _, _ in shards]
    grp = ShardGroup.local(gis, [b for _, b, _ in shards])
    rng = np.random.default_rng(9)
    batches = [np.ascontiguousarray(rng.standard_normal((32, dim)).astype(np.float32)) for _ in range(9)]
    want = [_oracle_answer(shards, b, k, ef) for b in batches]
    tickets, got = [], []
    for b in batches:  # at most 4 tickets outstanding
        if len(tickets) == 4:
            got.append(grp.wait(tickets.pop(0)))
        tickets.append(grp.submit(b, k, ef))
    while tickets:
        got.append(grp.wait(tickets.pop(0)))
    for g_, w_ in zip(got, want):
        _assert_same(g_, w_)
    # device-resident form, several batches queued back to back on one stream
    dev = torch.device("cuda", 0)
    outs = []
    stream = torch.cuda.Stream(device=dev)
    for b in batches:
        dq = torch.from_numpy(b).to(dev)
        di = torch.zeros((32, k), dtype=torch.int32, device=dev)
        ds = torch.zeros((32, k), dtype=torch.float64, device=dev)
        dc = torch.zeros(32, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        grp.search_device(dq.data_ptr(), 32, k, ef, di.data_ptr(), ds.data_ptr(), dc.data_ptr(), stream.cuda_stream)
        outs.append((dq, di, ds, dc))
    st = grp.sync()
    stream.synchronize()
    assert (st.dist_evals, st.hops, st.hops_l0) == want[-1][3]
    for (dq, di, ds, dc), w_ in zip(outs, want):
        _assert_same((di.cpu().numpy().astype(np.uint32), ds.cpu().numpy(), dc.cpu().numpy().astype(np.uint32)), w_)
    grp.close()
    for g in gis:
        g.close()


@pytest.mark.parametrize("prefilter", [False, True])
def test_sharded_flat_scan_equals_the_unsharded_scan_bit_for_bit(prefilter):
    from kektordb_b200 import GpuIndex
    from kektordb_b200.sharding import ShardGroup, shard_range
    _gpu_count()
    n, dim, k, world = 9000, 96, 20, 3
    X, Q = _make_corpus(n, dim, 33)
    u = np.random.default_rng(1).random(n)
    full = GpuIndex(dim, "euclidean", 8, n)
    full.AddBatch(X[:200], u[:200], 40)
    full.AddBatch(X[200:], u[200:], 40)
    gis, bases = [], []
    for r in range(world):
        base, count = shard_range(n, world, r)
        g = GpuIndex(dim, "euclidean", 8, count)
        g.AddBatch(X[base:base + 200], u[base:base + 200], 40)
        g.AddBatch(X[base + 200:base + count], u[base + 200:base + count], 40)
        gis.append(g)
        bases.append(base)
    grp = ShardGroup.local(gis, bases)
    for mode in (0, 1):
        want = full.flat_search(Q, k, mode)
        got = grp.flat_search(Q, k, mode, prefilter=prefilter)
        _assert_same(got, want)
    allow = O.dense_bitset(np.where(np.random.default_rng(2).random(n) < 0.3)[0] + 1, n)
    _assert_same(grp.flat_search(Q, k, 0, allow, prefilter=prefilter), full.flat_search(Q, k, 0, allow))
    grp.close()
    full.close()
    for g in gis:
        g.close()


def test_candidate_heap_overflow_is_reported_on_every_entry_point():
    import ctypes as C
    import torch
    from kektordb_b200 import ffi
    from kektordb_b200.sharding import ShardGroup
    _gpu_count()
    n, dim, k, ef = 4000, 16, 10, 200
    X, Q = _make_corpus(n, dim, 44, dup=False)
    shards = _oracle_shards(X, 1, O.METRIC_L2, 8, 60, 400)
    gi = _mirror(shards[0][0], O.METRIC_L2, 8)
    lib = ffi.lib()
    gi.set_tuning(cand_smem=8)
    ffi.check(lib.kdbgpu_set_candidate_bound(gi._h, 8))  # a candidate heap of 16 entries: ef = 200 overflows it
    with pytest.raises(ffi.GpuError) as ei:
        gi.SearchWithScores(Q, k, None, ef)
    assert ei.value.code == ffi.ERR_OVERFLOW
    # device-resident entry point: the launch itself cannot fail, kdbgpu_last_search_stats reports it
    dev = torch.device("cuda", 0)
    dq = torch.from_numpy(Q).to(dev)
    di = torch.zeros((len(Q), k), dtype=torch.int32, device=dev)
    ds = torch.zeros((len(Q), k), dtype=torch.float64, device=dev)
    dc = torch.full((len(Q),), 7, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    gi.search_device(dq.data_ptr(), len(Q), k, ef, di.data_ptr(), ds.data_ptr(), dc.data_ptr())
    st = ffi.Stats()
    assert lib.kdbgpu_last_search_stats(gi._h, C.byref(st)) == ffi.ERR_OVERFLOW
    assert st.dist_evals > 0
    assert (dc.cpu().numpy() == 0).any()  # the overflowing queries came back empty
    # the shard group: host form fails at wait, device form at sync
    grp = ShardGroup.local([gi], [0])
    with pytest.raises(ffi.GpuError) as ei:
        grp.SearchWithScores(Q, k, None, ef)
    assert ei.value.code == ffi.ERR_OVERFLOW
    grp.search_device(dq.data_ptr(), len(Q), k, ef, di.data_ptr(), ds.data_ptr(), dc.data_ptr())
    with pytest.raises(ffi.GpuError) as ei:
        grp.sync()
    assert ei.value.code == ffi.ERR_OVERFLOW
    grp.sync()  # the flag is cleared once reported
    # with the default bound restored the same search is exact again
    ffi.check(lib.kdbgpu_set_candidate_bound(gi._h, 1 << 15))
    gi.set_tuning(cand_smem=192)
    want = shards[0][0].search_batch(Q, k, ef, threads=8)
    got = gi.SearchWithScores(Q, k, None, ef)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    grp.close()
    gi.close()


def test_removed_node_named_by_a_stale_row_is_skipped_at_search_time():
    """hnsw_index.go:2553-2561: a neighbour whose node is nil is skipped when the row is read, not only
    when the row is staged (Vacuum nils nodes; a row the host has not re-patched may still name them)."""
    _gpu_count()
    n, dim, k, ef = 1500, 24, 10, 64
    X, Q = _make_corpus(n, dim, 55, dup=False)
    shards = _oracle_shards(X, 1, O.METRIC_COSINE, 8, 60, 500)
    oi = shards[0][0]
    gi = _mirror(oi, O.METRIC_COSINE, 8)
    base_ids = gi.SearchWithScores(Q, k, None, ef)[0]
    victims = np.unique(base_ids[:, 0])
    victims = victims[victims != oi.entry][:5].astype(np.uint32)  # nodes the queries certainly reach
    gi.remove_nodes(victims)  # rows naming them are NOT re-patched
    ids, sc, cnt, _ = gi.SearchWithScores(Q, k, None, ef)
    assert not np.isin(ids[ids > 0], victims).any()
    assert (cnt == k).all()
    gi.close()


# ---- two devices ------------------------------------------------------------------------------------
def test_local_group_over_two_devices_uses_peer_copies():
    from kektordb_b200.sharding import ShardGroup
    if _gpu_count() < 2:
        pytest.skip("needs two CUDA devices")
    n, dim, k, ef, world = 3000, 48, 10, 48, 2
    X, Q = _make_corpus(n, dim, 66)
    shards = _oracle_shards(X, world, O.METRIC_COSINE, 8, 60, 600)
    gis = [_mirror(oi, O.METRIC_COSINE, 8, device=r) for r, (oi, _, _) in enumerate(shards)]
    grp = ShardGroup.local(gis, [b for _, b, _ in shards])
    _assert_same(grp.SearchWithScores(Q, k, None, ef), _oracle_answer(shards, Q, k, ef))
    _assert_same(grp.flat_search(Q, k, 1), grp.flat_search(Q, k, 1, prefilter=True))
    grp.close()
    for g in gis:
        g.close()


def _rank_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from kektordb_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # plumbing only: carries the 128-byte id
    n, dim, k, ef = 3000, 48, 10, 48
    X, Q = _make_corpus(n, dim, 77)
    shards = _oracle_shards(X, world, O.METRIC_COSINE, 8, 60, 700)
    oi, base, count = shards[rank]
    gi = _mirror(oi, O.METRIC_COSINE, 8, device=rank)
    uid = [sharding.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    grp = sharding.ShardGroup.rank(gi, rank, world, uid[0], base)
    got = grp.SearchWithScores(Q, k, None, ef)
    tickets = [grp.submit(np.ascontiguousarray(Q[i::3]), k, ef) for i in range(3)]
    parts = [grp.wait(t) for t in tickets]
    fl = grp.flat_search(Q, k, 1, prefilter=True)
    np.savez(out_path % rank, ids=got[0], sc=got[1], cnt=got[2], e=got[3].dist_evals, xms=got[3].exchange_ms,
             p_ids=np.concatenate([p[0] for p in parts]), f_ids=fl[0], f_sc=fl[1])
    dist.barrier()
    grp.close()
    gi.close()
    dist.destroy_process_group()


def test_rank_group_over_nccl_two_processes_matches_the_oracle(tmp_path):
    import torch.multiprocessing as mp
    if _gpu_count() < 2:
        pytest.skip("needs two CUDA devices")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_rank_worker, args=(2, port, out), nprocs=2, join=True)
    n, dim, k, ef, world = 3000, 48, 10, 48, 2
    X, Q = _make_corpus(n, dim, 77)
    shards = _oracle_shards(X, world, O.METRIC_COSINE, 8, 60, 700)
    want = _oracle_answer(shards, Q, k, ef)
    full = O.OracleIndex(dim, O.METRIC_COSINE, 8, 60, O.ARITH_KERNEL, n)
    full.build_batched(X, np.random.default_rng(1).random(n), batch=256, threads=8)
    f_ids, f_sc, _ = full.flat_search_batch(Q, k, mode=1, threads=8)
    for r in range(world):  # every rank holds the merged result
        got = np.load(out % r)
        _assert_same((got["ids"], got["sc"], got["cnt"]), want)
        assert int(got["e"]) == want[3][0]
        assert np.array_equal(got["p_ids"], np.concatenate([want[0][i::3] for i in range(3)]))
        assert np.array_equal(got["f_ids"], f_ids) and np.array_equal(got["f_sc"], f_sc)
