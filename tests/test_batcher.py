"""The micro-batcher's host logic (csrc/batcher.cpp) against a test-double batch executor: grouping by
(k, ef, allow-list), result routing, dispatch rules, error behaviour.  No GPU needed — the executor is
a Python callback behind kdbgpu_batcher_create_fn.  Mirrors the reference's call shape: one blocking
SearchWithScores per request goroutine (pkg/engine/ops.go:1006)."""
import threading
import time

import numpy as np
import pytest

from kektordb_b200 import Batcher, ffi

DIM = 8


def _executor(log, delay=0.0, fail_on=None):
    """ids[i] = floor(q[i,0]) + j, scores = q[i,1] + j, count = k - (floor(q[i,0]) % 2)."""
    def fn(q, k, ef, allow):
        log.append((q.shape[0], k, ef, None if allow is None else allow.copy()))
        if delay:
            time.sleep(delay)
        if fail_on is not None and (q[:, 0] == fail_on).any():
            raise ffi.GpuError(ffi.ERR_OVERFLOW, "boom")
        base = np.floor(q[:, 0]).astype(np.uint32)
        ids = base[:, None] + np.arange(k, dtype=np.uint32)[None, :]
        sc = q[:, 1].astype(np.float64)[:, None] + np.arange(k)[None, :]
        return ids, sc, (k - (base % 2)).astype(np.uint32)
    return fn


def _q(i):
    v = np.zeros(DIM, np.float32)
    v[0], v[1] = i, i * 0.5
    return v


def _run(b, reqs):
    """reqs: list of (i, k, ef, allow); returns {i: (ids, scores)} from one thread per request."""
    out, threads = {}, []

    def call(i, k, ef, allow):
        out[i] = b.SearchWithScores(_q(i), k, allow, ef)

    for r in reqs:
        threads.append(threading.Thread(target=call, args=r))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return out


def test_lone_caller_is_dispatched_immediately():
    log = []
    b = Batcher(fn=_executor(log), dim=DIM, max_batch=64, max_wait_us=500_000)
    t0 = time.perf_counter()
    ids, sc = b.SearchWithScores(_q(4), 5, None, 20)
    assert time.perf_counter() - t0 < 0.2            # did not wait for the 0.5 s deadline: the device was idle
    assert ids.tolist() == [4, 5, 6, 7, 8] and sc.tolist() == [2.0, 3.0, 4.0, 5.0, 6.0]
    st = b.stats()
    assert (st.queries, st.batches, st.dispatched_idle) == (1, 1, 1)
    assert log == [(1, 5, 20, None)]
    b.close()


def test_concurrent_callers_share_batches_and_get_their_own_results():
    log = []
    b = Batcher(fn=_executor(log, delay=0.02), dim=DIM, max_batch=256, max_wait_us=100_000)
    n = 120
    out = _run(b, [(i, 4, 32, None) for i in range(n)])
    for i in range(n):
        ids, sc = out[i]
        cnt = 4 - (i % 2)
        assert ids.tolist() == [i + j for j in range(cnt)]
        assert sc.tolist() == [i * 0.5 + j for j in range(cnt)]
    st = b.stats()
    assert st.queries == n and st.batches < n / 4 and st.max_batch_seen > 8     # batched for real
    assert sum(e[0] for e in log) == n
    b.close()


def test_groups_are_keyed_by_k_ef_and_filter():
    log = []
    b = Batcher(fn=_executor(log, delay=0.03), dim=DIM, max_batch=256, max_wait_us=200_000)
    a1 = np.array([0b1010, 7], np.uint64)
    a2 = np.array([0b1010, 7], np.uint64)            # same membership, different buffer
    a3 = np.array([0b0110, 7], np.uint64)
    reqs = [(0, 3, 10, None)]                        # occupies the executor so the rest have to queue
    reqs += [(10 + i, 3, 10, None) for i in range(6)]
    reqs += [(20 + i, 5, 10, None) for i in range(6)]
    reqs += [(30 + i, 3, 99, None) for i in range(6)]
    reqs += [(40 + i, 3, 10, a1 if i % 2 else a2) for i in range(6)]
    reqs += [(50 + i, 3, 10, a3) for i in range(6)]
    out = _run(b, reqs)
    assert all(len(out[r[0]][0]) == r[1] - (r[0] % 2) for r in reqs)
    for nq, k, ef, allow in log:                     # no call mixes parameters; filters are passed through
        assert (k, ef) in ((3, 10), (5, 10), (3, 99))
    seen_filters = {None if a is None else tuple(a.tolist()) for _, _, _, a in log}
    assert seen_filters == {None, (0b1010, 7), (0b0110, 7)}
    assert len(log) <= 12 and b.stats().queries == len(reqs)
    b.close()


def test_full_group_is_dispatched_without_waiting_for_the_deadline():
    log = []
    b = Batcher(fn=_executor(log, delay=0.05), dim=DIM, max_batch=8, max_wait_us=5_000_000)
    t0 = time.perf_counter()
    _run(b, [(i, 2, 2, None) for i in range(1 + 8 * 3)])
    assert time.perf_counter() - t0 < 2.0            # far below the 5 s deadline
    st = b.stats()
    assert st.max_batch_seen == 8 and st.dispatched_full >= 2
    assert all(e[0] <= 8 for e in log)
    b.close()


def test_deadline_dispatch_while_device_is_busy():
    log = []
    b = Batcher(fn=_executor(log, delay=0.25), dim=DIM, max_batch=1000, max_wait_us=20_000)
    out = {}
    t1 = threading.Thread(target=lambda: out.__setitem__(1, b.SearchWithScores(_q(1), 2, None, 2)))
    t1.start()
    time.sleep(0.05)                                  # batch 1 is executing (0.25 s): the device is busy
    t0 = time.perf_counter()
    ids, _ = b.SearchWithScores(_q(2), 2, None, 2)    # waits 20 ms for company, then goes alone
    waited = time.perf_counter() - t0
    t1.join()
    assert ids.tolist() == [2, 3] and out[1][0].tolist() == [1]
    assert 0.015 < waited < 0.45 and b.stats().dispatched_deadline == 1
    b.close()


def test_failed_batch_yields_empty_results_and_the_error_code():
    log = []
    b = Batcher(fn=_executor(log, delay=0.02, fail_on=13.0), dim=DIM, max_batch=64, max_wait_us=50_000)
    ids, sc = b.SearchWithScores(_q(13), 3, None, 3)
    assert len(ids) == 0 and b.last_rc == ffi.ERR_OVERFLOW      # hnsw_index.go:355-359: error -> empty result
    ids, sc = b.SearchWithScores(_q(2), 3, None, 3)             # the batcher keeps working
    assert ids.tolist() == [2, 3, 4]
    b.close()


def test_argument_validation_and_shutdown():
    lib = ffi.lib()
    import ctypes as C
    h = C.c_void_p()
    assert lib.kdbgpu_batcher_create(None, 8, 10, C.byref(h)) == ffi.ERR_INVALID
    log = []
    b = Batcher(fn=_executor(log), dim=DIM)
    with pytest.raises(ValueError):
        b.SearchWithScores(np.zeros(DIM + 1, np.float32), 3)
    assert lib.kdbgpu_batcher_search(b._h, None, 3, 3, None, 0, None, None, None) == ffi.ERR_INVALID
    b.close()
    b.close()


# ---- asynchronous form: submit / poll / take (the Go shim's shape: no caller blocks inside the C call) --------
def test_submit_poll_take_routes_every_result_and_needs_no_caller_thread():
    log = []
    b = Batcher(fn=_executor(log, delay=0.01), dim=DIM, max_batch=64, max_wait_us=20_000)
    n, k = 300, 4
    t0 = time.perf_counter()
    tickets = {b.submit(_q(i), k, 32): i for i in range(n)}   # returns at once: nothing has been answered yet
    assert time.perf_counter() - t0 < 1.0 and len(tickets) == n
    seen = {}
    deadline = time.time() + 20
    while len(seen) < n and time.time() < deadline:
        for t in b.poll(128, 50_000):
            ids, sc, rc = b.take(t, k)
            i = tickets[t]
            assert rc == ffi.OK and i not in seen
            seen[i] = (ids, sc)
    assert len(seen) == n
    for i, (ids, sc) in seen.items():
        cnt = k - (i % 2)
        assert ids.tolist() == [i + j for j in range(cnt)] and sc.tolist() == [i * 0.5 + j for j in range(cnt)]
    st = b.stats()
    assert st.queries == n and st.batches <= n / 8 and st.max_batch_seen == 64
    assert b.poll(8, 1000) == []                                # every completion was announced exactly once
    b.close()


def test_take_blocks_until_done_and_rejects_unknown_or_spent_tickets():
    log = []
    b = Batcher(fn=_executor(log, delay=0.05), dim=DIM, max_batch=8, max_wait_us=1000)
    t = b.submit(_q(6), 3, 9)
    ids, sc, rc = b.take(t, 3)                                  # no poll needed: take waits for its query
    assert rc == ffi.OK and ids.tolist() == [6, 7, 8]
    assert b.take(t, 3)[2] == ffi.ERR_INVALID                   # a ticket is taken exactly once
    assert b.take(123456789, 3)[2] == ffi.ERR_INVALID
    assert b.poll(8, 1000) == [t]                               # the announcement is still delivered
    b.close()


def test_a_ticket_is_spent_by_one_take_while_its_group_still_has_results_outstanding():
    """A second take of the same ticket must not count the group's results out early (its staging would be
    recycled under the members that have not copied yet), and a slot nobody was given is not a ticket."""
    log = []
    b = Batcher(fn=_executor(log, delay=0.05), dim=DIM, max_batch=8, max_wait_us=200_000)
    hold = b.submit(_q(2), 3, 9)                                # dispatched as soon as a worker wakes up (idle device)
    t1, t2, t3 = (b.submit(_q(i), 3, 9) for i in (10, 20, 30))  # these collect in ONE group (with or behind `hold`)
    assert t1 >> 16 == t2 >> 16 == t3 >> 16
    assert b.take(t1, 3)[0].tolist() == [10, 11, 12]
    assert b.take(t1, 3)[2] == ffi.ERR_INVALID and b.take(t1, 3)[2] == ffi.ERR_INVALID
    assert b.take(((t1 >> 16) << 16) | 7, 3)[2] == ffi.ERR_INVALID   # slot 7 of that group was never handed out
    assert b.take(t2, 3)[0].tolist() == [20, 21, 22] and b.take(t3, 3)[0].tolist() == [30, 31, 32]
    assert b.take(hold, 3)[0].tolist() == [2, 3, 4]
    b.close()


def test_registered_filters_group_by_identity_and_reach_the_executor():
    log = []
    b = Batcher(fn=_executor(log, delay=0.03), dim=DIM, max_batch=256, max_wait_us=100_000)
    a1 = np.array([0b1010, 7], np.uint64)
    a3 = np.array([0b0110, 7], np.uint64)
    f1, f3 = b.register_filter(a1), b.register_filter(a3)
    b.submit(_q(0), 3, 10)                                      # occupies the executor so the rest queue up
    ts = [b.submit(_q(10 + i), 3, 10, filter_id=f1 if i % 2 else f3) for i in range(20)]
    ts += [b.submit(_q(40 + i), 3, 10, allowList=a1) for i in range(4)]   # raw bitset, same membership as f1
    got = {}
    while len(got) < len(ts) + 1:
        for t in b.poll(64, 50_000):
            got[t] = b.take(t, 3)
    assert all(got[t][2] == ffi.OK for t in ts)
    filters = [None if a is None else tuple(a.tolist()) for _, _, _, a in log]
    assert set(filters) == {None, (0b1010, 7), (0b0110, 7)}
    assert len(log) <= 5                                         # one batch per filter (+ raw / first), not one per query
    with pytest.raises(ffi.GpuError):
        b.submit(_q(1), 3, 10, filter_id=999)
    b.release_filter(f1)
    with pytest.raises(ffi.GpuError):
        b.release_filter(f1)
    b.close()


def test_failed_batch_reaches_async_callers_as_empty_result_plus_code():
    log = []
    b = Batcher(fn=_executor(log, delay=0.01, fail_on=13.0), dim=DIM, max_batch=64, max_wait_us=1000)
    t = b.submit(_q(13), 3, 3)
    ids, sc, rc = b.take(t, 3)
    assert rc == ffi.ERR_OVERFLOW and len(ids) == 0
    assert b.take(b.submit(_q(2), 3, 3), 3)[0].tolist() == [2, 3, 4]
    b.close()


def test_native_async_driver_equals_blocking_callers():
    """The load generator bench.py uses (tools/native): submit / poll / take from 4 submitter threads and one
    dispatcher gives every query the result the blocking call shape gives."""
    from tools.native import driver
    log = []
    b = Batcher(fn=_executor(log), dim=DIM, max_batch=32, max_wait_us=2000)
    Q = np.stack([_q(i) for i in range(500)])
    a = driver.run_async(b, Q, 4, 16, 4, 96)
    c = driver.run_callers(b, Q, 4, 16, 48)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and np.array_equal(a[2], c[2])
    assert a[0][7].tolist()[:3] == [7, 8, 9] and a[2].tolist() == [4 - (i % 2) for i in range(500)]
    b.close()
