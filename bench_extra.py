"""bench_extra.py — the other single-GPU configurations of BASELINE.json, behind `bench.py --workload`:

  flat    configs[2]: 1M x 768-d L2, flat brute force, top-100, 1024 queries per step.  The product path
          is kdbgpu_flat_search_batch(mode 0 | KDBGPU_FLAT_PREFILTER): tcgen05 bf16 Q x K^T nomination +
          exact float64 re-score under a certificate — results bit-identical to the exhaustive float64
          scan (BruteForceIndex.SearchWithScores, reference pkg/core/vector_index.go:104-162).
  hybrid  configs[4]: 1M x 1536-d cosine HNSW with an allow-list of 10 % selectivity (the dense form of
          the roaring bitmap DB.FindIDsByFilter returns, reference pkg/core/core.go:1766), top-10.

  quantized  configs[1]'s corpus and graph with the rows held as int8 (cosine) or float16 (euclidean) —
          the reference's other two precisions (distance.Int8 / distance.Float16).  The
          traversal kernel is HBM-bound, so 1 / 2 bytes per element instead of 4 is the lever.

Same JSON contract as bench.py's main line (value / e2e / roofline / cpu_baseline / clocks).  They are
extra lines: the driver's headline stays `bench.py` with no --workload flag (configs[1]).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1650.0, "bf16_tflops_sustained": None, "src": "fallback (B200_PROFILING.md)"}
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            out.update({"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
                        "bf16_tflops_sustained": j.get("bf16_tflops_sustained"), "src": "measured (MEASURED_PEAKS.json)"})
        except Exception:
            pass
    return out


def _flat_traffic():
    """DRAM bytes of one nomination-pass launch from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_flat.json")
    try:
        return int(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


def _rows_only_graph(gi, n):
    """Flat scans need rows and liveness only: every node on level 0, no links."""
    lv = np.zeros(n + 1, np.int32)
    lv[0] = -1
    gi.set_graph(n, lv, np.concatenate([[0], np.arange(n + 1)]).astype(np.uint64), np.zeros(n + 1, np.uint64),
                 np.zeros(1, np.uint32), 1, 0)


def run_flat(args, torch, bench):
    from kektordb_b200 import GpuIndex
    N, D, B = args.n, args.dim, args.batch
    k = args.k if args.k != 10 else 100  # configs[2] asks for top-100
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    ncores = len(os.sched_getaffinity(0))
    n_total = args.warmup + args.steps
    X = bench.make_data(torch, N, D, args.latent, args.noise, 42, dev)
    gi = GpuIndex(D, "euclidean", 8, N, device=local_rank)
    ffi = bench_ffi()
    ffi.check(ffi.lib().kdbgpu_upload_vectors_device(gi._h, 1, N, X.data_ptr(), D))
    _rows_only_graph(gi, N)
    Qd = bench.make_data(torch, n_total * B, D, args.latent, args.noise, 4242, dev)
    Qh = torch.empty((n_total * B, D), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()
    Q = Qh.numpy()

    if args.impl == "reference":
        return _flat_reference_arm(args, gi, X, Q, k, ncores)

    # parity (untimed): pre-filter vs exhaustive float64 scan on the GPU, and vs the CPU oracle
    n_par = 32
    a = gi.flat_search(Q[:n_par], k, 0)
    b0 = gi.flat_search(Q[:B], k, 0, prefilter=True)  # also the first warm-up (builds the bf16 mirror)
    parity = {"queries": n_par, "vs_exhaustive_f64_scan": {
        "ids_equal": bool(np.array_equal(a[0], b0[0][:n_par])), "scores_bit_equal": bool(np.array_equal(a[1], b0[1][:n_par]))}}
    for i in range(args.warmup):
        gi.flat_search(Q[i * B:(i + 1) * B], k, 0, prefilter=True)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(local_rank, args.clock_sampler)
    sampler.start()
    comp_ms = tens_ms = 0.0
    rescored = fallbacks = 0
    t0 = time.perf_counter()
    for i in range(args.warmup, n_total):
        _, _, _, st = gi.flat_search(Q[i * B:(i + 1) * B], k, 0, prefilter=True)
        comp_ms += st.kernel_ms
        tens_ms += st.hops_l0 / 1e6
        rescored += st.dist_evals
        fallbacks += st.hops
    torch.cuda.synchronize()
    e2e_one_s = time.perf_counter() - t0
    # sustained: the same steps repeated for --sustain-seconds (the K-step region above lasts ~30 ms: too short for
    # the clock sampler and for the power cap to show); device time of every kernel of the call, as `value`
    sustained = None
    if args.sustain_seconds > 0:
        sus_ms, sus_steps, t1 = 0.0, 0, time.perf_counter()
        while time.perf_counter() - t1 < args.sustain_seconds:
            i = args.warmup + sus_steps % args.steps
            sus_ms += gi.flat_search(Q[i * B:(i + 1) * B], k, 0, prefilter=True)[3].kernel_ms
            sus_steps += 1
        sustained = {"value": round(B * sus_steps / (sus_ms / 1e3), 1), "steps": sus_steps,
                     "seconds": round(time.perf_counter() - t1, 2)}
    # e2e: the same host-buffer calls from two caller threads — the library keeps two flat calls in flight (its own
    # stream and staging each), so one call's copies and host-side work overlap the other's kernels
    call_ms = []

    def one_call(i):
        c0 = time.perf_counter()
        gi.flat_search(Q[i * B:(i + 1) * B], k, 0, prefilter=True)
        call_ms.append((time.perf_counter() - c0) * 1e3)

    bench.e2e_threads(torch, local_rank, one_call, 2, 0, 4)  # both call slots of the handle have their buffers
    call_ms.clear()
    e2e_s = bench.e2e_threads(torch, local_rank, one_call, 2, args.warmup, args.steps)
    clocks = sampler.stop()
    pk = _peaks()
    value = B * args.steps / (comp_ms / 1e3)
    e2e = B * args.steps / e2e_s
    dp = (D + 63) // 64 * 64
    n_pad = (N + 255) // 256 * 256
    flops_full = 2.0 * B * n_pad * dp                      # one full Q x K^T pass (pass B, the dominant kernel)
    stride = int(os.environ.get("KDBGPU_FLAT_SAMPLE", "0")) or max(1, min(12, 8192 // (6 * k)))
    flops_step = flops_full * (1.0 + 1.0 / stride)         # + the sampled threshold pass
    achieved = flops_step * args.steps / (tens_ms / 1e3) / 1e12
    cpu = None
    if not args.no_cpu_baseline:
        cpu, par2 = _flat_cpu_baseline(args, gi, Q, k, ncores, b0)
        parity["vs_cpu_oracle"] = par2
    line = {
        "metric": f"top-{k} queries/sec, exact, {N}x{D}-d L2 flat brute force (batch={B})",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(comp_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 nomination (tcgen05, f32 accumulate) + f64 exact re-score", "data": "synthetic",
        "recall_at_k": 1.0 if parity["vs_exhaustive_f64_scan"]["ids_equal"] else None, "sustained": sustained,
        "config": {"workload": f"{N}x{D} L2 flat, top-{k}, batch={B} queries/step, 1xB200 (BASELINE configs[2])",
                   "data_model": "random-normal, i.i.d. isotropic" if args.latent <= 0 else
                                 f"random-normal, low-rank covariance (latent {args.latent})",
                   "l2_policy": "inputs larger than L2: 1.5 GB bf16 corpus + 3.07 GB f32 rows, new queries every step",
                   "pass_a_tile_stride": stride, "host_cores": ncores,
                   "value_is": "device time of every kernel of the call (CUDA events inside the library), copies excluded"},
        "e2e": {"value": round(e2e, 1), "unit": "queries/s", "h2d_bytes_per_step": B * D * 4,
                "d2h_bytes_per_step": B * k * 12 + B * 8, "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                "calls_in_flight": 2, "one_call_at_a_time": round(B * args.steps / e2e_one_s, 1),
                "call_ms_median": round(float(np.median(call_ms)), 3), "call_ms_max": round(max(call_ms), 3)},
        "gpu_launches": 8 * args.steps,
        "roofline": {"bound": "tensor", "kernel": "flat_tc_kernel (threshold pass + nomination pass)",
                     "achieved": round(achieved, 1), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": round(achieved / pk["bf16_tflops"], 4), "traffic": _flat_traffic(), "peak_source": pk["src"],
                     "peak_sustained": pk["bf16_tflops_sustained"],
                     "flops_per_step": flops_step, "tensor_passes_ms_per_step": round(tens_ms / args.steps, 4),
                     "exact_rescored_rows_per_query": round(rescored / (B * args.steps), 1),
                     "queries_sent_to_exhaustive_scan": int(fallbacks)},
        "cpu_baseline": cpu, "parity": parity, "clocks": clocks,
    }
    return line


def run_flat_sharded(args, torch, bench):
    """configs[2] over G GPUs (flat path): the corpus is split by contiguous id range, every rank
    scans its N/G rows for the same queries (tensor-core pre-filter + exact float64 re-score), and the library's
    shard group does the rest on the device: ONE ncclAllGather of the packed per-shard top-k and the merge kernel
    (kdbgpu_shard_flat_search_batch — no host round trip between scan and merge).  Work per rank shrinks with G,
    so this is reported with "scaling": "strong"."""
    import torch.distributed as dist
    from kektordb_b200 import GpuIndex, sharding
    ffi = bench_ffi()
    N, D, B = args.n, args.dim, args.batch
    k = args.k if args.k != 10 else 100
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # the banner would land on stdout, before the JSON line
        os.environ.pop("NCCL_DEBUG", None)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    n_total = args.warmup + args.steps
    X = bench.make_data(torch, N, D, args.latent, args.noise, 42, dev)
    base, n_local = sharding.shard_range(N, world, rank)
    gi = GpuIndex(D, "euclidean", 8, n_local, device=local_rank)
    ffi.check(ffi.lib().kdbgpu_upload_vectors_device(gi._h, 1, n_local, X[base:base + n_local].data_ptr(), D))
    _rows_only_graph(gi, n_local)
    full = None
    if rank == 0:  # the unsharded answer, for the exactness check of the merged result
        full = GpuIndex(D, "euclidean", 8, N, device=local_rank)
        ffi.check(ffi.lib().kdbgpu_upload_vectors_device(full._h, 1, N, X.data_ptr(), D))
        _rows_only_graph(full, N)
    del X
    uid = [sharding.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    grp = sharding.ShardGroup.rank(gi, rank, world, uid[0], base)
    Qd = bench.make_data(torch, n_total * B, D, args.latent, args.noise, 4242, dev)
    Qh = torch.empty((n_total * B, D), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()
    Q = Qh.numpy()
    for i in range(args.warmup):
        grp.flat_search(Q[i * B:(i + 1) * B], k, 0, prefilter=True)
    dist.barrier()
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(local_rank, args.clock_sampler)
    if rank == 0:
        sampler.start()
    scan_ms = xch_ms = mrg_ms = 0.0
    t0 = time.perf_counter()
    for i in range(args.warmup, n_total):
        out = grp.flat_search(Q[i * B:(i + 1) * B], k, 0, prefilter=True)
        scan_ms += out[3].traversal_ms
        xch_ms += out[3].exchange_ms
        mrg_ms += out[3].merge_ms
    dist.barrier()
    torch.cuda.synchronize()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([wall_s, scan_ms, xch_ms, mrg_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_s, scan_ms, xch_ms, mrg_ms = (float(x) for x in t)
    if rank == 0:
        n_par = 64
        f = full.flat_search(Q[(n_total - 1) * B:(n_total - 1) * B + n_par], k, 0, prefilter=True)
        exact = {"queries": n_par, "ids_equal_to_unsharded_scan": bool(np.array_equal(f[0], out[0][:n_par])),
                 "scores_bit_equal": bool(np.array_equal(f[1], out[1][:n_par]))}
        pk = _peaks()
        line = {
            "metric": f"top-{k} queries/sec, exact, {N}x{D}-d L2 flat brute force sharded over {world} GPUs (batch={B})",
            "value": round(B * args.steps / wall_s, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(wall_s / args.steps * 1e3, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16 nomination (tcgen05) + f64 exact re-score", "data": "synthetic",
            "recall_at_k": 1.0 if exact["ids_equal_to_unsharded_scan"] else None,
            "config": {"workload": f"{N}x{D} L2 flat, top-{k}, batch={B}; rows split by id range over {world} GPUs, library shard "
                                   f"group: ONE ncclAllGather of the packed per-shard top-{k} + merge kernel, no host bounce",
                       "value_is": "wall clock per step through the host-buffer call on every rank (H2D, scan, all-gather, "
                                   "merge, D2H), barrier on both sides, max over ranks",
                       "per_shard_scan_ms_per_step": round(scan_ms / args.steps, 4),
                       "allgather_ms_per_step": round(xch_ms / args.steps, 4), "merge_ms_per_step": round(mrg_ms / args.steps, 4)},
            "e2e": {"value": round(B * args.steps / wall_s, 1), "unit": "queries/s", "h2d_bytes_per_step": B * D * 4,
                    "d2h_bytes_per_step": B * k * 12 + B * 4 + 48},
            "gpu_launches": 10 * args.steps,
            "roofline": {"bound": "tensor", "kernel": "flat_tc_kernel", "achieved": None, "peak": pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": None, "traffic": None,
                         "note": "per-shard tensor passes; see the single-GPU line for the roofline of the kernel"},
            "cpu_baseline": None, "parity": exact, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    grp.close()
    dist.destroy_process_group()
    return 0


def bench_ffi():
    from kektordb_b200 import ffi
    return ffi


def _oracle_flat_index(gi, X_rows, n):
    from oracle import oracle as O
    oi = O.OracleIndex(gi.dim, O.METRIC_L2, 8, 16, O.ARITH_KERNEL, n)
    lv = np.zeros(n + 1, np.int32)
    lv[0] = -1
    g = O.Graph(n, lv, np.concatenate([[0], np.arange(n + 1)]).astype(np.uint64), np.zeros(n + 1, np.uint64),
                np.zeros(0, np.uint32), np.zeros(n + 1, np.uint8), 1, 0)
    oi.import_graph(X_rows, g)
    return oi


def _download_rows(gi, n):
    vec = np.zeros((n + 1, gi.dim), dtype=np.float32)
    step = 1 << 17
    for i in range(1, n + 1, step):
        c = min(step, n + 1 - i)
        vec[i:i + c] = gi.download_vectors(i, c)
    return vec


def _flat_cpu_baseline(args, gi, Q, k, ncores, gpu_first):
    oi = _oracle_flat_index(gi, _download_rows(gi, args.n), args.n)
    n_par = 8
    wi, ws, wc = oi.flat_search_batch(Q[:n_par], k, mode=0, threads=ncores)
    par = {"queries": n_par, "ids_equal": bool(np.array_equal(wi, gpu_first[0][:n_par])),
           "scores_bit_equal": bool(np.array_equal(ws, gpu_first[1][:n_par]))}
    done, t0 = 0, time.perf_counter()
    while True:
        oi.flat_search_batch(Q[done:done + ncores], k, mode=0, threads=ncores)
        done += ncores
        el = time.perf_counter() - t0
        if el >= args.cpu_seconds or done >= 512:
            break
    return {"value": round(done / el, 2), "unit": "queries/s", "cores": ncores, "kind": "port",
            "sample": f"{done} queries of the same workload in {el:.1f} s, oracle restatement of "
                      "BruteForceIndex.SearchWithScores (float64 scan + sort), one query per thread"}, par


def _flat_reference_arm(args, gi, X, Q, k, ncores):
    oi = _oracle_flat_index(gi, _download_rows(gi, args.n), args.n)
    gi.close()
    B = min(args.batch, 2 * ncores)  # a bounded sample per step: the CPU scan needs ~0.5 core-seconds per query
    for i in range(min(args.warmup, 1)):
        oi.flat_search_batch(Q[:B], k, mode=0, threads=ncores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        oi.flat_search_batch(Q[i * B:(i + 1) * B], k, mode=0, threads=ncores)
    el = time.perf_counter() - t0
    v = B * args.steps / el
    return ({
        "impl": "reference", "metric": f"top-{k} queries/sec, exact, {args.n}x{args.dim}-d L2 flat brute force",
        "value": round(v, 2), "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(el / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.n}x{args.dim} L2 flat, top-{k}; each step = {B} queries (bounded sample of the "
                               f"{args.batch}-query batch)", "parallelism": f"CPU only, {ncores} threads"},
        "cpu_baseline": {"value": round(v, 2), "unit": "queries/s", "cores": ncores, "kind": "port",
                         "sample": f"{args.steps} steps of {B} queries, oracle restatement of BruteForceIndex"},
        "e2e": {"value": round(v, 2), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})



def _measure(args, torch, bench, gi, Qd, Q, B, k, ef, local_rank, dev, n_ov, allow):
    """Device-resident timing with consecutive batches alternating over n_ov streams (as bench.py's headline),
    isolated launches, and end-to-end timing through the C ABI with host buffers from n_ov caller threads.
    allow: dense uint64 bitset (numpy) shared by every batch, or None.  Returns (dev_ms, iso_ms, e2e_s, stats, clocks)."""
    import threading
    n_total = args.warmup + args.steps
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_ov)]
    d_ids = [torch.zeros((B, k), dtype=torch.int32, device=dev) for _ in range(n_ov)]
    d_sc = [torch.zeros((B, k), dtype=torch.float64, device=dev) for _ in range(n_ov)]
    d_cnt = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(n_ov)]
    d_allow, allow_words, allow_first = None, 0, 0
    if allow is not None:
        d_allow = torch.from_numpy(allow.view(np.int64)).to(dev)
        allow_words = int(allow.size)
        nz = np.flatnonzero(allow)
        allow_first = int(nz[0] * 64 + (int(allow[nz[0]]) & -int(allow[nz[0]])).bit_length() - 1) if nz.size else 0

    def launch(i, j):
        gi.search_device(Qd[i * B:(i + 1) * B].data_ptr(), B, k, ef, d_ids[j].data_ptr(), d_sc[j].data_ptr(),
                         d_cnt[j].data_ptr(), streams[j].cuda_stream, d_allow.data_ptr() if d_allow is not None else 0,
                         allow_words, allow_first)

    for i in range(args.warmup):
        launch(i, i % n_ov)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(local_rank, args.clock_sampler)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(streams[0])
    for s_ in streams[1:]:
        s_.wait_event(ev0)
    for i in range(args.warmup, n_total):
        launch(i, i % n_ov)
    for s_ in streams[1:]:
        streams[0].wait_stream(s_)
    ev1.record(streams[0])
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    st = gi.last_search_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for r in range(5):
        launch(r, 0)
    e1.record(streams[0])
    torch.cuda.synchronize()
    iso_ms = e0.elapsed_time(e1) / 5
    for i in range(args.warmup):
        gi.SearchWithScores(Q[i * B:(i + 1) * B], k, allow, ef)

    def worker(j):
        torch.cuda.set_device(local_rank)
        for i in range(args.warmup + j, n_total, n_ov):
            gi.SearchWithScores(Q[i * B:(i + 1) * B], k, allow, ef)

    ws = [threading.Thread(target=worker, args=(j,)) for j in range(n_ov)]
    t1 = time.perf_counter()
    for t in ws:
        t.start()
    for t in ws:
        t.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t1
    return dev_ms, iso_ms, e2e_s, st, sampler.stop()

# --------------------------------------------------------------------------------------------------
def run_hybrid(args, torch, bench):
    """configs[4]: HNSW + allow-list (10 % selectivity).  Semantics of the reference: non-members are
    marked visited and skipped before any distance and are never traversed through
    (hnsw_index.go:2542-2549); the smallest member replaces a non-member entry point (:437-446)."""
    from kektordb_b200 import GpuIndex, dense_allow_list
    from oracle import oracle as O
    N, B, k, ef = args.n, args.batch, args.k, args.ef
    D = args.dim if args.dim != 768 else 1536
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    ncores = len(os.sched_getaffinity(0))
    n_total = args.warmup + args.steps
    X = bench.make_data(torch, N, D, args.latent, args.noise, 42, dev)
    gi, build_s = bench.build_index(torch, GpuIndex, X, args.m, args.efc, args.build_batch, 1, local_rank)
    del X
    Qd = bench.make_data(torch, n_total * B, D, args.latent, args.noise, 4242, dev)
    Qh = torch.empty((n_total * B, D), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()
    Q = Qh.numpy()
    sel = float(os.environ.get("KDB_SELECTIVITY", "0.1"))
    member = np.random.default_rng(7).random(N + 1) < sel
    member[0] = False
    allow = dense_allow_list(np.where(member)[0], N)

    if args.impl == "reference":
        oi = bench.oracle_from_gpu(gi, args.m, args.efc, O.ARITH_AVX2)
        gi.close()
        for i in range(args.warmup):
            oi.search_batch(Q[i * B:(i + 1) * B], k, ef, allow=allow, threads=ncores)
        t0 = time.perf_counter()
        for i in range(args.warmup, n_total):
            oi.search_batch(Q[i * B:(i + 1) * B], k, ef, allow=allow, threads=ncores)
        el = time.perf_counter() - t0
        v = B * args.steps / el
        return ({"impl": "reference", "metric": f"top-{k} queries/sec, {N}x{D}-d cosine HNSW + allow-list {sel:.0%}",
                          "value": round(v, 1), "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{N}x{D} cosine HNSW M={args.m} efSearch={ef} top-{k}, allow-list {sel:.0%}",
                                     "parallelism": f"CPU only, {ncores} threads"},
                          "cpu_baseline": {"value": round(v, 1), "unit": "queries/s", "cores": ncores, "kind": "port",
                                           "sample": f"{args.steps} steps of {B} queries, oracle port"},
                          "e2e": {"value": round(v, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0})

    gi.prepare_search(B, k, ef)
    ids0, sc0, cnt0, st0 = gi.SearchWithScores(Q[:B], k, allow, ef)
    n_gt = min(256, B)
    gt_ids, _, gt_cnt, _ = gi.flat_search(Q[:n_gt], k, 1, allow, prefilter=True)
    recall = bench.recall_at_k(ids0[:n_gt], gt_ids)
    only_members = bool(member[ids0[ids0 > 0]].all())
    n_ov = max(1, min(4, args.overlap))
    dev_ms, iso_ms, e2e_s, st, clocks = _measure(args, torch, bench, gi, Qd, Q, B, k, ef, local_rank, dev, n_ov, allow)
    kern_ms = dev_ms
    E, H, H0 = st.dist_evals * args.steps, st.hops * args.steps, st.hops_l0 * args.steps
    pk = _peaks()
    stride = (D + 127) // 128 * 128
    byts = E * stride * 4 + H0 * 2 * args.m * 4 + (H - H0) * args.m * 4
    achieved = byts / (kern_ms / 1e3) / 1e9
    cpu = parity = None
    if not args.no_cpu_baseline:
        oi = bench.oracle_from_gpu(gi, args.m, args.efc, O.ARITH_KERNEL)
        npar = min(256, B)
        pid, psc, pcnt, pst = oi.search_batch(Q[:npar], k, ef, allow=allow, threads=ncores)
        parity = {"queries": npar, "ids_equal": bool(np.array_equal(pid, ids0[:npar])),
                  "scores_bit_equal": bool(np.array_equal(psc, sc0[:npar])),
                  "counts_equal": bool(np.array_equal(pcnt.astype(np.uint32), cnt0[:npar])),
                  "results_are_allow_list_members": only_members}
        oi.set_arith(O.ARITH_AVX2)
        oi.search_batch(Q[:npar], k, ef, allow=allow, threads=ncores)
        done, t1, i = 0, time.perf_counter(), 0
        while True:
            oi.search_batch(Q[(i % n_total) * B:(i % n_total + 1) * B], k, ef, allow=allow, threads=ncores)
            done += B
            i += 1
            el = time.perf_counter() - t1
            if el >= args.cpu_seconds or i >= 64:
                break
        cpu = {"value": round(done / el, 1), "unit": "queries/s", "cores": ncores, "kind": "port",
               "sample": f"{done} queries in {el:.1f} s, oracle port, AVX2-FMA order, same graph and allow-list"}
    line = {
        "metric": f"top-{k} queries/sec @ recall@{k}, {N}x{D}-d cosine HNSW + allow-list ({sel:.0%} selectivity)",
        "value": round(B * args.steps / (kern_ms / 1e3), 1), "unit": "queries/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(kern_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "recall_at_10": round(recall, 4),
        "config": {"workload": f"{N}x{D} cosine, HNSW M={args.m} efC={args.efc} efSearch={ef}, top-{k}, allow-list "
                               f"Bernoulli({sel}) seed 7 shared by the batch, batch={B} (BASELINE configs[4])",
                   "l2_policy": "inputs larger than L2", "build_seconds": round(build_s, 2), "host_cores": ncores,
                   "batches_in_flight": n_ov},
        "e2e": {"value": round(B * args.steps / e2e_s, 1), "unit": "queries/s",
                "h2d_bytes_per_step": B * D * 4 + allow.nbytes, "d2h_bytes_per_step": B * k * 12 + B * 4 + 40,
                "ms_per_step": round(e2e_s / args.steps * 1e3, 4)},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "kernel": "hnsw_search_kernel", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": None, "peak_source": pk["src"],
                     "dist_evals_per_query": round(E / (B * args.steps), 1), "hops_per_query": round(H / (B * args.steps), 1)},
        "cpu_baseline": cpu, "parity": parity, "clocks": clocks,
    }
    return line


# --------------------------------------------------------------------------------------------------
def run_quantized(args, torch, bench):
    """the HNSW traversal over int8 (cosine) / float16 (euclidean) rows.  Distances follow
    the reference's arithmetic for the precision (hnsw_index.go:2398-2449, distance_go.go:93-118); int8
    results are bit-identical to the CPU path in any summation order (integer dot), float16 ones in
    kernel order.  The graph is built on the device with the precision's own distances (TrainQuantizer +
    AddBatch, as DB.Compress does) and is bit-identical to the oracle's build; both arms search the same graph."""
    import threading
    from kektordb_b200 import GpuIndex, ffi
    from oracle import oracle as O
    N, D, B, k, ef = args.n, args.dim, args.batch, args.k, args.ef
    prec = args.precision
    metric = "cosine" if prec == "int8" else "euclidean"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    ncores = len(os.sched_getaffinity(0))
    n_total = args.warmup + args.steps
    X = bench.make_data(torch, N, D, args.latent, args.noise, 42, dev)
    if metric == "cosine":  # DB.Compress feeds the new index the stored (unit-norm) float32 rows (core.go:1150-1165)
        X /= X.norm(dim=1, keepdim=True).clamp_min(1e-30)
    Qd = bench.make_data(torch, n_total * B, D, args.latent, args.noise, 4242, dev)
    Qh = torch.empty((n_total * B, D), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()
    Q = Qh.numpy()
    # exact float32 ground truth for recall: flat scan over the float32 rows
    n_gt = min(256, B)
    gf = GpuIndex(D, metric, 8, N, device=local_rank)
    gf.upload_vectors_device(1, X.data_ptr(), N, D)
    _rows_only_graph(gf, N)
    gt_ids, _, _, _ = gf.flat_search(Q[:n_gt], k, 1, prefilter=True)
    gf.close()
    # the quantized index, built on the device with its own distances: TrainQuantizer, then AddBatch
    # (what DB.Compress does, core.go:1224-1270)
    gi = GpuIndex(D, metric, args.m, N, device=local_rank, precision=prec)
    t0 = time.time()
    abs_max = gi.train_quantizer_device(X.data_ptr(), D, N) if prec == "int8" else None
    u = np.random.default_rng(1).random(N)
    pos = 0
    for b in bench.build_schedule(N, args.efc, args.build_batch):
        gi.add_batch_device(X[pos:pos + b].data_ptr(), b, D, u[pos:pos + b], args.efc)
        pos += b
    torch.cuda.synchronize()
    build_s = time.time() - t0
    stage_s = 0.0
    del X
    graph = gi.get_graph()
    step = 1 << 17
    oprec = O.PREC_I8 if prec == "int8" else O.PREC_F16

    def oracle_index(arith):
        n, levels, node_row, row_off, nbrs, entry, max_level = graph
        rows = np.zeros((n + 1, D), dtype=np.int8 if prec == "int8" else np.uint16)
        for i in range(1, n + 1, step):
            c = min(step, n + 1 - i)
            rows[i:i + c] = gi.download_rows_raw(i, c)
        oi = O.OracleIndex(D, O.METRIC_COSINE if prec == "int8" else O.METRIC_L2, args.m, args.efc, arith, n, precision=oprec)
        if prec == "int8":
            oi.set_quantizer(abs_max)
        oi.import_graph(rows, O.Graph(n, levels, node_row, row_off, nbrs, np.zeros(n + 1, np.uint8), entry, max_level))
        return oi

    metric_name = f"top-{k} queries/sec @ recall@{k}, {N}x{D}-d {metric} HNSW over {prec} rows (M={args.m}, efSearch={ef}, batch={B})"
    if args.impl == "reference":
        oi = oracle_index(O.ARITH_AVX2)
        gi.close()
        for i in range(args.warmup):
            oi.search_batch(Q[i * B:(i + 1) * B], k, ef, threads=ncores)
        t1 = time.perf_counter()
        for i in range(args.warmup, n_total):
            oi.search_batch(Q[i * B:(i + 1) * B], k, ef, threads=ncores)
        el = time.perf_counter() - t1
        v = B * args.steps / el
        return ({"impl": "reference", "metric": metric_name, "value": round(v, 1), "unit": "queries/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": prec, "data": "synthetic",
                          "config": {"workload": f"{N}x{D} {metric} HNSW over {prec} rows, efSearch={ef}, top-{k}",
                                     "parallelism": f"CPU only, {ncores} threads"},
                          "cpu_baseline": {"value": round(v, 1), "unit": "queries/s", "cores": ncores, "kind": "port",
                                           "sample": f"{args.steps} steps of {B} queries, oracle port"},
                          "e2e": {"value": round(v, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0})

    gi.prepare_search(B, k, ef)
    ids0, sc0, cnt0, st0 = gi.SearchWithScores(Q[:B], k, None, ef)
    recall = bench.recall_at_k(ids0[:n_gt], gt_ids)
    n_ov = max(1, min(4, args.overlap if "--overlap" in os.sys.argv else 4))
    dev_ms, iso_ms, e2e_s, st, clocks = _measure(args, torch, bench, gi, Qd, Q, B, k, ef, local_rank, dev, n_ov, None)
    pk = _peaks()
    esize = 1 if prec == "int8" else 2
    row_bytes = (D * esize + 127) // 128 * 128
    byts = st.dist_evals * (row_bytes + (4 if prec == "int8" else 0)) + st.hops_l0 * 2 * args.m * 4 + \
        (st.hops - st.hops_l0) * args.m * 4
    launch_ms = dev_ms / args.steps
    achieved = byts / (launch_ms / 1e3) / 1e9
    cpu = parity = None
    if not args.no_cpu_baseline:
        oi = oracle_index(O.ARITH_KERNEL)
        npar = min(256, B)
        pid, psc, pcnt, pst = oi.search_batch(Q[:npar], k, ef, threads=ncores)
        parity = {"queries": npar, "ids_equal": bool(np.array_equal(pid, ids0[:npar])),
                  "scores_bit_equal": bool(np.array_equal(psc, sc0[:npar])),
                  "counts_equal": bool(np.array_equal(pcnt.astype(np.uint32), cnt0[:npar]))}
        oi.set_arith(O.ARITH_AVX2)
        rid, rsc, _, _ = oi.search_batch(Q[:npar], k, ef, threads=ncores)
        parity["ids_equal_vs_reference_order"] = bool(np.array_equal(rid, ids0[:npar]))
        same = rid == ids0[:npar]
        parity["max_abs_score_diff_vs_reference_order"] = float(np.max(np.abs(rsc[same] - sc0[:npar][same]))) if same.any() else None
        done, t2, i = 0, time.perf_counter(), 0
        while True:
            oi.search_batch(Q[(i % n_total) * B:(i % n_total + 1) * B], k, ef, threads=ncores)
            done += B
            i += 1
            el = time.perf_counter() - t2
            if el >= args.cpu_seconds or i >= 64:
                break
        cpu = {"value": round(done / el, 1), "unit": "queries/s", "cores": ncores, "kind": "port",
               "sample": f"{done} queries in {el:.1f} s, oracle port over the same {prec} rows and graph, "
                         + ("exact int32 dot (dotProductGoInt8), " if prec == "int8" else "AVX2-FMA + F16C order (lib.rs:101-141), ")
                         + "one query per thread"}
    line = {
        "metric": metric_name, "value": round(B * args.steps / (dev_ms / 1e3), 1), "unit": "queries/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(launch_ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "i8 dot -> f64" if prec == "int8" else "f16 rows, f32 accumulate",
        "data": "synthetic", "recall_at_10": round(recall, 4),
        "config": {"workload": f"{N}x{D} {metric}, HNSW M={args.m} efC={args.efc} efSearch={ef}, top-{k}, batch={B}, rows held as "
                               f"{prec} ({row_bytes} B per row); recall is against the exact float32 scan",
                   "graph": f"built on the GPU over the {prec} rows with {prec} distances (kdbgpu_train_quantizer + "
                            "kdbgpu_add_batch); same graph for both arms",
                   "quantizer_abs_max": abs_max, "l2_policy": "inputs larger than L2, new query batch every step",
                   "batches_in_flight": n_ov, "build_seconds": round(build_s, 2), "stage_seconds": round(stage_s, 2),
                   "host_cores": ncores},
        "e2e": {"value": round(B * args.steps / e2e_s, 1), "unit": "queries/s", "h2d_bytes_per_step": B * D * 4,
                "d2h_bytes_per_step": B * k * 12 + B * 4 + 40, "ms_per_step": round(e2e_s / args.steps * 1e3, 4)},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "kernel": "hnsw_search_kernel", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": None, "peak_source": pk["src"],
                     "algorithmic_bytes_per_launch": int(byts), "avg_launch_ms": round(launch_ms, 4),
                     "isolated_launch_ms": round(iso_ms, 4),
                     "isolated_frac": round(byts / (iso_ms / 1e3) / 1e9 / pk["hbm_gbs"], 4),
                     "dist_evals_per_query": round(st.dist_evals / B, 1), "hops_per_query": round(st.hops / B, 1)},
        "cpu_baseline": cpu, "parity": parity, "clocks": clocks,
    }
    return line
