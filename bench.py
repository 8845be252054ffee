#!/usr/bin/env python
"""bench.py — top-10 queries/sec at measured recall@10 on 1M x 768-d cosine, HNSW M=32 efSearch=128,
batch = 1024 queries (BASELINE.json configs[1]), on 1/2/4/8 B200s, with the reference's CPU path
timed beside it.

One "step" = one pass of the hot path over one batch of 1024 synthetic queries:
(*Index).SearchWithScores for every query of the batch (normalise, level descent, level-0 beam
search with ef=128, top-10).  `value` times it with queries and results resident in HBM;
`e2e` times the same call through the reference-facing C ABI with HOST buffers (pinned), H2D and
D2H copies inside the timed region.

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Data: synthetic random-normal vectors with a low-rank covariance (latent dimension 32 + isotropic
noise); see DESIGN.md §6 for why i.i.d. isotropic N(0,1) cannot meet the recall bar with ANY HNSW.
The graph is built on the GPU by kdbgpu_add_batch (bit-identical to the oracle's restatement of the
reference's AddBatch); the same graph feeds the GPU and the CPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--mode", default="replica", choices=["replica", "shard"],
                   help="N>1: replica = full corpus per GPU, queries split (no collective); "
                        "shard = corpus split by id range, NCCL all-gather of per-shard top-k + merge")
    p.add_argument("--n", "--corpus-size", dest="n", type=int, default=1_000_000,
                   help="rows in the corpus (under torchrun spell it --corpus-size: torchrun's own parser chokes on --n)")
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--m", type=int, default=32)
    p.add_argument("--efc", type=int, default=200)
    p.add_argument("--ef", type=int, default=128)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--latent", type=int, default=32)
    p.add_argument("--noise", type=float, default=0.1)
    p.add_argument("--data-model", default="lowrank", choices=["lowrank", "iid"],
                   help="lowrank: x = zW + noise*e (default); iid: isotropic N(0,1) (recall collapses, DESIGN.md §6)")
    p.add_argument("--build-batch", type=int, default=16384)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="bound of the cpu_baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-single-call", action="store_true", help="skip the one-query-per-call (micro-batcher) measurement")
    p.add_argument("--single-call-threads", type=int, default=0,
                   help="native caller threads of the one-query-per-call measurement (0 = batches in flight x batch)")
    p.add_argument("--batcher-wait-us", type=int, default=1000,
                   help="micro-batcher deadline; a query's own service time is ~4 ms, and waking 1024 blocked OS "
                        "threads spreads their next requests over ~1-2 ms, so shorter deadlines form small batches")
    p.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "none"])
    p.add_argument("--workload", default="hnsw", choices=["hnsw", "flat", "hybrid", "quantized"],
                   help="hnsw = BASELINE configs[1] (the headline); flat = configs[2] (tensor-core flat top-100); "
                        "hybrid = configs[4] (1536-d HNSW + 10 %% allow-list); quantized = configs[1]'s corpus held "
                        "as int8 (cosine) or float16 (euclidean) rows, SURVEY.md §8 f-4 — see bench_extra.py")
    p.add_argument("--precision", default="int8", choices=["int8", "float16"], help="--workload quantized only")
    p.add_argument("--overlap", type=int, default=3,
                   help="batches in flight: consecutive steps alternate over this many streams / caller threads")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# data + index construction (setup, untimed)
# ------------------------------------------------------------------------------------------------
def make_data(torch, n, dim, latent, noise, seed, device):
    """Random-normal vectors with low-rank covariance: x = z W + noise * e, z ~ N(0, I_latent).
    latent <= 0 selects i.i.d. isotropic N(0, 1)."""
    if latent <= 0:
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        out = torch.empty(n, dim, device=device, dtype=torch.float32)
        for i in range(0, n, 1 << 18):
            c = min(1 << 18, n - i)
            out[i:i + c] = torch.randn(c, dim, generator=g, device=device)
        return out
    g = torch.Generator(device=device)
    g.manual_seed(777)
    W = torch.randn(latent, dim, generator=g, device=device) / latent ** 0.5
    g.manual_seed(seed)
    out = torch.empty(n, dim, device=device, dtype=torch.float32)
    step = 1 << 18
    for i in range(0, n, step):
        c = min(step, n - i)
        z = torch.randn(c, latent, generator=g, device=device)
        e = torch.randn(c, dim, generator=g, device=device)
        out[i:i + c] = z @ W + noise * e
    return out


def build_schedule(n, ef_const, bmax):
    """AddBatch call sizes: the first call (index smaller than efConstruction) goes through
    sequential single Adds in the reference, so it is kept small; later calls never exceed the
    current index size (batch members do not see each other, hnsw_index.go:1789-1853)."""
    sched = [min(n, ef_const)]
    while sum(sched) < n:
        sched.append(min(bmax, sum(sched), n - sum(sched)))
    return sched


def build_index(torch, GpuIndex, X, m, efc, bmax, level_seed, device_index, metric="cosine"):
    n, dim = X.shape
    gi = GpuIndex(dim, metric, m, n, device=device_index)
    u = np.random.default_rng(level_seed).random(n)
    pos = 0
    t0 = time.time()
    for b in build_schedule(n, efc, bmax):
        gi.add_batch_device(X[pos:pos + b].data_ptr(), b, dim, u[pos:pos + b], efc)
        pos += b
    torch.cuda.synchronize()
    return gi, time.time() - t0


def recall_at_k(ids, gt):
    return float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / gt.shape[1] for i in range(len(gt))]))


class ClockSampler:
    """Samples SM clocks and throttle reasons DURING the timed region (B200_PROFILING.md's clocks
    line).  kind "nvml" polls NVML in-process every 50 ms (the same counters nvidia-smi prints,
    without a second process hammering the driver); kind "smi" runs `nvidia-smi -lms 200`."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, kind="nvml"):
        self.gpu, self.kind = gpu_index, kind
        self.sm, self.smax, self.reasons, self.power = [], [], set(), []
        self.proc = self.thread = None
        self.stop_flag = False

    def start(self):
        if self.kind == "none":
            return
        try:
            if self.kind == "nvml":
                import pynvml
                pynvml.nvmlInit()
                idx = self.gpu
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                if vis:
                    try:
                        idx = int(vis.split(",")[self.gpu])
                    except Exception:
                        idx = self.gpu
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
                self.nv = pynvml
                self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            else:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                     "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception as ex:
            self.kind, self.err = "failed", repr(ex)

    def _poll_nvml(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                 else nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                         getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = get_reasons(self.h)
                for name, bit in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                self.sm.append(float(f[1]))
                self.smax.append(float(f[2]))
                self.power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        if self.kind in ("none", "failed"):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                    "note": "sampler " + (self.kind if self.kind == "none" else getattr(self, "err", "failed"))}
        self.stop_flag = True
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.smax) if self.smax else None, "samples": len(self.sm),
                "power_w_max": round(max(self.power), 1) if self.power else None, "reasons": sorted(self.reasons),
                "source": self.kind}


# ------------------------------------------------------------------------------------------------
def oracle_from_gpu(gi, m, efc, arith):
    """Load the GPU-built graph + stored rows into the CPU oracle (same graph for both arms)."""
    from oracle import oracle as O
    n, levels, node_row, row_off, nbrs, entry, max_level = gi.get_graph()
    vec = np.zeros((n + 1, gi.dim), dtype=np.float32)
    step = 1 << 17
    for i in range(1, n + 1, step):
        c = min(step, n + 1 - i)
        vec[i:i + c] = gi.download_vectors(i, c)
    oi = O.OracleIndex(gi.dim, O.METRIC_COSINE, m, efc, arith, n)
    oi.import_graph(vec, O.Graph(n, levels, node_row, row_off, nbrs, np.zeros(n + 1, np.uint8), entry, max_level))
    del vec
    return oi


def main():
    args = parse_args()
    if args.data_model == "iid":
        args.latent = 0
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 alone

    import torch
    from kektordb_b200 import GpuIndex, ffi

    if not torch.cuda.is_available() or ffi.lib().kdbgpu_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.workload != "hnsw":
        import bench_extra
        if world > 1 and args.workload == "flat" and args.impl == "ours":
            if args.data_model == "lowrank" and "--data-model" not in sys.argv:
                args.latent = 0
            return bench_extra.run_flat_sharded(args, torch, sys.modules[__name__])
        if world > 1:
            raise SystemExit("--workload hybrid/quantized are single-GPU lines")
        if args.workload == "flat" and args.data_model == "lowrank" and "--data-model" not in sys.argv:
            args.latent = 0  # configs[2] is quoted on plain random-normal vectors; exact search has no recall issue
        fn = {"flat": bench_extra.run_flat, "hybrid": bench_extra.run_hybrid, "quantized": bench_extra.run_quantized}
        return fn[args.workload](args, torch, sys.modules[__name__])
    dist = None
    if world > 1 and args.impl == "ours":
        # keep stdout to the one JSON line: NCCL's version banner / warnings go to stderr
        os.environ.pop("NCCL_DEBUG", None) if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION" else None
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    ncores = len(os.sched_getaffinity(0))
    k, ef, B, D, N = args.k, args.ef, args.batch, args.dim, args.n
    n_steps_total = args.warmup + args.steps
    mode = args.mode if (world > 1 and args.impl == "ours") else "single"

    # ---- setup (untimed): corpus, graph, queries -------------------------------------------------
    X = make_data(torch, N, D, args.latent, args.noise, 42, dev)
    if mode == "shard":
        from kektordb_b200.sharding import shard_range
        base, n_local = shard_range(N, world, rank)
        Xl = X[base:base + n_local].contiguous()
        del X
        X = Xl
    else:
        n_local, base = N, 0
    gi, build_s = build_index(torch, GpuIndex, X, args.m, args.efc, args.build_batch, 1, local_rank)
    # queries: replicas answer different queries per rank, shards answer the same ones
    q_seed = 4242 + (rank if mode == "replica" else 0)
    Qd = make_data(torch, n_steps_total * B, D, args.latent, args.noise, q_seed, dev)
    Qh = torch.empty((n_steps_total * B, D), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()
    Qh_np = Qh.numpy()

    if args.impl == "reference":
        return run_reference_arm(args, gi, Qh_np, ncores, build_s)

    # ---- ground truth + recall (untimed) ---------------------------------------------------------
    n_gt = min(256, B)
    gi.prepare_search(B, k, ef)  # every launch workspace allocated up front (no allocation inside a timed region)
    ids0, sc0, cnt0, st0 = gi.SearchWithScores(Qh_np[:B], k, None, ef)  # also the first warm-up
    gt_ids, gt_sc, _, _ = gi.flat_search(Qh_np[:n_gt], k, 1)
    recall_local = recall_at_k(ids0[:n_gt], gt_ids)

    # ---- device-resident timing (`value`) --------------------------------------------------------
    n_ov = max(1, min(4, args.overlap)) if mode != "shard" else 1
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_ov)]
    stream = streams[0]
    d_ids_l = [torch.zeros((B, k), dtype=torch.int32, device=dev) for _ in range(n_ov)]
    d_sc_l = [torch.zeros((B, k), dtype=torch.float64, device=dev) for _ in range(n_ov)]
    d_cnt_l = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(n_ov)]
    d_ids, d_sc, d_cnt = d_ids_l[0], d_sc_l[0], d_cnt_l[0]
    g_ids = g_sc = g_cnt = m_ids = m_sc = m_cnt = None
    if mode == "shard":
        g_ids = torch.zeros((world, B, k), dtype=torch.int32, device=dev)
        g_sc = torch.zeros((world, B, k), dtype=torch.float64, device=dev)
        g_cnt = torch.zeros((world, B), dtype=torch.int32, device=dev)
        m_ids, m_sc, m_cnt = torch.zeros_like(d_ids), torch.zeros_like(d_sc), torch.zeros_like(d_cnt)

    def step_device(i):
        q = Qd[i * B:(i + 1) * B]
        j = i % n_ov  # consecutive batches alternate streams so that one batch's tail overlaps the next one's head
        gi.search_device(q.data_ptr(), B, k, ef, d_ids_l[j].data_ptr(), d_sc_l[j].data_ptr(), d_cnt_l[j].data_ptr(),
                         streams[j].cuda_stream)
        if mode == "shard":  # the one exchange step: all-gather of per-shard top-k, then merge
            with torch.cuda.stream(stream):
                gl = torch.where(d_ids > 0, d_ids + base, d_ids)
                dist.all_gather_into_tensor(g_ids.view(world * B, k), gl)
                dist.all_gather_into_tensor(g_sc.view(world * B, k), d_sc)
                dist.all_gather_into_tensor(g_cnt.view(world * B), d_cnt)
            ffi.check(ffi.lib().kdbgpu_merge_topk_device(gi._h, world, B, k, g_ids.data_ptr(), g_sc.data_ptr(),
                                                         g_cnt.data_ptr(), m_ids.data_ptr(), m_sc.data_ptr(),
                                                         m_cnt.data_ptr(), stream.cuda_stream))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank, args.clock_sampler)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_e = tot_h = tot_h0 = 0
    barrier()
    ev0.record(streams[0])
    for st_ in streams[1:]:
        st_.wait_event(ev0)  # every stream starts after the start mark
    for i in range(args.warmup, n_steps_total):
        step_device(i)
    for st_ in streams[1:]:
        streams[0].wait_stream(st_)  # the end mark follows the last kernel of every stream
    ev1.record(streams[0])
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    # counters of the last launch stand for the per-step work (same graph, i.i.d. query batches)
    st_last = gi.last_search_stats()
    tot_e, tot_h, tot_h0 = st_last.dist_evals, st_last.hops, st_last.hops_l0

    # ---- end-to-end timing through the C ABI with host buffers (`e2e`) -----------------------------
    if mode != "shard":
        for i in range(args.warmup):
            gi.SearchWithScores(Qh_np[i * B:(i + 1) * B], k, None, ef)
        barrier()

        def e2e_worker(j):  # one caller thread per in-flight batch; ctypes releases the GIL inside the C call
            torch.cuda.set_device(local_rank)
            for i in range(args.warmup + j, n_steps_total, n_ov):
                gi.SearchWithScores(Qh_np[i * B:(i + 1) * B], k, None, ef)

        workers = [threading.Thread(target=e2e_worker, args=(j,)) for j in range(n_ov)]
        t0 = time.perf_counter()
        for t in workers:
            t.start()
        for t in workers:
            t.join()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_d2h = B * k * 12 + B * 4 + 40
    else:
        # sharded: H2D of the batch, per-shard search, all-gather, merge, D2H of the merged top-k
        h_ids = torch.empty((B, k), dtype=torch.int32, pin_memory=True)
        h_sc = torch.empty((B, k), dtype=torch.float64, pin_memory=True)
        h_cnt = torch.empty(B, dtype=torch.int32, pin_memory=True)
        q_stage = torch.empty((B, D), dtype=torch.float32, device=dev)

        def e2e_step(i):
            with torch.cuda.stream(stream):
                q_stage.copy_(Qh[i * B:(i + 1) * B], non_blocking=True)
            gi.search_device(q_stage.data_ptr(), B, k, ef, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(),
                             stream.cuda_stream)
            with torch.cuda.stream(stream):
                gl = torch.where(d_ids > 0, d_ids + base, d_ids)
                dist.all_gather_into_tensor(g_ids.view(world * B, k), gl)
                dist.all_gather_into_tensor(g_sc.view(world * B, k), d_sc)
                dist.all_gather_into_tensor(g_cnt.view(world * B), d_cnt)
            ffi.check(ffi.lib().kdbgpu_merge_topk_device(gi._h, world, B, k, g_ids.data_ptr(), g_sc.data_ptr(),
                                                         g_cnt.data_ptr(), m_ids.data_ptr(), m_sc.data_ptr(),
                                                         m_cnt.data_ptr(), stream.cuda_stream))
            with torch.cuda.stream(stream):
                h_ids.copy_(m_ids, non_blocking=True)
                h_sc.copy_(m_sc, non_blocking=True)
                h_cnt.copy_(m_cnt, non_blocking=True)
            stream.synchronize()

        for i in range(args.warmup):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.warmup, n_steps_total):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_d2h = B * k * 12 + B * 4
    # ---- the reference's own call shape: one blocking single-query call per caller thread, batched by the
    # ---- library's micro-batcher; native caller threads, rank 0, single GPU only
    one_call = None
    if mode == "single" and not args.no_single_call:
        try:
            from kektordb_b200 import Batcher
            from tests.native import driver
            bt = Batcher(gi, max_batch=B, max_wait_us=args.batcher_wait_us)
            n_callers = args.single_call_threads or n_ov * B
            driver.run_callers(bt, Qh_np[:args.warmup * B], k, ef, n_callers)  # warm-up
            st_w = bt.stats()
            ids1, sc1, cnt1, secs1 = driver.run_callers(bt, Qh_np[args.warmup * B:], k, ef, n_callers)
            st_b = bt.stats()
            nb = st_b.batches - st_w.batches
            one_call = {"value": round(args.steps * B / secs1, 1), "unit": "queries/s", "caller_threads": n_callers,
                        "mean_batch": round((st_b.queries - st_w.queries) / max(1, nb), 1), "batches": nb,
                        "max_wait_us": args.batcher_wait_us, "first_batch_equal_to_batched_call": bool(
                            np.array_equal(ids1[:B], gi.SearchWithScores(Qh_np[args.warmup * B:(args.warmup + 1) * B], k, None, ef)[0]))}
            bt.close()
        except Exception as ex:  # the main line must still print
            one_call = {"error": repr(ex)}
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks: max time, summed work --------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, recall_local], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = times.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tmin = times.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dev_ms, e2e_ms, recall = float(tmax[0]), float(tmax[1]), float(tmin[2])
    else:
        dev_ms, e2e_ms, recall = float(times[0]), float(times[1]), float(times[2])
    queries_per_step = B * (world if mode == "replica" else 1)
    if mode == "shard":
        # recall of the merged result against the merged exact scan (per-shard flat + same merge)
        recall = shard_recall(torch, dist, ffi, gi, Qh_np[:n_gt], k, base, world, m_ids, m_sc, m_cnt, d_ids, d_sc,
                              d_cnt, g_ids, g_sc, g_cnt, Qd, B, ef, stream, dev)
    value = queries_per_step * args.steps / (dev_ms / 1e3)
    e2e_value = queries_per_step * args.steps / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel (hnsw_search_kernel) ---------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    stride = (D + 127) // 128 * 128
    bytes_per_launch = tot_e * stride * 4 + tot_h0 * (2 * args.m) * 4 + (tot_h - tot_h0) * args.m * 4
    # average duration of ONE traversal launch, run alone (overlapped launches would hide each other's tails)
    kernel_ms = None
    if True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
        nrep = 5
        for r in range(nrep):
            gi.search_device(Qd[r * B:(r + 1) * B].data_ptr(), B, k, ef, d_ids.data_ptr(), d_sc.data_ptr(),
                             d_cnt.data_ptr(), stream.cuda_stream)
        with torch.cuda.stream(stream):
            e1.record(stream)
        torch.cuda.synchronize()
        kernel_ms = e0.elapsed_time(e1) / nrep
    isolated = bytes_per_launch / (kernel_ms / 1e3) / 1e9
    # over the timed region: K launches' algorithmic bytes / the region's device time (launches of
    # consecutive batches overlap, so this is the kernel's sustained rate; shard mode also holds the
    # collective in the region and reports the isolated launch instead)
    launch_ms = dev_ms / args.steps if mode != "shard" else kernel_ms
    achieved = bytes_per_launch / (launch_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "hnsw_search_kernel", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(bytes_per_launch), "avg_launch_ms": round(launch_ms, 4),
                "isolated_launch_ms": round(kernel_ms, 4), "isolated_frac": round(isolated / peak, 4),
                "dist_evals_per_query": round(tot_e / B, 1), "hops_per_query": round(tot_h / B, 1)}
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- CPU baseline (rank 0, N=1): the oracle port on this box's host cores -----------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline, parity = run_cpu_baseline(args, gi, Qh_np, ncores, ids0, sc0)
        except Exception as ex:  # the main line must still print
            cpu_baseline = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "top-10 queries/sec @ recall@10, 1Mx768-d cosine HNSW (M=32, efSearch=128, batch=1024)",
            "value": round(value, 1), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak" if mode != "shard" else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "recall_at_10": round(recall, 4),
            "config": {"workload": f"{N}x{D} cosine, HNSW M={args.m} efC={args.efc} efSearch={ef}, top-{k}, "
                                   f"batch={B} queries/step/GPU, 1xB200 per rank",
                       "parallelism": {"single": "1 GPU", "replica": f"{world} replicas, queries split, no collective",
                                       "shard": f"corpus split by id range over {world} GPUs, NCCL all-gather of "
                                                f"per-shard top-{k} + merge kernel"}[mode],
                       "data_model": (f"random-normal, low-rank covariance (latent {args.latent}, noise {args.noise}), "
                                      if args.latent > 0 else "random-normal, i.i.d. isotropic, ") +
                                     "seeds 42/4242; graph built on GPU (kdbgpu_add_batch), levels seed 1",
                       "l2_policy": "inputs larger than L2: 3.07 GB corpus, new query batch every step",
                       "batches_in_flight": n_ov, "build_seconds": round(build_s, 2), "host_cores": ncores},
            "e2e": {"value": round(e2e_value, 1), "unit": "queries/s",
                    "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": e2e_d2h,
                    "ms_per_step": round(e2e_ms / args.steps, 4), "one_query_per_call": one_call},
            "gpu_launches": 2 * args.steps + (args.steps if mode == "shard" else 0),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def shard_recall(torch, dist, ffi, gi, Qgt, k, base, world, m_ids, m_sc, m_cnt, d_ids, d_sc, d_cnt, g_ids, g_sc,
                 g_cnt, Qd, B, ef, stream, dev):
    """Recall of the sharded search: merged HNSW top-k vs merged exact per-shard scans."""
    n_gt = Qgt.shape[0]
    fi, fs, fc, _ = gi.flat_search(Qgt, k, 1)
    pad = B - n_gt
    fi = np.concatenate([np.where(fi > 0, fi + base, fi), np.zeros((pad, k), np.uint32)]).astype(np.int32)
    fs = np.concatenate([fs, np.zeros((pad, k))])
    fc = np.concatenate([fc, np.zeros(pad, np.uint32)]).astype(np.int32)
    t_i, t_s, t_c = (torch.from_numpy(a).to(dev) for a in (fi, fs, fc))
    gi2, gs2, gc2 = torch.zeros_like(g_ids), torch.zeros_like(g_sc), torch.zeros_like(g_cnt)
    dist.all_gather_into_tensor(gi2.view(world * B, k), t_i)
    dist.all_gather_into_tensor(gs2.view(world * B, k), t_s)
    dist.all_gather_into_tensor(gc2.view(world * B), t_c)
    e_ids, e_sc, e_cnt = torch.zeros_like(d_ids), torch.zeros_like(d_sc), torch.zeros_like(d_cnt)
    torch.cuda.synchronize()
    ffi.check(ffi.lib().kdbgpu_merge_topk_device(gi._h, world, B, k, gi2.data_ptr(), gs2.data_ptr(), gc2.data_ptr(),
                                                 e_ids.data_ptr(), e_sc.data_ptr(), e_cnt.data_ptr(), None))
    torch.cuda.synchronize()
    # the merged HNSW answer for the first batch
    gi.search_device(Qd[:B].data_ptr(), B, k, ef, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), stream.cuda_stream)
    with torch.cuda.stream(stream):
        gl = torch.where(d_ids > 0, d_ids + base, d_ids)
        dist.all_gather_into_tensor(g_ids.view(world * B, k), gl)
        dist.all_gather_into_tensor(g_sc.view(world * B, k), d_sc)
        dist.all_gather_into_tensor(g_cnt.view(world * B), d_cnt)
    ffi.check(ffi.lib().kdbgpu_merge_topk_device(gi._h, world, B, k, g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(),
                                                 m_ids.data_ptr(), m_sc.data_ptr(), m_cnt.data_ptr(), stream.cuda_stream))
    torch.cuda.synchronize()
    return recall_at_k(m_ids[:n_gt].cpu().numpy(), e_ids[:n_gt].cpu().numpy())


def run_cpu_baseline(args, gi, Qh_np, ncores, gpu_ids0, gpu_sc0):
    """cpu_baseline: the oracle port (reference-faithful search, 8-lane FMA distance as
    native/compute/src/lib.rs) on all host cores, on a bounded sample of the same workload."""
    from oracle import oracle as O
    k, ef, B = args.k, args.ef, args.batch
    oi = oracle_from_gpu(gi, args.m, args.efc, O.ARITH_KERNEL)
    # parity on the first 256 queries: bit-exact in the kernel's summation order
    npar = min(256, B)
    pid, psc, pcnt, _ = oi.search_batch(Qh_np[:npar], k, ef, threads=ncores)
    parity = {"queries": npar, "ids_equal": bool(np.array_equal(pid, gpu_ids0[:npar])),
              "scores_bit_equal": bool(np.array_equal(psc, gpu_sc0[:npar]))}
    oi.set_arith(O.ARITH_AVX2)
    rid, rsc, _, _ = oi.search_batch(Qh_np[:npar], k, ef, threads=ncores)  # also the CPU warm-up
    parity["topk_set_agreement_vs_avx2_order"] = round(float(np.mean(
        [len(set(rid[i].tolist()) & set(gpu_ids0[i].tolist())) / k for i in range(npar)])), 4)
    same = rid == gpu_ids0[:npar]
    parity["max_abs_score_diff_vs_avx2_order"] = float(np.max(np.abs(rsc[same] - gpu_sc0[:npar][same]))) if same.any() else None
    done, t0, nb = 0, time.perf_counter(), Qh_np.shape[0] // B
    i = 0
    while True:
        q = Qh_np[(i % nb) * B:(i % nb + 1) * B]
        oi.search_batch(q, k, ef, threads=ncores)
        done += B
        i += 1
        el = time.perf_counter() - t0
        if el >= args.cpu_seconds or i >= 64:
            break
    cpu = {"value": round(done / el, 1), "unit": "queries/s", "cores": ncores, "kind": "port",
           "sample": f"{done} queries ({i} batches of {B}) of the same workload in {el:.1f} s, "
                     "oracle port, AVX2-FMA 8-lane distance order, one query per thread"}
    return cpu, parity


def run_reference_arm(args, gi, Qh_np, ncores, build_s):
    """--impl reference: the reference's CPU implementation of the path (the Go/Rust reference cannot
    be built in this image, so this is the oracle port), all host threads, same graph and queries."""
    from oracle import oracle as O
    k, ef, B = args.k, args.ef, args.batch
    oi = oracle_from_gpu(gi, args.m, args.efc, O.ARITH_AVX2)
    gt_ids, _, _, _ = gi.flat_search(Qh_np[:min(256, B)], k, 1)
    gi.close()
    for i in range(args.warmup):
        ids, _, _, _ = oi.search_batch(Qh_np[i * B:(i + 1) * B], k, ef, threads=ncores)
        if i == 0:
            recall = recall_at_k(ids[:gt_ids.shape[0]], gt_ids)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        oi.search_batch(Qh_np[i * B:(i + 1) * B], k, ef, threads=ncores)
    el = time.perf_counter() - t0
    value = B * args.steps / el
    line = {
        "impl": "reference",
        "metric": "top-10 queries/sec @ recall@10, 1Mx768-d cosine HNSW (M=32, efSearch=128, batch=1024)",
        "value": round(value, 1), "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "recall_at_10": round(recall, 4),
        "config": {"workload": f"{args.n}x{args.dim} cosine, HNSW M={args.m} efC={args.efc} efSearch={ef}, top-{k}, "
                               f"batch={B} queries/step", "parallelism": f"CPU only, {ncores} threads",
                   "graph": "built on the GPU by kdbgpu_add_batch (bit-identical to the oracle's AddBatch), "
                            "searched on the CPU only", "build_seconds": round(build_s, 2)},
        "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": ncores, "kind": "port",
                         "sample": f"{args.steps} steps of {B} queries; Go/Rust reference not buildable here "
                                   "(no go/rustc): oracle port, AVX2-FMA 8-lane distance order"},
        "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
