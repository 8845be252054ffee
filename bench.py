#!/usr/bin/env python
"""bench.py — top-10 queries/sec at measured recall@10 on 1M x 768-d cosine, HNSW M=32 efSearch=128,
batch = 1024 queries (BASELINE.json configs[1]), on 1/2/4/8 B200s, with the reference's CPU path
timed beside it.

One "step" = one pass of the hot path over one batch of 1024 synthetic queries:
(*Index).SearchWithScores for every query of the batch (normalise, level descent, level-0 beam
search with ef=128, top-10).  `value` times it with queries and results resident in HBM;
`e2e` times the same call through the reference-facing C ABI with HOST buffers, H2D and D2H copies
inside the timed region (pinned buffers; the pageable figure, what a Go slice is, sits beside it).

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

N > 1 reports two things in the one JSON line: `value` = N replicas, queries split, no collective
(labelled), and `shard` = the north-star's layout — the corpus split by id range, one HNSW per GPU,
the library's shard group (kdbgpu_shard_*): per-shard traversal, ONE ncclAllGather of the packed
per-shard top-k per batch, merge kernel, up to 4 batches in flight — with its own queries/s, e2e,
recall, all-gather / merge times and parity against "G oracle indexes + exact merge".

At N = 1 the line also carries `extras`: compact results for BASELINE configs[2] (flat top-100) and
configs[4] (1536-d + 10 % allow-list), the i.i.d. N(0,1) data model, and queries sampled from the corpus;
at N = 8, configs[3] (10M x 768 over 8 GPUs).  `--no-extras` skips them.

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

METRIC = "top-10 queries/sec @ recall@10, 1Mx768-d cosine HNSW (M=32, efSearch=128, batch=1024)"


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--mode", default="both", choices=["both", "replica", "shard"],
                   help="N>1: replica = full corpus per GPU, queries split (no collective) -> `value`; shard = corpus "
                        "split by id range, library shard group with one NCCL all-gather per batch -> `shard`; "
                        "both (default) measures the two in one run")
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
    p.add_argument("--no-single-call", action="store_true", help="skip the one-query-per-call (micro-batcher) measurements")
    p.add_argument("--submitters", type=int, default=8,
                   help="submitter threads of the asynchronous one-query-per-call measurement (+1 dispatcher thread)")
    p.add_argument("--single-call-threads", type=int, default=0,
                   help="caller threads of the blocking one-query-per-call measurement (0 = batches in flight x batch)")
    p.add_argument("--batcher-wait-us", type=int, default=1000, help="micro-batcher deadline")
    p.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "none"])
    p.add_argument("--workload", default="hnsw", choices=["hnsw", "flat", "hybrid", "quantized"],
                   help="hnsw = BASELINE configs[1] (the headline); flat = configs[2] (tensor-core flat top-100); "
                        "hybrid = configs[4] (1536-d HNSW + 10 %% allow-list); quantized = configs[1]'s corpus held "
                        "as int8 (cosine) or float16 (euclidean) rows, SURVEY.md §8 f-4 — see bench_extra.py")
    p.add_argument("--precision", default="int8", choices=["int8", "float16"], help="--workload quantized only")
    p.add_argument("--overlap", type=int, default=3,
                   help="batches in flight: consecutive steps alternate over this many streams / caller threads")
    p.add_argument("--sustain-seconds", type=float, default=2.0,
                   help="length of the sustained sub-measurement reported beside `value` (0 = skip)")
    p.add_argument("--no-extras", action="store_true", help="skip the `extras` (other BASELINE configs / data variants)")
    p.add_argument("--extras", default="auto",
                   help="comma list of flat,hybrid,iid,config3 or 'auto' (N=1: flat,hybrid,iid; N=8: config3)")
    p.add_argument("--config3-n", type=int, default=10_000_000, help="rows of the configs[3] extra (N = 8)")
    return p.parse_args(argv)


# ------------------------------------------------------------------------------------------------
# data + index construction (setup, untimed)
# ------------------------------------------------------------------------------------------------
def make_data(torch, n, dim, latent, noise, seed, device, out=None):
    """Random-normal vectors with low-rank covariance: x = z W + noise * e, z ~ N(0, I_latent).
    latent <= 0 selects i.i.d. isotropic N(0, 1)."""
    g = torch.Generator(device=device)
    if out is None:
        out = torch.empty(n, dim, device=device, dtype=torch.float32)
    step = 1 << 18
    if latent <= 0:
        g.manual_seed(seed)
        for i in range(0, n, step):
            c = min(step, n - i)
            out[i:i + c] = torch.randn(c, dim, generator=g, device=device)
        return out
    g.manual_seed(777)
    W = torch.randn(latent, dim, generator=g, device=device) / latent ** 0.5
    g.manual_seed(seed)
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


def merge_by_distance_then_id(ids, scores, counts, k):
    """Exact merge of per-shard top-k by (distance, id) in numpy — the checker of the merge kernel in the
    parity legs below (ids/scores [S][Q][k], counts [S][Q])."""
    S, Q, kk = ids.shape
    valid = np.arange(kk)[None, None, :] < np.asarray(counts)[:, :, None]
    d = np.where(valid, scores, np.inf).transpose(1, 0, 2).reshape(Q, S * kk)
    i = np.where(valid, ids, np.iinfo(np.uint32).max).transpose(1, 0, 2).reshape(Q, S * kk).astype(np.uint64)
    order = np.lexsort((i, d), axis=1)[:, :k]
    od, oi = np.take_along_axis(d, order, 1), np.take_along_axis(i, order, 1)
    ok = np.isfinite(od)
    return np.where(ok, oi, 0).astype(np.uint32), np.where(ok, od, 0.0), ok.sum(1).astype(np.uint32)


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
            return self
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
        return self

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


def peaks():
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


def workload_string(args):
    """The same string on both arms (the driver compares the two arms' config.workload)."""
    return (f"{args.n}x{args.dim} cosine, HNSW M={args.m} efC={args.efc} efSearch={args.ef}, top-{args.k}, "
            f"batch={args.batch} queries/step")


def shared_config(args):
    """`config` = the workload, identical on both arms (the driver compares the two arms' config); what is specific to
    an arm (how many GPUs / host threads, batches in flight, build time) goes under `arm`."""
    return {"workload": workload_string(args),
            "data_model": (f"random-normal, low-rank covariance (latent {args.latent}, noise {args.noise}), "
                           if args.latent > 0 else "random-normal, i.i.d. isotropic, ") + "seeds 42/4242, levels seed 1",
            "graph": "built on the GPU by kdbgpu_add_batch (bit-identical to the oracle's AddBatch, tests/test_gpu_build.py)",
            "l2_policy": f"inputs larger than L2: {args.n * ((args.dim + 127) // 128 * 128) * 4 / 1e9:.2f} GB corpus, new query batch every step"}


class DeviceRunner:
    """Device-resident stepping: batch i goes to stream i mod n_ov, so that one batch's straggler tail overlaps the
    next one's head.  `search` is GpuIndex.search_device or ShardGroup.search_device (same leading arguments)."""

    def __init__(self, torch, dev, search, Qd, B, k, ef, n_ov):
        self.torch, self.search, self.Qd, self.B, self.k, self.ef, self.n_ov = torch, search, Qd, B, k, ef, n_ov
        self.nb = Qd.shape[0] // B
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_ov)]
        self.ids = [torch.zeros((B, k), dtype=torch.int32, device=dev) for _ in range(n_ov)]
        self.sc = [torch.zeros((B, k), dtype=torch.float64, device=dev) for _ in range(n_ov)]
        self.cnt = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(n_ov)]

    def step(self, i, j=None):
        j = i % self.n_ov if j is None else j
        q = self.Qd[(i % self.nb) * self.B:(i % self.nb + 1) * self.B]
        self.search(q.data_ptr(), self.B, self.k, self.ef, self.ids[j].data_ptr(), self.sc[j].data_ptr(),
                    self.cnt[j].data_ptr(), self.streams[j].cuda_stream)

    def timed(self, first, count, barrier):
        """CUDA events on the launching streams around `count` steps; returns milliseconds."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(self.streams[0])
        for s in self.streams[1:]:
            s.wait_event(ev0)  # every stream starts after the start mark
        for i in range(first, first + count):
            self.step(i)
        for s in self.streams[1:]:
            self.streams[0].wait_stream(s)  # the end mark follows the last kernel of every stream
        ev1.record(self.streams[0])
        barrier()
        return ev0.elapsed_time(ev1)

    def sustained(self, seconds, barrier, chunk=64):
        """Back-to-back steps for at least `seconds` of device time (chunks of `chunk` steps, one event pair)."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(self.streams[0])
        for s in self.streams[1:]:
            s.wait_event(ev0)
        t0, n = time.perf_counter(), 0
        while True:
            for i in range(n, n + chunk):
                self.step(i)
            n += chunk
            for s in self.streams:
                s.synchronize()  # bounds the queue; the device stays busy through the other streams' tails
            if time.perf_counter() - t0 >= seconds:
                break
        for s in self.streams[1:]:
            self.streams[0].wait_stream(s)
        ev1.record(self.streams[0])
        barrier()
        return ev0.elapsed_time(ev1), n


def e2e_threads(torch, local_rank, call, n_threads, first, count):
    """`count` host-buffer calls from n_threads caller threads (ctypes releases the GIL inside the C call)."""
    def worker(j):
        torch.cuda.set_device(local_rank)
        for i in range(first + j, first + count, n_threads):
            call(i)
    ws = [threading.Thread(target=worker, args=(j,)) for j in range(n_threads)]
    t0 = time.perf_counter()
    for t in ws:
        t.start()
    for t in ws:
        t.join()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def one_query_per_call(args, gi, Q, k, ef, B, n_ov):
    """The reference's own call shape (one SearchWithScores per request) through the library's micro-batcher, native
    callers (tools/native): the asynchronous form a Go shim uses, and one blocked OS thread per query."""
    from kektordb_b200 import Batcher
    from tools.native import driver
    W = args.warmup * B
    out = {}
    bt = Batcher(gi, max_batch=B, max_wait_us=args.batcher_wait_us)
    try:
        window = n_ov * B
        driver.run_async(bt, Q[:W], k, ef, args.submitters, window)  # warm-up
        # the region is short (steps x batch queries, ~0.1 s): three repetitions, the median is reported (one
        # descheduled thread on the host otherwise decides the number), all three are kept
        reps = []
        for _ in range(3):
            s0 = bt.stats()
            ids1, sc1, cnt1, secs = driver.run_async(bt, Q[W:], k, ef, args.submitters, window)
            s1 = bt.stats()
            reps.append((secs, max(1, s1.batches - s0.batches), s1.queries - s0.queries))
        secs, nb, nqd = sorted(reps)[1]
        s0, s1 = None, None
        want = gi.SearchWithScores(Q[W:W + B], k, None, ef)
        out["async_submit_poll_take"] = {
            "value": round((Q.shape[0] - W) / secs, 1), "unit": "queries/s",
            "repetitions_qps": [round((Q.shape[0] - W) / r[0], 1) for r in reps], "os_threads": args.submitters + 1 + 4,
            "submitter_threads": args.submitters, "dispatcher_threads": 1, "batcher_worker_threads": 4,
            "queries_in_flight": window, "mean_batch": round(nqd / nb, 1), "batches": nb,
            "max_wait_us": args.batcher_wait_us,
            "first_batch_equal_to_batched_call": bool(np.array_equal(ids1[:B], want[0]) and np.array_equal(sc1[:B], want[1]))}
        n_callers = args.single_call_threads or n_ov * B
        driver.run_callers(bt, Q[:W], k, ef, n_callers)
        reps = []
        for _ in range(3):
            s0 = bt.stats()
            ids2, sc2, cnt2, secs2 = driver.run_callers(bt, Q[W:], k, ef, n_callers)
            s1 = bt.stats()
            reps.append((secs2, max(1, s1.batches - s0.batches), s1.queries - s0.queries))
        secs2, nb, nqd = sorted(reps)[1]
        out["blocking_one_thread_per_query"] = {
            "value": round((Q.shape[0] - W) / secs2, 1), "unit": "queries/s",
            "repetitions_qps": [round((Q.shape[0] - W) / r[0], 1) for r in reps], "caller_threads": n_callers,
            "mean_batch": round(nqd / nb, 1), "batches": nb,
            "first_batch_equal_to_batched_call": bool(np.array_equal(ids2[:B], want[0]))}
        out["value"] = out["async_submit_poll_take"]["value"]
        out["unit"] = "queries/s"
    finally:
        bt.close()
    return out


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
        this = sys.modules[__name__]
        if world > 1 and args.workload == "flat" and args.impl == "ours":
            if args.data_model == "lowrank" and "--data-model" not in sys.argv:
                args.latent = 0
            return bench_extra.run_flat_sharded(args, torch, this)
        if world > 1 and args.impl == "ours":
            raise SystemExit("--workload hybrid/quantized are single-GPU lines")
        if args.workload == "flat" and args.data_model == "lowrank" and "--data-model" not in sys.argv:
            args.latent = 0  # configs[2] is quoted on plain random-normal vectors; exact search has no recall issue
        fn = {"flat": bench_extra.run_flat, "hybrid": bench_extra.run_hybrid, "quantized": bench_extra.run_quantized}
        line = fn[args.workload](args, torch, this)
        print(json.dumps(line), flush=True)
        return 0
    dist = None
    if world > 1 and args.impl == "ours":
        # keep stdout to the one JSON line: NCCL's version banner / warnings go to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    ncores = len(os.sched_getaffinity(0))
    k, ef, B, D, N = args.k, args.ef, args.batch, args.dim, args.n
    n_steps_total = args.warmup + args.steps
    multi = world > 1 and args.impl == "ours"
    do_replica = (not multi) or args.mode in ("both", "replica")
    do_shard = multi and args.mode in ("both", "shard")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): corpus, graph, queries -------------------------------------------------
    X = make_data(torch, N, D, args.latent, args.noise, 42, dev)
    gi, build_s = (None, 0.0)
    if do_replica:
        gi, build_s = build_index(torch, GpuIndex, X, args.m, args.efc, args.build_batch, 1, local_rank)
    # queries: replicas answer different queries per rank, shards answer the same ones
    Qd = make_data(torch, n_steps_total * B, D, args.latent, args.noise, 4242 + (rank if do_replica else 0), dev)
    Qh = torch.empty((n_steps_total * B, D), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()
    Qh_np = Qh.numpy()

    if args.impl == "reference":
        return run_reference_arm(args, gi, Qh_np, ncores, build_s)

    line = {"metric": METRIC, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    pk = peaks()
    stride = (D + 127) // 128 * 128
    n_ov = max(1, min(4, args.overlap))
    clocks = None
    if do_replica:
        # ---- ground truth + recall (untimed) -----------------------------------------------------
        n_gt = min(256, B)
        gi.prepare_search(B, k, ef)  # every launch workspace allocated up front (no allocation inside a timed region)
        ids0, sc0, cnt0, st0 = gi.SearchWithScores(Qh_np[:B], k, None, ef)  # also the first warm-up
        gt_ids, gt_sc, _, _ = gi.flat_search(Qh_np[:n_gt], k, 1, prefilter=True)
        recall_local = recall_at_k(ids0[:n_gt], gt_ids)

        # ---- device-resident timing (`value`) ----------------------------------------------------
        run = DeviceRunner(torch, dev, gi.search_device, Qd, B, k, ef, n_ov)
        for i in range(args.warmup):
            run.step(i)
        barrier()
        sampler = ClockSampler(local_rank, args.clock_sampler)
        if rank == 0:
            sampler.start()
        dev_ms = run.timed(args.warmup, args.steps, barrier)
        # counters of the last launch stand for the per-step work (same graph, i.i.d. query batches)
        st_last = gi.last_search_stats()
        tot_e, tot_h, tot_h0 = st_last.dist_evals, st_last.hops, st_last.hops_l0
        sus_ms, sus_steps = (0.0, 0)
        if args.sustain_seconds > 0:
            sus_ms, sus_steps = run.sustained(args.sustain_seconds, barrier)

        # ---- end-to-end timing through the C ABI with host buffers (`e2e`) -----------------------
        for i in range(args.warmup):
            gi.SearchWithScores(Qh_np[i * B:(i + 1) * B], k, None, ef)
        barrier()
        # K steps per repetition, three repetitions, the median counts (the region is ~70 ms: one descheduled caller
        # thread otherwise decides it); all three are reported
        e2e_reps = [e2e_threads(torch, local_rank, lambda i: gi.SearchWithScores(Qh_np[i * B:(i + 1) * B], k, None, ef),
                                n_ov, args.warmup, args.steps) for _ in range(3)]
        e2e_s = sorted(e2e_reps)[1]
        # the same calls from PAGEABLE host memory — what a Go []float32 is (cgo passes &slice[0])
        Qpage = np.array(Qh_np, copy=True)
        barrier()
        e2e_page_s = e2e_threads(torch, local_rank, lambda i: gi.SearchWithScores(Qpage[i * B:(i + 1) * B], k, None, ef),
                                 n_ov, args.warmup, args.steps)
        del Qpage
        e2e_d2h = B * k * 12 + B * 4 + 40
        # ---- the reference's own call shape: one call per query through the micro-batcher ----------
        one_call = None
        if world == 1 and not args.no_single_call:
            try:
                one_call = one_query_per_call(args, gi, Qh_np, k, ef, B, n_ov)
            except Exception as ex:  # the main line must still print
                one_call = {"error": repr(ex)}
        barrier()
        clocks = sampler.stop() if rank == 0 else None

        # ---- reduce over ranks: max time, min recall ---------------------------------------------
        times = torch.tensor([dev_ms, e2e_s * 1e3, e2e_page_s * 1e3, sus_ms / max(1, sus_steps)], dtype=torch.float64, device=dev)
        rec = torch.tensor([recall_local], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(rec, op=dist.ReduceOp.MIN)
        dev_ms, e2e_ms, e2e_page_ms, sus_step_ms = (float(x) for x in times)
        recall = float(rec[0])
        qps_step = B * world
        value = qps_step * args.steps / (dev_ms / 1e3)

        # ---- roofline of the dominant kernel (hnsw_search_kernel) ---------------------------------
        bytes_per_launch = tot_e * stride * 4 + tot_h0 * (2 * args.m) * 4 + (tot_h - tot_h0) * args.m * 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(run.streams[0])
        nrep = 5
        for r in range(nrep):  # ONE launch at a time: its own duration, straggler tail included
            run.step(r, 0)
        e1.record(run.streams[0])
        torch.cuda.synchronize()
        kernel_ms = e0.elapsed_time(e1) / nrep
        step_ms = dev_ms / args.steps
        achieved = bytes_per_launch / (step_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "hnsw_search_kernel", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
                    "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": None, "peak_source": pk["src"],
                    "algorithmic_bytes_per_launch": int(bytes_per_launch),
                    "step_interval_ms": round(step_ms, 4),
                    "step_interval_is": f"timed region / steps with {n_ov} launches in flight (a throughput interval, "
                                        "not one kernel's duration; that is isolated_launch_ms)",
                    "isolated_launch_ms": round(kernel_ms, 4),
                    "isolated_frac": round(bytes_per_launch / (kernel_ms / 1e3) / 1e9 / pk["hbm_gbs"], 4),
                    "dist_evals_per_query": round(tot_e / B, 1), "hops_per_query": round(tot_h / B, 1)}
        if sus_steps:
            roofline["sustained_frac"] = round(bytes_per_launch / (sus_step_ms / 1e3) / 1e9 / pk["hbm_gbs"], 4)
        prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = "committed ncu --set full capture of this kernel on this workload (" + \
                    str(tj.get("source", "profiles/")) + "), not measured in this run"
            except Exception:
                pass
        line.update({
            "value": round(value, 1), "ms_per_step": round(step_ms, 4), "scaling": "weak", "recall_at_10": round(recall, 4),
            "sustained": None if not sus_steps else {
                "value": round(qps_step / (sus_step_ms / 1e3), 1), "unit": "queries/s", "steps": sus_steps,
                "seconds": round(sus_step_ms * sus_steps / 1e3, 3), "ms_per_step": round(sus_step_ms, 4)},
            "e2e": {"value": round(qps_step * args.steps / (e2e_ms / 1e3), 1), "unit": "queries/s",
                    "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": e2e_d2h,
                    "ms_per_step": round(e2e_ms / args.steps, 4), "host_buffers": "pinned",
                    "repetitions_qps_this_rank": [round(B * args.steps / r, 1) for r in e2e_reps],
                    "value_is": "median of 3 repetitions of `steps` calls (max over ranks)",
                    "pageable_host_buffers": {"value": round(qps_step * args.steps / (e2e_page_ms / 1e3), 1),
                                              "unit": "queries/s", "ms_per_step": round(e2e_page_ms / args.steps, 4)},
                    "one_query_per_call": one_call},
            "gpu_launches": 2 * args.steps, "roofline": roofline})

    variants = []
    if do_replica and world == 1 and not args.no_extras:
        try:
            variants.append(extra_corpus_queries(args, torch, gi, X, dev))
        except Exception as ex:
            variants.append({"name": "queries sampled from the corpus", "error": repr(ex)})

    # ---- the north-star's multi-GPU layout: id-range shards + one NCCL all-gather per batch -------
    shard = None
    if do_shard:
        try:
            shard = run_shard_section(args, torch, dist, ffi, X, Qd, Qh, Qh_np, N, rank, world, local_rank, dev, barrier,
                                      ncores, n_total=n_steps_total)
        except Exception as ex:
            shard = {"error": repr(ex)}
        if not do_replica and isinstance(shard, dict) and "value" in shard:
            line.update({"value": shard["value"], "ms_per_step": shard["ms_per_step"], "scaling": "strong",
                         "recall_at_10": shard["recall_at_10"], "e2e": shard["e2e"], "gpu_launches": 4 * args.steps,
                         "roofline": shard.get("roofline")})
    del X

    # ---- CPU baseline (rank 0, N=1): the oracle port on this box's host cores -----------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline, parity = run_cpu_baseline(args, gi, Qh_np, ncores, ids0, sc0)
        except Exception as ex:  # the main line must still print
            cpu_baseline = {"error": repr(ex)}

    # ---- extras: the other BASELINE configs and data variants, compact ------------------------------
    extras = None
    if not args.no_extras:
        try:
            extras = run_extras(args, torch, dist, rank, world, local_rank, dev, barrier, first=variants)
        except Exception as ex:
            extras = [{"error": repr(ex)}]

    if rank == 0:
        line.update({
            "config": shared_config(args),
            "arm": {"parallelism": "1 GPU" if world == 1 else
                                   (f"`value`: {world} replicas (full corpus per GPU), queries split, NO collective; " if do_replica else "") +
                                   (f"`shard`: corpus split by id range over {world} GPUs, library shard group: ONE ncclAllGather "
                                    f"of the packed per-shard top-{k} per batch + merge kernel" if do_shard else ""),
                    "per_rank": "1xB200 per rank, batch per GPU per step" if world > 1 else "1xB200",
                    "batches_in_flight": n_ov, "build_seconds": round(build_s, 2), "host_cores": ncores},
            "shard": shard, "cpu_baseline": cpu_baseline, "parity": parity, "extras": extras, "clocks": clocks})
        print(json.dumps(line), flush=True)
    if gi is not None:
        gi.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_shard_section(args, torch, dist, ffi, X, Qd, Qh, Qh_np, N, rank, world, local_rank, dev, barrier, ncores, n_total,
                      parity_queries=128):
    """shard g owns ids [g N/G, (g+1) N/G) and its own HNSW over them; every rank answers the SAME
    queries through the library's shard group."""
    from kektordb_b200 import GpuIndex
    from kektordb_b200 import sharding
    k, ef, B, D = args.k, args.ef, args.batch, args.dim
    base, n_local = sharding.shard_range(N, world, rank)
    gs, build_s = build_index(torch, GpuIndex, X[base:base + n_local], args.m, args.efc, args.build_batch, 1 + rank, local_rank)
    uid = [sharding.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)  # plumbing: carries the 128-byte NCCL id
    grp = sharding.ShardGroup.rank(gs, rank, world, uid[0], base)
    gs.prepare_search(B, k, ef)
    # shards answer the same queries on every rank: use rank 0's
    if Qd is not None:
        dist.broadcast(Qd, src=0)
        Qh.copy_(Qd)
        torch.cuda.synchronize()
    n_ov = 4
    run = DeviceRunner(torch, dev, grp.search_device, Qd, B, k, ef, n_ov)
    for i in range(args.warmup):
        run.step(i)
    grp.sync()
    barrier()
    dev_ms = run.timed(args.warmup, args.steps, barrier)
    st = grp.sync()
    sus_ms, sus_steps = (0.0, 0)
    if args.sustain_seconds > 0:
        sus_ms, sus_steps = run.sustained(args.sustain_seconds, barrier)
        grp.sync()
    # one batch at a time: the latency of a batch (traversal + all-gather + merge), and its parts
    iso = []
    for r in range(3):
        ids_i, sc_i, cnt_i, s_i = grp.SearchWithScores(Qh_np[r * B:(r + 1) * B], k, None, ef)
        iso.append(s_i)
    barrier()
    # e2e: host buffers in, merged result out, 4 batches in flight from ONE thread (collectives stay in issue order)
    def e2e_run(first, count):
        tickets = []
        t0 = time.perf_counter()
        for i in range(first, first + count):
            if len(tickets) == 4:
                grp.wait(tickets.pop(0))
            tickets.append(grp.submit(Qh_np[i * B:(i + 1) * B], k, ef))
        while tickets:
            last = grp.wait(tickets.pop(0))
        torch.cuda.synchronize()
        return time.perf_counter() - t0, last
    e2e_run(0, args.warmup)
    barrier()
    e2e_s, _ = e2e_run(args.warmup, args.steps)
    barrier()
    # recall of the merged answer against the merged exact scan (sharded flat scan: same exchange + merge)
    n_gt = min(256, B)
    got = grp.SearchWithScores(Qh_np[:B], k, None, ef)
    gt = grp.flat_search(Qh_np[:n_gt], k, 1, prefilter=True)
    recall = recall_at_k(got[0][:n_gt], gt[0])
    # parity: G oracle indexes (each rank's own GPU-built graph, searched on the CPU) + exact merge, vs the group
    parity = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        npar = min(parity_queries, B)
        oi = oracle_from_gpu(gs, args.m, args.efc, O.ARITH_KERNEL)
        pid, psc, pcnt, pst = oi.search_batch(Qh_np[:npar], k, ef, threads=max(1, ncores // world))
        del oi
        gl = sharding.globalize_ids(pid, pcnt, base)
        parts = [None] * world
        dist.all_gather_object(parts, (gl, psc, pcnt.astype(np.uint32)))
        w_ids, w_sc, w_cnt = merge_by_distance_then_id(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]),
                                                       np.stack([p[2] for p in parts]), k)
        parity = {"oracle": f"{world} oracle indexes (kernel-order arithmetic) over the same GPU-built graphs + exact merge "
                            "by (distance, id)", "queries": npar,
                  "ids_equal": bool(np.array_equal(w_ids, got[0][:npar])),
                  "scores_bit_equal": bool(np.array_equal(w_sc, got[1][:npar])),
                  "counts_equal": bool(np.array_equal(w_cnt, got[2][:npar]))}
    t = torch.tensor([dev_ms, e2e_s * 1e3, sus_ms / max(1, sus_steps), build_s], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, sus_step_ms, build_s = (float(x) for x in t)
    pk = peaks()
    stride = (D + 127) // 128 * 128
    byts = st.dist_evals * stride * 4 + st.hops_l0 * 2 * args.m * 4 + (st.hops - st.hops_l0) * args.m * 4  # all shards
    step_ms = dev_ms / args.steps
    out = {
        "what": f"corpus split by id range over {world} GPUs ({n_local} ids and one HNSW per GPU), every query answered by "
                f"every shard; kdbgpu_shard_search_*: traversal (epilogue writes global ids into one packed buffer) -> ONE "
                f"ncclAllGather per batch on a high-priority stream -> merge kernel; {n_ov} batches in flight",
        "value": round(B * args.steps / (dev_ms / 1e3), 1), "unit": "queries/s", "ms_per_step": round(step_ms, 4),
        "scaling": "strong", "recall_at_10": round(recall, 4),
        "sustained": None if not sus_steps else {"value": round(B / (sus_step_ms / 1e3), 1), "unit": "queries/s",
                                                 "steps": sus_steps, "seconds": round(sus_step_ms * sus_steps / 1e3, 3)},
        "e2e": {"value": round(B * args.steps / (e2e_ms / 1e3), 1), "unit": "queries/s", "h2d_bytes_per_step": B * D * 4,
                "d2h_bytes_per_step": B * k * 12 + B * 4 + 48, "ms_per_step": round(e2e_ms / args.steps, 4),
                "host_buffers": "pinned", "batches_in_flight": 4},
        "one_batch_alone": {"traversal_ms": round(float(np.median([s.traversal_ms for s in iso])), 4),
                            "allgather_ms": round(float(np.median([s.exchange_ms for s in iso])), 4),
                            "merge_ms": round(float(np.median([s.merge_ms for s in iso])), 4),
                            "total_ms_incl_copies": round(float(np.median([s.total_ms for s in iso])), 4)},
        "allgather_ms": round(float(np.median([s.exchange_ms for s in iso])), 4),
        "merge_ms": round(float(np.median([s.merge_ms for s in iso])), 4),
        "allgather_bytes_per_rank": int(B * k * 12 + B * 4 + 48),
        "collective_share_of_step": round(float(np.median([s.exchange_ms + s.merge_ms for s in iso])) / step_ms, 4),
        "parity_vs_oracle": parity, "build_seconds": round(build_s, 2),
        "roofline": {"bound": "hbm", "kernel": "hnsw_search_kernel (per shard)", "unit": "GB/s", "peak": pk["hbm_gbs"],
                     "achieved": round(byts / world / (step_ms / 1e3) / 1e9, 1),
                     "frac": round(byts / world / (step_ms / 1e3) / 1e9 / pk["hbm_gbs"], 4), "traffic": None,
                     "dist_evals_per_query_all_shards": round(st.dist_evals / B, 1)},
    }
    grp.close()
    gs.close()
    return out


def compact(line):
    """The fields of a full bench line an `extras` entry keeps."""
    if not isinstance(line, dict) or "value" not in line:
        return line
    rf = line.get("roofline") or {}
    keep = {"metric": line.get("metric"), "value": line.get("value"), "unit": line.get("unit"),
            "ms_per_step": line.get("ms_per_step"), "steps": line.get("steps"),
            "e2e": {kk: (line.get("e2e") or {}).get(kk) for kk in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "roofline": {kk: rf.get(kk) for kk in ("bound", "kernel", "achieved", "peak", "unit", "frac", "isolated_frac") if kk in rf},
            "cpu_baseline": line.get("cpu_baseline"), "parity": line.get("parity"), "clocks": line.get("clocks"),
            "workload": (line.get("config") or {}).get("workload")}
    for kk in ("recall_at_10", "recall_at_k", "n_gpus", "scaling"):
        if kk in line:
            keep[kk] = line[kk]
    return keep


def run_extras(args, torch, dist, rank, world, local_rank, dev, barrier, first=None):
    """Compact lines for the other BASELINE configs and the data variants BASELINE.md §3 names, each time-boxed."""
    import copy
    import bench_extra
    this = sys.modules[__name__]
    want = args.extras.split(",") if args.extras != "auto" else (
        ["flat", "hybrid", "iid"] if world == 1 else (["config3"] if world == 8 else []))
    out = list(first or [])
    for name in want:
        t0 = time.time()
        try:
            if name in ("flat", "hybrid") and world == 1:
                a = copy.copy(args)
                a.cpu_seconds = min(args.cpu_seconds, 6.0)
                a.no_single_call = True
                a.steps = max(args.steps, 150 if name == "flat" else 100)  # >= 0.2 s timed: several clock samples
                if name == "flat":
                    a.latent = 0  # configs[2] is quoted on plain random-normal vectors
                line = (bench_extra.run_flat if name == "flat" else bench_extra.run_hybrid)(a, torch, this)
                e = compact(line)
                e["name"] = "configs[2] flat top-100" if name == "flat" else "configs[4] hybrid (1536-d + 10% allow-list)"
                out.append(e)
            elif name == "iid" and world == 1:
                out.append(extra_iid(args, torch, dev, local_rank))
            elif name == "config3" and world > 1:
                e = extra_config3(args, torch, dist, rank, world, local_rank, dev, barrier)
                if rank == 0:
                    out.append(e)
        except Exception as ex:
            out.append({"name": name, "error": repr(ex)})
        if out and isinstance(out[-1], dict):
            out[-1]["wall_seconds"] = round(time.time() - t0, 1)
        torch.cuda.empty_cache()
    return out


def extra_corpus_queries(args, torch, gi, X, dev):
    """The reference's own benchmark samples its queries FROM the indexed dataset (clients/python/benchmark2.py:443-445)."""
    k, ef, B = args.k, args.ef, args.batch
    nb = 8
    idx = torch.from_numpy(np.random.default_rng(99).integers(0, X.shape[0], nb * B)).to(dev)
    Qall = X[idx].contiguous()
    Qc = Qall[:B].cpu().numpy()
    ids, sc, cnt, _ = gi.SearchWithScores(Qc, k, None, ef)
    gt, _, _, _ = gi.flat_search(Qc[:256], k, 1, prefilter=True)
    run = DeviceRunner(torch, dev, gi.search_device, Qall, B, k, ef, max(1, min(4, args.overlap)))
    for i in range(3):
        run.step(i)
    torch.cuda.synchronize()
    ms = run.timed(0, 16, torch.cuda.synchronize)
    return {"name": "queries sampled from the corpus (benchmark2.py:443-445)", "value": round(16 * B / (ms / 1e3), 1),
            "unit": "queries/s", "steps": 16, "recall_at_10": round(recall_at_k(ids[:256], gt), 4),
            "self_is_top1": round(float(np.mean(ids[:, 0] == (idx[:B].cpu().numpy() + 1))), 4),
            "note": "same index as the headline; a query that IS a stored row finds itself at distance ~0"}


def extra_iid(args, torch, dev, local_rank):
    """BASELINE.md §3's literal data model: corpus and queries i.i.d. N(0,1)."""
    from kektordb_b200 import GpuIndex
    k, ef, B, D, N = args.k, args.ef, args.batch, args.dim, args.n
    X = make_data(torch, N, D, 0, 0.0, 42, dev)
    g2, build_s = build_index(torch, GpuIndex, X, args.m, args.efc, args.build_batch, 1, local_rank)
    del X
    nb = 8
    Q = make_data(torch, nb * B, D, 0, 0.0, 4242, dev)
    Qh = Q[:B].cpu().numpy()
    g2.prepare_search(B, k, ef)
    ids, sc, cnt, st = g2.SearchWithScores(Qh, k, None, ef)
    gt, _, _, _ = g2.flat_search(Qh[:256], k, 1, prefilter=True)
    run = DeviceRunner(torch, dev, g2.search_device, Q, B, k, ef, max(1, min(4, args.overlap)))
    for i in range(3):
        run.step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank, args.clock_sampler).start()
    ms = run.timed(0, 16, torch.cuda.synchronize)
    clocks = sampler.stop()
    stl = g2.last_search_stats()
    stride = (D + 127) // 128 * 128
    byts = stl.dist_evals * stride * 4 + stl.hops_l0 * 2 * args.m * 4 + (stl.hops - stl.hops_l0) * args.m * 4
    pk = peaks()
    g2.close()
    return {"name": "i.i.d. N(0,1) corpus and queries (BASELINE.md §3 data model)", "value": round(16 * B / (ms / 1e3), 1),
            "unit": "queries/s", "steps": 16, "recall_at_10": round(recall_at_k(ids[:256], gt), 4),
            "dist_evals_per_query": round(st.dist_evals / B, 1), "build_seconds": round(build_s, 1),
            "roofline": {"bound": "hbm", "achieved": round(byts / (ms / 16 / 1e3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": round(byts / (ms / 16 / 1e3) / 1e9 / pk["hbm_gbs"], 4)},
            "clocks": clocks,
            "note": "isotropic 768-d data has no neighbourhood structure: recall collapses for ANY HNSW (DESIGN.md §6); "
                    "CPU and GPU return the same bits"}


def extra_config3(args, torch, dist, rank, world, local_rank, dev, barrier):
    """BASELINE configs[3]: 10M x 768 cosine, corpus sharded over 8 GPUs, NCCL top-k merge."""
    import copy
    from kektordb_b200 import ffi
    a = copy.copy(args)
    a.n = args.config3_n
    a.no_cpu_baseline = True  # a 1.25M-row oracle per rank is minutes of CPU; parity at this shape is the shard section's
    a.sustain_seconds = min(args.sustain_seconds, 1.0)
    from kektordb_b200.sharding import shard_range
    base, n_local = shard_range(a.n, world, rank)
    # only the local shard's rows are ever materialised (30.7 GB in total)
    Xl = torch.empty((n_local, a.dim), dtype=torch.float32, device=dev)
    full_like = make_data_range(torch, a.n, a.dim, a.latent, a.noise, 42, dev, base, n_local, Xl)
    n_total = a.warmup + a.steps
    Qd = make_data(torch, n_total * a.batch, a.dim, a.latent, a.noise, 4242, dev)
    Qh = torch.empty((n_total * a.batch, a.dim), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Qd)
    torch.cuda.synchronize()

    class _Rows:  # run_shard_section slices X[base:base+n_local]
        def __getitem__(self, s):
            return full_like
    ncores = len(os.sched_getaffinity(0))
    e = run_shard_section(a, torch, dist, ffi, _Rows(), Qd, Qh, Qh.numpy(), a.n, rank, world, local_rank, dev, barrier, ncores,
                          n_total)
    e["name"] = f"configs[3] {a.n}x{a.dim} cosine sharded over {world} GPUs"
    e["workload"] = f"{a.n}x{a.dim} cosine, per-shard HNSW M={a.m} efC={a.efc} efSearch={a.ef}, top-{a.k}, batch={a.batch}"
    return e


def make_data_range(torch, n, dim, latent, noise, seed, device, base, count, out):
    """Rows [base, base+count) of make_data(n, ...) without materialising the rest (same generator stream)."""
    g = torch.Generator(device=device)
    step = 1 << 18
    W = None
    if latent > 0:
        g.manual_seed(777)
        W = torch.randn(latent, dim, generator=g, device=device) / latent ** 0.5
    g.manual_seed(seed)
    for i in range(0, n, step):
        c = min(step, n - i)
        if latent > 0:
            z = torch.randn(c, latent, generator=g, device=device)
            e = torch.randn(c, dim, generator=g, device=device)
        else:
            z, e = None, torch.randn(c, dim, generator=g, device=device)
        lo, hi = max(i, base), min(i + c, base + count)
        if lo < hi:
            blk = (z[lo - i:hi - i] @ W + noise * e[lo - i:hi - i]) if latent > 0 else e[lo - i:hi - i]
            out[lo - base:hi - base] = blk
        if i + c >= base + count:
            break
    return out


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
    for name, arith in (("avx2", O.ARITH_AVX2), ("seq", O.ARITH_SEQ)):
        oi.set_arith(arith)
        rid, rsc, _, _ = oi.search_batch(Qh_np[:npar], k, ef, threads=ncores)  # avx2: also the CPU warm-up
        parity[f"topk_set_agreement_vs_{name}_order"] = round(float(np.mean(
            [len(set(rid[i].tolist()) & set(gpu_ids0[i].tolist())) / k for i in range(npar)])), 4)
        same = rid == gpu_ids0[:npar]
        parity[f"max_abs_score_diff_vs_{name}_order"] = float(np.max(np.abs(rsc[same] - gpu_sc0[:npar][same]))) if same.any() else None
    oi.set_arith(O.ARITH_AVX2)
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
    gt_ids, _, _, _ = gi.flat_search(Qh_np[:min(256, B)], k, 1, prefilter=True)
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
        "impl": "reference", "metric": METRIC,
        "value": round(value, 1), "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "recall_at_10": round(recall, 4),
        "config": shared_config(args),
        "arm": {"parallelism": f"CPU only, {ncores} threads (the graph is searched on the CPU only)",
                "build_seconds": round(build_s, 2), "host_cores": ncores},
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
