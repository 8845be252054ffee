// api.cu — the C ABI of include/kektordb_gpu.h: handle lifecycle, staging of the corpus and the
// graph into HBM, and the batched query entry points.  No torch types, no exceptions across the
// boundary, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <new>
#include <string>
#include <vector>

#include "handle.h"

using namespace kdb;

namespace kdb {
struct BuildLaunch {
  uint32_t start_id, count;
  uint32_t pre_entry;
  int pre_max;
  int efc;
  uint32_t n_slots;
  uint32_t up_base;
  uint32_t n_rows;
  const uint32_t *out_off;
  const uint32_t *slot_node;
  const uint8_t *slot_level;
  uint32_t *cand_ids, *cand_cnt;
  uint32_t *row_cnt, *row_off, *srcs, *active, *n_active, *work_counter;
  uint32_t *adj0, *upper_adj;
  const uint32_t *upper_node;
  const uint8_t *upper_level;
  double *scratch_d;
  uint32_t *scratch_ids;
  uint32_t scratch_cap;
  int commit_grid;
};
size_t seq_add_smem_bytes(const DevIndex &ix, int efc, uint32_t cand_smem);
int build_search_occupancy(const DevIndex &ix, int efc, uint32_t cand_smem);
cudaError_t launch_add_batch(const DevIndex &ix, const SearchArgs &a, const BuildLaunch &L, int search_grid,
                             cudaStream_t stream);
cudaError_t launch_seq_add(const DevIndex &ix, const SearchArgs &a, uint32_t start_id, uint32_t count, int efc,
                           uint32_t *adj0, uint32_t *upper_adj, uint32_t *entry_io, cudaStream_t stream);
}  // namespace kdb

namespace kdb {
int arena_check_header(const unsigned char *hdr, uint32_t dim, int precision, const char *what);
size_t arena_chunk_size();
size_t arena_header_size();
uint32_t arena_vecs_per_chunk(uint32_t vector_bytes);
int arena_max_chunk(const char *dir, std::string *err);
long arena_read_chunk(const char *dir, int id, unsigned char *hdr, unsigned char *dst, size_t want, std::string *err);
}  // namespace kdb

static thread_local std::string g_last_error;

namespace kdb {
int set_error(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
int arena_fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace kdb

#define fail kdb::set_error
// Nothing unwinds across the C boundary: the entry points (and the helpers behind them) that size host arrays from the
// caller's arguments run inside this pair.  RAII guards (workspace release, add_batch's rollback) have run by then.
#define KDB_NOTHROW_BEGIN try {
#define KDB_NOTHROW_END                                                             \
  }                                                                                 \
  catch (const std::bad_alloc &) {                                                  \
    return fail(KDBGPU_ERR_NOMEM, "out of host memory");                            \
  }                                                                                 \
  catch (...) {                                                                     \
    return fail(KDBGPU_ERR_CUDA, "unexpected C++ exception inside the library");    \
  }

namespace {

int ensure_search_workspace(kdbgpu_index *h, int grid) {
  const uint32_t words = (((h->capacity + 1 + 31) / 32) + 3) & ~3u;
  if (grid > h->ws_grid || words != h->vis_words) {
    h->vis_words = words;
    h->visited.release();
    CUDA_TRY(h->visited.reserve((size_t)grid * words, true));
    h->cand_overflow.release();
    CUDA_TRY(h->cand_overflow.reserve((size_t)grid * h->ovf_cap));
    h->ws_grid = grid;
  }
  CUDA_TRY(h->stats.reserve(4, true));
  CUDA_TRY(h->work_counter.reserve(1, true));
  CUDA_TRY(h->err_flag.reserve(1, true));
  return KDBGPU_OK;
}

}  // namespace

namespace kdb {

// ---- per-launch search workspaces ---------------------------------------------------------------
int ensure_ws(kdbgpu_index *h, kdbgpu_index::SearchWs &w, int grid) {
  const uint32_t words = (((h->capacity + 1 + 31) / 32) + 3) & ~3u;
  if (grid > w.grid || words != w.vis_words) {
    w.vis_words = words;
    w.visited.release();
    CUDA_TRY(w.visited.reserve((size_t)grid * words, true));
    w.cand_overflow.release();
    CUDA_TRY(w.cand_overflow.reserve((size_t)grid * h->ovf_cap));
    w.grid = grid;
  }
  CUDA_TRY(w.stats.reserve(4, true));
  CUDA_TRY(w.work_counter.reserve(4, true));  // [0] fast kernel, [1] exact kernel, [2] queries handed over
  CUDA_TRY(w.err_flag.reserve(1, true));
  return KDBGPU_OK;
}

// blocks until one of the workspaces is free (host-buffer path: at most kNumSearchWs batches in flight)
int acquire_ws(kdbgpu_index *h) {
  std::unique_lock<std::mutex> lk(h->ws_mu);
  for (;;) {
    for (int i = 0; i < kdbgpu_index::kNumSearchWs; ++i) {
      const int j = (int)((h->ws_next + (unsigned)i) % kdbgpu_index::kNumSearchWs);
      if (!h->sws[j].busy) {
        h->sws[j].busy = true;
        h->ws_next = (unsigned)j + 1;
        return j;
      }
    }
    h->ws_cv.wait(lk);
  }
}
void release_ws(kdbgpu_index *h, int j) {
  {
    std::lock_guard<std::mutex> lk(h->ws_mu);
    h->sws[j].busy = false;
  }
  h->ws_cv.notify_one();
}

// queue one traversal launch on `stream` using workspace `w`; buffers are device pointers.  The
// launch waits for the previous user of the workspace and records its own completion on w.done.
int enqueue_search(kdbgpu_index *h, kdbgpu_index::SearchWs &w, const float *d_q_prepared, uint32_t nq, int k, int ef,
                   const uint32_t *d_allow, uint32_t allow_entry, uint32_t *d_ids, double *d_scores,
                   uint32_t *d_counts, cudaStream_t stream, unsigned long long *d_stats, int *d_err,
                   uint32_t id_base) {
  DevIndex ix = h->dev();
  // the launch shape: the throughput shape while other batches of this handle are in flight (their query-warps
  // fill this launch's straggler tail), the latency shape when this batch is alone on the device
  SearchTuning tn = h->tuning;
  if (tn.slots_idle > 0 && tn.slots_idle != tn.slots) {
    // "in flight" = a launch of another workspace, queued on ANOTHER stream (launches of one stream run one after
    // the other: nothing fills their tails), that has not finished on the device — every path records the
    // workspace's `done` event behind its launch
    int n_other = 0;
    {
      std::lock_guard<std::mutex> lk(h->ws_mu);
      for (int i = 0; i < kdbgpu_index::kNumSearchWs; ++i) {
        kdbgpu_index::SearchWs &o = h->sws[i];
        if (&o == &w || o.launch_stream == nullptr || o.launch_stream == stream) continue;
        if (o.done && cudaEventQuery(o.done) == cudaErrorNotReady) n_other++;
      }
      w.launch_stream = stream;
    }
    (void)cudaGetLastError();  // cudaErrorNotReady is an answer, not a failure
    if (n_other == 0) {
      tn.slots = tn.slots_idle;
      // a lone batch small enough to be resident in ONE wave of the next larger shape takes it: only the per-query
      // time counts then, and twice the rows in flight per query-warp shorten a hop (1 M x 768 float32, one call at a
      // time: 2.39 -> 2.20 ms for 1 query, 3.24 -> 3.06 ms for 512; profiles/r2_latency_by_batch.json)
      SearchTuning t2 = tn;
      t2.slots = tn.slots * 2;
      if (search_slots_supported(t2.slots)) {
        const int o2 = search_occupancy(ix, ef, t2);
        if (o2 > 0 && (uint64_t)nq <= (uint64_t)o2 * (uint64_t)h->num_sms) tn = t2;
      }
    }
  }
  int occ = search_occupancy(ix, ef, tn);
  if (occ <= 0 && tn.slots != h->tuning.slots) {  // the latency shape does not fit: fall back
    tn.slots = h->tuning.slots;
    occ = search_occupancy(ix, ef, tn);
  }
  if (occ <= 0)
    return fail(KDBGPU_ERR_INVALID, "search configuration does not fit shared memory (dim=%d ef=%d smem=%zu)", h->dim,
                ef, search_smem_bytes(ix, ef, tn));
  const bool fast = search_fast_eligible(ix, ef, tn);
  const int occ_fast = fast ? search_fast_occupancy(ix, ef, tn) : 0;
  const int occ_max = occ_fast > occ ? occ_fast : occ;
  int rc = ensure_ws(h, w, occ_max * h->num_sms);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(w.work_counter.p, 0, 4 * sizeof(uint32_t), stream));
  if (d_stats && d_err) {  // contiguous in the caller's blob: stats[4] then err
    CUDA_TRY(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long) + sizeof(long long), stream));
  } else {
    d_stats = w.stats.p;
    d_err = w.err_flag.p;
    CUDA_TRY(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), stream));
    CUDA_TRY(cudaMemsetAsync(d_err, 0, sizeof(int), stream));
  }
  SearchArgs a{};
  a.queries = d_q_prepared;
  a.qnorms = w.qnorms.p;
  a.nq = nq;
  a.k = k;
  a.ef = ef;
  a.allow = d_allow;
  a.allow_entry = allow_entry;
  a.out_ids = d_ids;
  a.out_scores = d_scores;
  a.out_counts = d_counts;
  a.visited = w.visited.p;
  a.vis_words = w.vis_words;
  a.cand_overflow = w.cand_overflow.p;
  a.ovf_cap = h->ovf_cap;
  a.cand_smem = (uint32_t)tn.cand_smem;
  a.stats = d_stats;
  a.work_counter = w.work_counter.p;
  a.err_flag = d_err;
  a.id_base = id_base;
  a.rows_evict_first = tn.rows_evict_first;
  int grid = occ * h->num_sms;
  if ((uint32_t)grid > nq) grid = (int)nq;
  if (fast && occ_fast > 0) {
    // pass 1: sorted-list fast path; pass 2: the heap path over whatever met a distance tie
    CUDA_TRY(w.redo.reserve(nq));
    int gfast = occ_fast * h->num_sms;
    if ((uint32_t)gfast > nq) gfast = (int)nq;
    a.redo_count = w.work_counter.p + 2;
    if (search_fast_hands_over(ix)) {
      a.redo_list = w.redo.p;
      CUDA_TRY(launch_search_fast(ix, a, tn, gfast, stream));
      a.work_counter = w.work_counter.p + 1;
      a.query_list = w.redo.p;
      a.query_count = w.work_counter.p + 2;
      a.redo_list = nullptr;
      a.redo_count = nullptr;
      CUDA_TRY(launch_search(ix, a, tn, grid, stream));
    } else {  // tied queries are re-run by the heap path inside the same launch
      a.redo_list = nullptr;
      CUDA_TRY(launch_search_fast(ix, a, tn, gfast, stream));
    }
  } else {
    CUDA_TRY(launch_search(ix, a, tn, grid, stream));
  }
  if (fast && occ_fast > 0)  // stats[3] = queries the heap pass answered after a tie in the fast pass
    CUDA_TRY(cudaMemcpyAsync(d_stats + 3, w.work_counter.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  return KDBGPU_OK;
}

// query preparation of searchInternal (hnsw_index.go:401-434) into w.q_prep (+ w.qnorms for int8)
int prepare_queries(kdbgpu_index *h, kdbgpu_index::SearchWs &w, const float *d_q_raw, uint32_t nq, cudaStream_t s) {
  CUDA_TRY(w.q_prep.reserve((size_t)nq * h->stride));
  if (h->precision == KDBGPU_PRECISION_F32) {
    CUDA_TRY(launch_prep_queries(d_q_raw, (size_t)h->dim, w.q_prep.p, nq, (uint32_t)h->dim, h->stride, h->metric, s));
    return KDBGPU_OK;
  }
  if (h->precision == KDBGPU_PRECISION_INT8) {
    if (h->abs_max == 0.f) return fail(KDBGPU_ERR_STATE, "int8 index without a trained quantizer (kdbgpu_set_quantizer)");
    CUDA_TRY(w.qnorms.reserve(nq));
  }
  CUDA_TRY(launch_convert_rows(d_q_raw, (size_t)h->dim, w.q_prep.p, h->stride, nq, (uint32_t)h->dim, h->kind,
                               h->metric == KDBGPU_METRIC_COSINE, h->abs_max, w.qnorms.p, true, s));
  return KDBGPU_OK;
}

uint32_t first_set_bit(const uint64_t *bits, size_t words, bool *found) {
  for (size_t w = 0; w < words; ++w)
    if (bits[w]) {
      *found = true;
      return (uint32_t)(w * 64 + (size_t)__builtin_ctzll(bits[w]));
    }
  *found = false;
  return 0;
}

// copy a host allow-list bitset into `dst`, padded/truncated to cover ids 0..capacity
int stage_allow(kdbgpu_index *h, DevBuf<uint32_t> &dst, const uint64_t *allow, size_t allow_words,
                cudaStream_t stream) {
  const size_t need32 = ((size_t)h->capacity + 1 + 31) / 32 + 2;
  CUDA_TRY(dst.reserve(need32));
  CUDA_TRY(cudaMemsetAsync(dst.p, 0, need32 * sizeof(uint32_t), stream));
  size_t copy32 = allow_words * 2;
  if (copy32 > need32) copy32 = need32;
  CUDA_TRY(cudaMemcpyAsync(dst.p, allow, copy32 * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  return KDBGPU_OK;
}


}  // namespace kdb

namespace {

// ---- flat scan -------------------------------------------------------------------------------------
// Every flat call runs on one of the handle's two flat workspaces (FlatWs: its own stream, staging and result
// buffers), under the handle's SHARED lock — so the copies and host-side work of one call overlap the kernels
// of the next, and flat scans run next to traversals.  Results leave the device through a pinned staging blob
// (one asynchronous copy per array, one synchronisation) or, for device-resident outputs, device-to-device.

bool is_device_pointer(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// results of `c` queries: device buffers of the workspace -> the caller's arrays (+ per-query flags to `flags`, host)
int flat_deliver(kdbgpu_index *h, kdbgpu_index::FlatWs &w, uint32_t c, int k, uint32_t *out_ids, double *out_scores,
                 uint32_t *out_counts, uint32_t *flags, cudaStream_t s) {
  const size_t nk = (size_t)c * k;
  if (is_device_pointer(out_ids)) {
    CUDA_TRY(cudaMemcpyAsync(out_ids, w.out_ids.p, nk * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(out_scores, w.out_scores.p, nk * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(out_counts, w.out_counts.p, (size_t)c * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    if (flags) CUDA_TRY(cudaMemcpyAsync(flags, w.t_flags.p, (size_t)c * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return KDBGPU_OK;
  }
  const size_t o_ids = nk * sizeof(double), o_cnt = o_ids + nk * sizeof(uint32_t), o_flags = o_cnt + (size_t)c * sizeof(uint32_t);
  const size_t bytes = o_flags + (size_t)c * sizeof(uint32_t);
  if (w.h_out_bytes < bytes) {
    if (w.h_out) cudaFreeHost(w.h_out);
    w.h_out = nullptr;
    w.h_out_bytes = 0;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&w.h_out), bytes + bytes / 4, cudaHostAllocDefault));
    w.h_out_bytes = bytes + bytes / 4;
  }
  CUDA_TRY(cudaMemcpyAsync(w.h_out, w.out_scores.p, nk * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(w.h_out + o_ids, w.out_ids.p, nk * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(w.h_out + o_cnt, w.out_counts.p, (size_t)c * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (flags) CUDA_TRY(cudaMemcpyAsync(w.h_out + o_flags, w.t_flags.p, (size_t)c * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(out_scores, w.h_out, nk * sizeof(double));
  memcpy(out_ids, w.h_out + o_ids, nk * sizeof(uint32_t));
  memcpy(out_counts, w.h_out + o_cnt, (size_t)c * sizeof(uint32_t));
  if (flags) memcpy(flags, w.h_out + o_flags, (size_t)c * sizeof(uint32_t));
  (void)h;
  return KDBGPU_OK;
}

// exhaustive float64 scan (flat.cu)
int flat_scan_impl(kdbgpu_index *h, kdbgpu_index::FlatWs &w, const float *queries, uint32_t nq, int k, int mode,
                   const uint32_t *d_allow, uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
  cudaStream_t s = w.stream;
  // bound the dist[chunk][n] workspace to ~2 GiB
  uint32_t chunk = (uint32_t)((2ull << 30) / ((size_t)h->n * sizeof(double)));
  if (chunk < 16) chunk = 16;
  chunk &= ~15u;
  if (chunk > nq) chunk = nq;
  CUDA_TRY(w.q_raw.reserve((size_t)chunk * h->dim));
  CUDA_TRY(w.q_prep.reserve((size_t)chunk * h->stride));
  CUDA_TRY(w.out_ids.reserve((size_t)chunk * k));
  CUDA_TRY(w.out_scores.reserve((size_t)chunk * k));
  CUDA_TRY(w.out_counts.reserve(chunk));
  CUDA_TRY(w.flat_dist.reserve((size_t)chunk * h->n));
  DevIndex ix = h->dev();
  for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
    const uint32_t c = nq - q0 < chunk ? nq - q0 : chunk;
    CUDA_TRY(cudaMemcpyAsync(w.q_raw.p, queries + (size_t)q0 * h->dim, (size_t)c * h->dim * sizeof(float),
                             cudaMemcpyDefault, s));
    CUDA_TRY(launch_prep_queries(w.q_raw.p, (size_t)h->dim, w.q_prep.p, c, (uint32_t)h->dim, h->stride,
                                 mode == 1 ? h->metric : KDBGPU_METRIC_L2, s));
    CUDA_TRY(launch_flat_distances(ix, w.q_raw.p, w.q_prep.p, c, mode, w.flat_dist.p, s));
    CUDA_TRY(launch_flat_select(ix, w.flat_dist.p, c, k, d_allow, w.out_ids.p, w.out_scores.p, w.out_counts.p, s));
    int rc = flat_deliver(h, w, c, k, out_ids + (size_t)q0 * k, out_scores + (size_t)q0 * k, out_counts + q0, nullptr, s);
    if (rc) return rc;
  }
  return KDBGPU_OK;
}

// ---- flat scan with the tensor-core pre-filter (flat_tc.cu) ---------------------------------------
struct TcPlan {
  uint32_t bm, bn, n_pad, dp, n_groups;
  int use_norm;
  float alpha;
};

// bf16 mirror of the rows (built once, lazily, under tc_mu), beta for this call's liveness / allow-list
int tc_prepare(kdbgpu_index *h, kdbgpu_index::FlatWs &w, int mode, const uint32_t *d_allow, TcPlan *P, cudaStream_t s) {
  P->bm = flat_tc_bm();
  P->bn = flat_tc_bn();
  P->n_pad = (h->n + P->bn - 1) / P->bn * P->bn;
  P->dp = ((uint32_t)h->dim + flat_tc_bk() - 1) / flat_tc_bk() * flat_tc_bk();
  P->n_groups = P->n_pad / 32;
  P->use_norm = (mode == 0 || h->metric == KDBGPU_METRIC_L2) ? 1 : 0;
  P->alpha = P->use_norm ? -2.f : -1.f;  // L2: |x|^2 - 2<q,x> (+|q|^2) ; cosine: -<q^,x> (+1)
  {
    std::lock_guard<std::mutex> tl(h->tc_mu);
    if (!h->tc_valid || h->tc_n != h->n) {
      CUDA_TRY(cudaDeviceSynchronize());  // the other flat workspace may still be reading the old mirror
      CUDA_TRY(h->x_bf16.reserve((size_t)P->n_pad * P->dp));
      CUDA_TRY(h->x_sumsq.reserve(P->n_pad));
      CUDA_TRY(h->x_resid2.reserve(P->n_pad));
      CUDA_TRY(h->x_max.reserve(2));
      CUDA_TRY(launch_to_bf16(h->vecs.p + h->stride, h->stride, h->n, (uint32_t)h->dim, h->x_bf16.p, P->dp, P->n_pad,
                              h->x_sumsq.p, h->x_resid2.p, s));
      CUDA_TRY(launch_tc_max(h->x_sumsq.p, h->x_resid2.p, h->n, h->x_max.p, s));
      CUDA_TRY(cudaStreamSynchronize(s));  // visible to every stream before tc_valid says so
      h->tc_valid = true;
      h->tc_n = h->n;
    }
  }
  CUDA_TRY(w.tc_beta.reserve(P->n_pad));
  CUDA_TRY(launch_tc_beta(h->dev(), h->x_sumsq.p, d_allow, P->use_norm, P->n_pad, w.tc_beta.p, s));
  return KDBGPU_OK;
}

// H2D of `c` raw queries, normalisation where the mode asks for it, bf16 copy + norms
int tc_stage_queries(kdbgpu_index *h, kdbgpu_index::FlatWs &w, const float *queries, uint32_t c, uint32_t c_pad, int mode,
                     const TcPlan &P, cudaStream_t s, cudaEvent_t after_h2d = nullptr) {
  CUDA_TRY(w.q_raw.reserve((size_t)c * h->dim));
  CUDA_TRY(w.q_prep.reserve((size_t)c * h->stride));
  CUDA_TRY(w.tq_bf16.reserve((size_t)c_pad * P.dp));
  CUDA_TRY(w.tq_sumsq.reserve(c_pad));
  CUDA_TRY(w.tq_resid2.reserve(c_pad));
  CUDA_TRY(cudaMemcpyAsync(w.q_raw.p, queries, (size_t)c * h->dim * sizeof(float), cudaMemcpyDefault, s));
  if (after_h2d) CUDA_TRY(cudaEventRecord(after_h2d, s));
  CUDA_TRY(launch_prep_queries(w.q_raw.p, (size_t)h->dim, w.q_prep.p, c, (uint32_t)h->dim, h->stride,
                               mode == 1 ? h->metric : KDBGPU_METRIC_L2, s));
  const bool prepared = mode == 1 && h->metric == KDBGPU_METRIC_COSINE;
  CUDA_TRY(launch_to_bf16(prepared ? w.q_prep.p : w.q_raw.p, prepared ? (size_t)h->stride : (size_t)h->dim, c,
                          (uint32_t)h->dim, w.tq_bf16.p, P.dp, c_pad, w.tq_sumsq.p, w.tq_resid2.p, s));
  return KDBGPU_OK;
}

FlatTcLaunch tc_launch_desc(kdbgpu_index *h, kdbgpu_index::FlatWs &w, const TcPlan &P, uint32_t c, uint32_t c_pad) {
  FlatTcLaunch L;
  memset(&L, 0, sizeof L);
  L.q_bf16 = w.tq_bf16.p;
  L.x_bf16 = h->x_bf16.p;
  L.nq = c;
  L.nq_pad = c_pad;
  L.n = h->n;
  L.n_pad = P.n_pad;
  L.dp = P.dp;
  L.alpha = P.alpha;
  L.beta = w.tc_beta.p;
  const uint64_t tiles = (uint64_t)(c_pad / P.bm) * (P.n_pad / P.bn);
  L.grid = tiles < (uint64_t)h->num_sms ? (int)tiles : h->num_sms;
  return L;
}

int flat_prefilter_impl(kdbgpu_index *h, kdbgpu_index::FlatWs &w, const float *queries, uint32_t nq, int k, int mode,
                        const uint32_t *d_allow, uint32_t *out_ids, double *out_scores, uint32_t *out_counts,
                        uint64_t *evals, uint64_t *fallbacks, float *gemm_ms, float *compute_ms) {
  KDB_NOTHROW_BEGIN
  cudaStream_t s = w.stream;
  TcPlan P;
  int rc = tc_prepare(h, w, mode, d_allow, &P, s);
  if (rc) return rc;
  const uint32_t cap = 2048u;                      // spill slots per query beyond the per-(query, CTA) slots
  const uint32_t sub_slots = flat_tc_sub_slots();
  const uint32_t want = 8192u;                     // nominees per query the sampling aims below
  const uint32_t fcap = k <= 256 ? 1024u : 4096u;  // survivors of the refinement, re-scored exactly
  // pass A may look at every ct_stride-th corpus tile only: any k rows bound the k-th best score from
  // above, a sample just makes theta looser (about ct_stride * k nominees instead of k)
  const uint32_t n_ctiles = P.n_pad / P.bn, gpt = flat_tc_groups_per_tile();
  uint32_t ct_stride = want / (6u * (uint32_t)k);
  // measured at 1 M x 768, k = 100 (profiles/r2_flat_rescore_ab.log): every 8th tile 704 k queries/s, every 12th 721 k,
  // every 16th 711 k (the looser threshold costs the nomination pass more than the threshold pass saves)
  if (ct_stride > 12) ct_stride = 12;
  while (ct_stride > 1 && (uint64_t)((n_ctiles + ct_stride - 1) / ct_stride) * gpt < 4ull * (uint64_t)k) --ct_stride;
  if (ct_stride < 1) ct_stride = 1;
  const char *env_stride = getenv("KDBGPU_FLAT_SAMPLE");
  if (env_stride && atoi(env_stride) >= 1 && atoi(env_stride) <= 64) ct_stride = (uint32_t)atoi(env_stride);
  const uint32_t n_groups = (n_ctiles + ct_stride - 1) / ct_stride * gpt;
  if (n_groups < (uint32_t)k || h->n <= cap || tc_rescore_smem((uint32_t)h->dim, fcap) > 200 * 1024) {
    // too few rows for a threshold to exist (or to be worth it): the exhaustive scan answers
    *evals = (uint64_t)nq * h->n;
    *fallbacks = nq;
    return flat_scan_impl(h, w, queries, nq, k, mode, d_allow, out_ids, out_scores, out_counts);
  }
  // chunk so that gmin[chunk][n_groups] stays within 512 MiB
  uint32_t chunk = (uint32_t)((512ull << 20) / ((size_t)n_groups * sizeof(float)));
  chunk = chunk / P.bm * P.bm;
  if (chunk < P.bm) chunk = P.bm;
  if (chunk > flat_tc_max_queries()) chunk = flat_tc_max_queries();
  const uint32_t nq_pad = (nq + P.bm - 1) / P.bm * P.bm;
  if (chunk > nq_pad) chunk = nq_pad;
  const uint32_t emit_grid = (uint32_t)h->num_sms;
  CUDA_TRY(w.t_gmin.reserve((size_t)chunk * n_groups));
  CUDA_TRY(w.t_theta.reserve(chunk));
  CUDA_TRY(w.t_bound.reserve(chunk));
  CUDA_TRY(w.t_thf.reserve(chunk));
  CUDA_TRY(w.t_cnt.reserve(chunk + 4));  // [chunk] spill counters + the two tile counters of the tensor passes
  CUDA_TRY(w.t_fcnt.reserve(chunk));
  CUDA_TRY(w.t_flags.reserve(chunk));
  CUDA_TRY(w.t_sub.reserve((size_t)chunk * emit_grid * sub_slots));
  CUDA_TRY(w.t_subcnt.reserve((size_t)chunk * emit_grid));
  CUDA_TRY(w.t_ovf.reserve((size_t)chunk * cap));
  CUDA_TRY(w.t_fid.reserve((size_t)chunk * fcap));
  CUDA_TRY(w.t_nres.reserve(1));
  CUDA_TRY(w.out_ids.reserve((size_t)chunk * k));
  CUDA_TRY(w.out_scores.reserve((size_t)chunk * k));
  CUDA_TRY(w.out_counts.reserve(chunk));
  CUDA_TRY(cudaMemsetAsync(w.t_nres.p, 0, sizeof(unsigned long long), s));
  DevIndex ix = h->dev();
  const bool prepared = mode == 1 && h->metric == KDBGPU_METRIC_COSINE;
  std::vector<uint32_t> flags(nq, 0u);
  *gemm_ms = 0.f;
  *compute_ms = 0.f;
  for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
    const uint32_t c = nq - q0 < chunk ? nq - q0 : chunk;
    const uint32_t c_pad = (c + P.bm - 1) / P.bm * P.bm;
    rc = tc_stage_queries(h, w, queries + (size_t)q0 * h->dim, c, c_pad, mode, P, s, w.ev[2]);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(w.t_cnt.p, 0, (size_t)(chunk + 4) * sizeof(uint32_t), s));
    FlatTcLaunch L = tc_launch_desc(h, w, P, c, c_pad);
    L.tile_counter = w.t_cnt.p + chunk;
    L.gmin = w.t_gmin.p;
    L.theta = w.t_theta.p;
    L.sub = w.t_sub.p;
    L.sub_cnt = w.t_subcnt.p;
    L.ovf_cnt = w.t_cnt.p;
    L.ovf = w.t_ovf.p;
    L.cap = cap;
    CUDA_TRY(cudaEventRecord(w.ev[3], s));
    L.epi = 1;  // pass A: group minima over the sampled tiles
    L.ct_stride = ct_stride;
    {
      const uint64_t tiles = (uint64_t)(c_pad / P.bm) * ((n_ctiles + ct_stride - 1) / ct_stride);
      L.grid = tiles < (uint64_t)h->num_sms ? (int)tiles : h->num_sms;
    }
    CUDA_TRY(launch_flat_tc(L, s));
    CUDA_TRY(launch_tc_threshold(w.t_gmin.p, n_groups, c, k, w.tq_sumsq.p, w.tq_resid2.p, h->x_max.p, P.alpha,
                                 P.use_norm, P.dp, w.t_theta.p, w.t_bound.p, s));
    L.epi = 2;  // pass B: ids + scores below theta, every tile
    L.tile_counter = w.t_cnt.p + chunk + 1;
    L.ct_stride = 1;
    L.grid = tc_launch_desc(h, w, P, c, c_pad).grid;
    CUDA_TRY(launch_flat_tc(L, s));
    const uint32_t grid_b = (uint32_t)L.grid;
    CUDA_TRY(cudaEventRecord(w.ev[4], s));
    CUDA_TRY(launch_tc_refine(w.t_sub.p, w.t_subcnt.p, grid_b, w.t_cnt.p, w.t_ovf.p, cap, c, k, w.t_theta.p,
                              w.t_bound.p, w.t_fcnt.p, w.t_fid.p, fcap, w.t_thf.p, w.t_flags.p, s));
    CUDA_TRY(launch_tc_rescore(ix, mode, prepared ? w.q_prep.p : w.q_raw.p, prepared ? (size_t)h->stride : (size_t)h->dim,
                               c, k, w.t_fcnt.p, w.t_fid.p, fcap, w.t_thf.p, w.t_bound.p, w.tq_sumsq.p,
                               w.out_ids.p, w.out_scores.p, w.out_counts.p, w.t_flags.p, w.t_nres.p, s));
    CUDA_TRY(cudaEventRecord(w.ev[5], s));
    rc = flat_deliver(h, w, c, k, out_ids + (size_t)q0 * k, out_scores + (size_t)q0 * k, out_counts + q0, flags.data() + q0, s);
    if (rc) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, w.ev[3], w.ev[4]);
    *gemm_ms += ms;
    cudaEventElapsedTime(&ms, w.ev[2], w.ev[5]);
    *compute_ms += ms;
  }
  unsigned long long nres = 0;
  CUDA_TRY(cudaMemcpyAsync(&nres, w.t_nres.p, sizeof nres, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  // queries whose certificate did not close (candidate buffer overflow) take the exhaustive scan
  std::vector<uint32_t> redo;
  for (uint32_t q = 0; q < nq; ++q)
    if (flags[q]) redo.push_back(q);
  if (!redo.empty()) {
    const size_t r = redo.size();
    std::vector<float> rq(r * (size_t)h->dim);
    std::vector<uint32_t> rid(r * (size_t)k), rcnt(r);
    std::vector<double> rsc(r * (size_t)k);
    // `queries` / the outputs may live on either side (cudaMemcpyDefault); this path is rare
    for (size_t i = 0; i < r; ++i)
      CUDA_TRY(cudaMemcpy(&rq[i * h->dim], queries + (size_t)redo[i] * h->dim, (size_t)h->dim * sizeof(float),
                          cudaMemcpyDefault));
    rc = flat_scan_impl(h, w, rq.data(), (uint32_t)r, k, mode, d_allow, rid.data(), rsc.data(), rcnt.data());
    if (rc) return rc;
    for (size_t i = 0; i < r; ++i) {
      CUDA_TRY(cudaMemcpy(out_ids + (size_t)redo[i] * k, &rid[i * k], (size_t)k * sizeof(uint32_t), cudaMemcpyDefault));
      CUDA_TRY(cudaMemcpy(out_scores + (size_t)redo[i] * k, &rsc[i * k], (size_t)k * sizeof(double), cudaMemcpyDefault));
      CUDA_TRY(cudaMemcpy(out_counts + redo[i], &rcnt[i], sizeof(uint32_t), cudaMemcpyDefault));
    }
  }
  *evals = nres + (uint64_t)redo.size() * h->n;
  *fallbacks = redo.size();
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

}  // namespace

namespace kdb {
int acquire_fws(kdbgpu_index *h) {
  std::unique_lock<std::mutex> lk(h->fws_mu);
  for (;;) {
    for (int i = 0; i < kdbgpu_index::kNumFlatWs; ++i)
      if (!h->fws[i].busy) {
        h->fws[i].busy = true;
        return i;
      }
    h->fws_cv.wait(lk);
  }
}
void release_fws(kdbgpu_index *h, int i) {
  {
    std::lock_guard<std::mutex> lk(h->fws_mu);
    h->fws[i].busy = false;
  }
  h->fws_cv.notify_one();
}

int flat_search_ws(kdbgpu_index *h, kdbgpu_index::FlatWs &w, const float *queries, uint32_t nq, int k, int mode,
                   bool prefilter, const uint32_t *d_allow, uint32_t *out_ids, double *out_scores, uint32_t *out_counts,
                   kdbgpu_stats *stats) {
  cudaStream_t s = w.stream;
  CUDA_TRY(cudaEventRecord(w.ev[0], s));
  uint64_t evals = 0, fallbacks = 0;
  float gemm_ms = 0.f, compute_ms = 0.f;
  int rc;
  if (prefilter)
    rc = flat_prefilter_impl(h, w, queries, nq, k, mode, d_allow, out_ids, out_scores, out_counts, &evals, &fallbacks,
                             &gemm_ms, &compute_ms);
  else {
    rc = flat_scan_impl(h, w, queries, nq, k, mode, d_allow, out_ids, out_scores, out_counts);
    evals = (uint64_t)nq * h->n;
  }
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(w.ev[1], s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (stats) {
    stats->dist_evals = evals;
    stats->hops = fallbacks;
    cudaEventElapsedTime(&stats->total_ms, w.ev[0], w.ev[1]);
    // pre-filter: kernel_ms = every kernel of the call (no copies); hops_l0 = the two tensor-core
    // passes + threshold alone, in nanoseconds
    stats->kernel_ms = prefilter && compute_ms > 0.f ? compute_ms : stats->total_ms;
    stats->hops_l0 = prefilter ? (uint64_t)(gemm_ms * 1e6f) : 0;
  }
  return KDBGPU_OK;
}
}  // namespace kdb

extern "C" {

int kdbgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

const char *kdbgpu_last_error(void) { return g_last_error.c_str(); }

const char *kdbgpu_version(void) { return "kektordb_gpu 0.1.0 sm_100a"; }

int kdbgpu_index_create(int device, int dim, int metric, int m, uint32_t capacity, kdbgpu_index **out) {
  return kdbgpu_index_create_ex(device, dim, metric, KDBGPU_PRECISION_F32, m, capacity, out);
}

int kdbgpu_index_create_ex(int device, int dim, int metric, int precision, int m, uint32_t capacity,
                           kdbgpu_index **out) {
  if (!out) return fail(KDBGPU_ERR_INVALID, "out is NULL");
  *out = nullptr;
  // GetFloat16Func / GetInt8Func (distance_go.go:159-177): float16 is Euclidean only, int8 Cosine only
  if (precision != KDBGPU_PRECISION_F32 && precision != KDBGPU_PRECISION_F16 && precision != KDBGPU_PRECISION_INT8)
    return fail(KDBGPU_ERR_INVALID, "unknown precision %d", precision);
  if (precision == KDBGPU_PRECISION_F16 && metric != KDBGPU_METRIC_L2)
    return fail(KDBGPU_ERR_INVALID, "metric 'cosine' not supported for float16 precision");
  if (precision == KDBGPU_PRECISION_INT8 && metric != KDBGPU_METRIC_COSINE)
    return fail(KDBGPU_ERR_INVALID, "metric 'euclidean' not supported for int8 precision");
  if (dim <= 0 || dim > 8192) return fail(KDBGPU_ERR_INVALID, "dim %d out of range (1..8192)", dim);
  if (metric != KDBGPU_METRIC_L2 && metric != KDBGPU_METRIC_COSINE)
    return fail(KDBGPU_ERR_INVALID, "unknown metric %d", metric);
  if (m <= 0) m = 16;  // hnsw.New default (hnsw_index.go:140-142)
  if (m > 128) return fail(KDBGPU_ERR_INVALID, "m %d too large (max 128)", m);
  if (capacity == 0 || capacity > 0xfffffff0u) return fail(KDBGPU_ERR_INVALID, "capacity out of range");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(KDBGPU_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
  }
  if (device < 0 || device >= ndev) return fail(KDBGPU_ERR_INVALID, "device %d of %d", device, ndev);
  DeviceGuard g(device);
  if (!g.ok) return fail(KDBGPU_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(KDBGPU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  kdbgpu_index *h = new (std::nothrow) kdbgpu_index();
  if (!h) return fail(KDBGPU_ERR_NOMEM, "host allocation failed");
  h->device = device;
  h->dim = dim;
  h->metric = metric;
  h->m = m;
  h->capacity = capacity;
  h->precision = precision;
  if (precision == KDBGPU_PRECISION_F32) {
    h->kind = metric == KDBGPU_METRIC_COSINE ? KIND_COS_F32 : KIND_L2_F32;
    h->stride = ((uint32_t)dim + 127u) & ~127u;  // whole float4 columns for every lane (searcher.cuh)
    h->row_words = h->stride;
  } else {
    const uint32_t esize = precision == KDBGPU_PRECISION_F16 ? 2u : 1u;
    h->kind = precision == KDBGPU_PRECISION_F16 ? KIND_L2_F16 : KIND_COS_I8;
    h->row_words = (((uint32_t)dim * esize + 127u) & ~127u) / 4u;  // row pitch: whole 128-byte lines
    h->stride = (h->row_words + 127u) & ~127u;                     // shared-memory slot: whole 16-byte columns per lane
  }
  h->num_sms = prop.multiProcessorCount;
  const char *env;
  // float16 / int8 rows: the traversal is latency-bound, so resident query-warps matter more than
  // shared-memory heap capacity (measured: profiles/README.md, quantized sweep)
  if (precision != KDBGPU_PRECISION_F32) h->tuning.cand_smem = precision == KDBGPU_PRECISION_F16 ? 128 : 64;
  // launch shapes measured at 1 M x 768 (profiles/README.md): rows per group = slots / 2.  float32 / float16 rows are
  // HBM-bound with batches in flight (more resident query-warps win: 4 slots) and latency-bound alone (8 slots);
  // int8 rows are issue-bound either way (8 slots halve the per-row overhead, 16 when alone)
  h->tuning.slots = precision == KDBGPU_PRECISION_INT8 ? 8 : 4;
  h->tuning.slots_idle = precision == KDBGPU_PRECISION_INT8 ? 16 : 8;
  // long rows: residency first.  Four 6 KB slots (1536-d float32) leave 7 query-warps per SM; with two there are 13,
  // and the hybrid workload (BASELINE configs[4], ~7 rows per hop) answers 765 k instead of 643 k queries/s
  if (4u * h->row_words * (uint32_t)sizeof(float) > 16384u) {
    h->tuning.slots = 2;
    h->tuning.slots_idle = 4;
  }
  if ((env = getenv("KDBGPU_FAST"))) h->tuning.fast = atoi(env);
  if ((env = getenv("KDBGPU_SLOTS"))) h->tuning.slots = atoi(env);
  if ((env = getenv("KDBGPU_SLOTS_IDLE"))) h->tuning.slots_idle = atoi(env);
  if ((env = getenv("KDBGPU_ROWS_EVICT_FIRST"))) h->tuning.rows_evict_first = atoi(env);
  if ((env = getenv("KDBGPU_CAND_SMEM"))) h->tuning.cand_smem = atoi(env);
  if ((env = getenv("KDBGPU_MAX_CTAS_PER_SM"))) h->tuning.max_ctas_per_sm = atoi(env);
  auto cleanup = [&](int rc) {
    kdbgpu_index_destroy(h);
    return rc;
  };
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
  for (auto &w : h->sws) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&w.ev[i]);
  }
  for (auto &w : h->fws) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&w.ev[i]);
  }
  const size_t n1 = (size_t)capacity + 1;
  if (e == cudaSuccess) e = h->vecs.reserve(n1 * h->row_words + 128, true);
  if (e == cudaSuccess && precision == KDBGPU_PRECISION_INT8) e = h->norms.reserve(n1, true);
  if (e == cudaSuccess) e = h->adj0.reserve(n1 * (size_t)(2 * m), true);
  if (e == cudaSuccess) e = h->levels.reserve(n1);
  if (e == cudaSuccess) e = cudaMemset(h->levels.p, 0xff, n1);
  if (e == cudaSuccess) e = h->upper_first.reserve(n1, true);
  if (e == cudaSuccess) e = h->deleted.reserve((n1 + 31) / 32 + 2, true);
  const size_t up_rows = (size_t)((double)capacity / (double)(m > 1 ? m - 1 : 1) * 1.3) + 4096;
  if (e == cudaSuccess) e = h->upper_adj.reserve(up_rows * (size_t)m, true);
  if (e == cudaSuccess) e = h->upper_node.reserve(up_rows, true);
  if (e == cudaSuccess) e = h->upper_level.reserve(up_rows, true);
  h->h_levels.assign(n1, (int8_t)-1);
  h->h_upper_first.assign(n1, 0u);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return cleanup(fail(e == cudaErrorMemoryAllocation ? KDBGPU_ERR_NOMEM : KDBGPU_ERR_CUDA,
                        "index allocation failed: %s", cudaGetErrorString(e)));
  }
  *out = h;
  return KDBGPU_OK;
}

int kdbgpu_index_destroy(kdbgpu_index *h) {
  if (!h) return KDBGPU_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  for (auto &w : h->sws) {
    w.release();
    if (w.done) cudaEventDestroy(w.done);
    for (auto &e : w.ev)
      if (e) cudaEventDestroy(e);
    if (w.stream) cudaStreamDestroy(w.stream);
  }
  h->vecs.release();
  h->norms.release();
  h->conv_tmp.release();
  h->qnorm1.release();
  h->adj0.release();
  h->upper_adj.release();
  h->upper_first.release();
  h->deleted.release();
  h->levels.release();
  h->q_raw.release();
  h->q_prep.release();
  h->out_ids.release();
  h->out_counts.release();
  h->allow.release();
  h->visited.release();
  h->ids_tmp.release();
  h->out_scores.release();
  h->flat_dist.release();
  h->dist_tmp.release();
  h->cand_overflow.release();
  h->stats.release();
  h->work_counter.release();
  h->err_flag.release();
  h->upper_node.release();
  h->upper_level.release();
  h->b_out_off.release();
  h->b_slot_node.release();
  h->b_cand_ids.release();
  h->b_cand_cnt.release();
  h->b_row_cnt.release();
  h->b_row_off.release();
  h->b_srcs.release();
  h->b_active.release();
  h->b_scalars.release();
  h->b_scratch_ids.release();
  h->b_slot_level.release();
  h->b_scratch_d.release();
  h->x_bf16.release(); h->x_sumsq.release(); h->x_resid2.release(); h->x_max.release();
  for (auto &w : h->fws) {
    w.release();
    for (auto &e : w.ev)
      if (e) cudaEventDestroy(e);
    if (w.stream) cudaStreamDestroy(w.stream);
  }
  for (auto &e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  (void)cudaGetLastError();
  delete h;
  return KDBGPU_OK;
}

int kdbgpu_upload_vectors(kdbgpu_index *h, uint32_t first_id, uint32_t count, const float *rows) {
  if (!h || (!rows && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (first_id == 0 || (uint64_t)first_id + count - 1 > h->capacity)
    return fail(KDBGPU_ERR_INVALID, "ids %u..%llu outside 1..%u", first_id, (unsigned long long)first_id + count - 1,
                h->capacity);
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());  // drain searches queued through the asynchronous entry point
  h->tc_valid = false;
  if (h->precision != KDBGPU_PRECISION_F32) {
    // float32 in, stored form out: what Add / AddBatch do before the row reaches the arena
    // (hnsw_index.go:497-520, :1553-1577); chunked through a device staging buffer
    if (h->precision == KDBGPU_PRECISION_INT8 && h->abs_max == 0.f)
      return fail(KDBGPU_ERR_STATE, "int8 index without a trained quantizer (kdbgpu_set_quantizer)");
    const uint32_t chunk = 1u << 16;
    CUDA_TRY(h->conv_tmp.reserve((size_t)(count < chunk ? count : chunk) * h->dim));
    for (uint32_t i = 0; i < count; i += chunk) {
      const uint32_t c = count - i < chunk ? count - i : chunk;
      CUDA_TRY(cudaMemcpyAsync(h->conv_tmp.p, rows + (size_t)i * h->dim, (size_t)c * h->dim * sizeof(float),
                               cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(launch_convert_rows(h->conv_tmp.p, (size_t)h->dim, h->vecs.p + (size_t)(first_id + i) * h->row_words,
                                   h->row_words, c, (uint32_t)h->dim, h->kind, false, h->abs_max,
                                   h->norms.p ? h->norms.p + first_id + i : nullptr, false, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return KDBGPU_OK;
  }
  // padded columns were zeroed at creation and are never written
  CUDA_TRY(cudaMemcpy2DAsync(h->vecs.p + (size_t)first_id * h->stride, (size_t)h->stride * sizeof(float), rows,
                             (size_t)h->dim * sizeof(float), (size_t)h->dim * sizeof(float), count,
                             cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return KDBGPU_OK;
}

int kdbgpu_upload_rows_raw(kdbgpu_index *h, uint32_t first_id, uint32_t count, const void *rows) {
  if (!h || (!rows && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (first_id == 0 || (uint64_t)first_id + count - 1 > h->capacity)
    return fail(KDBGPU_ERR_INVALID, "ids %u..%llu outside 1..%u", first_id, (unsigned long long)first_id + count - 1,
                h->capacity);
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  h->tc_valid = false;
  const size_t esize = h->precision == KDBGPU_PRECISION_F32 ? 4 : (h->precision == KDBGPU_PRECISION_F16 ? 2 : 1);
  const size_t row_bytes = (size_t)h->dim * esize;
  CUDA_TRY(cudaMemcpy2DAsync(h->vecs.p + (size_t)first_id * h->row_words, (size_t)h->row_words * sizeof(float), rows,
                             row_bytes, row_bytes, count, cudaMemcpyHostToDevice, h->stream));
  if (h->precision == KDBGPU_PRECISION_INT8)
    CUDA_TRY(launch_int8_norms(h->vecs.p + (size_t)first_id * h->row_words, h->row_words, count, (uint32_t)h->dim,
                               h->norms.p + first_id, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return KDBGPU_OK;
}

int kdbgpu_download_rows_raw(kdbgpu_index *h, uint32_t first_id, uint32_t count, void *rows) {
  if (!h || (!rows && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (first_id == 0 || (uint64_t)first_id + count - 1 > h->capacity)
    return fail(KDBGPU_ERR_INVALID, "ids outside 1..%u", h->capacity);
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t esize = h->precision == KDBGPU_PRECISION_F32 ? 4 : (h->precision == KDBGPU_PRECISION_F16 ? 2 : 1);
  const size_t row_bytes = (size_t)h->dim * esize;
  CUDA_TRY(cudaMemcpy2D(rows, row_bytes, h->vecs.p + (size_t)first_id * h->row_words,
                        (size_t)h->row_words * sizeof(float), row_bytes, count, cudaMemcpyDeviceToHost));
  return KDBGPU_OK;
}

int kdbgpu_download_norms(kdbgpu_index *h, uint32_t first_id, uint32_t count, float *norms) {
  if (!h || (!norms && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (h->precision != KDBGPU_PRECISION_INT8) return fail(KDBGPU_ERR_INVALID, "norms exist for int8 indexes only");
  if (first_id == 0 || (uint64_t)first_id + count - 1 > h->capacity)
    return fail(KDBGPU_ERR_INVALID, "ids outside 1..%u", h->capacity);
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(norms, h->norms.p + first_id, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
  return KDBGPU_OK;
}

int kdbgpu_set_quantizer(kdbgpu_index *h, float abs_max) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (h->precision != KDBGPU_PRECISION_INT8) return fail(KDBGPU_ERR_INVALID, "the quantizer belongs to int8 indexes");
  if (!(abs_max >= 0.f)) return fail(KDBGPU_ERR_INVALID, "AbsMax must be >= 0");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  h->abs_max = abs_max;
  return KDBGPU_OK;
}

namespace {
// The sample of Quantizer.Train (pkg/core/distance/quantizer.go:62-94): everything up to 10 000 rows,
// above that every step-th row (10 % of the rows, capped at 25 000, floor 10 000).
void train_sample(uint32_t n, uint32_t *step, uint32_t *count) {
  *step = 1;
  *count = n;
  if (n > 10000u) {
    uint32_t target = n / 10u;
    if (target > 25000u) target = 25000u;
    if (target < 10000u) target = 10000u;
    uint32_t st = n / target;
    if (st < 1) st = 1;
    uint32_t c = 0;
    for (uint64_t i = 0; i < n; i += st) {
      ++c;
      if (c >= target) break;
    }
    *step = st;
    *count = c;
  }
}
// AbsMax = the value at index int(len * 0.999) of the ascending |values| of the sample (:96-125) —
// an exact 3-pass radix select on the device instead of the reference's sort.  Sampled row i is
// d_rows[i * step * row_stride ..].
int train_select(kdbgpu_index *h, const float *d_rows, size_t row_stride, uint32_t count, uint32_t step,
                 float *abs_max) {
  KDB_NOTHROW_BEGIN
  const uint64_t total = (uint64_t)count * (uint64_t)h->dim;
  long long qi = (long long)((double)total * 0.999);
  if (qi >= (long long)total) qi = (long long)total - 1;
  if (qi < 0) qi = 0;
  DevBuf<unsigned long long> hist;
  CUDA_TRY(hist.reserve(2048));
  std::vector<unsigned long long> hh(2048);
  uint64_t rem = (uint64_t)qi + 1;  // 1-based rank of the wanted value
  uint32_t prefix = 0;
  int shift = 32, rc = KDBGPU_OK;
  cudaStream_t s = h->stream;
  for (int pass = 0; pass < 3 && rc == KDBGPU_OK; ++pass) {
    const int bits = pass < 2 ? 11 : 10;
    const int hi_shift = shift;
    shift -= bits;
    cudaError_t e = cudaMemsetAsync(hist.p, 0, 2048 * sizeof(unsigned long long), s);
    if (e == cudaSuccess)
      e = launch_abs_hist(d_rows, row_stride, count, step, (uint32_t)h->dim, prefix, hi_shift, shift, bits, hist.p, s);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(hh.data(), hist.p, 2048 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      rc = fail(KDBGPU_ERR_CUDA, "quantizer training: %s", cudaGetErrorString(e));
      break;
    }
    uint32_t b = 0;
    for (; b < (1u << bits); ++b) {
      if (hh[b] >= rem) break;
      rem -= hh[b];
    }
    if (b == (1u << bits)) {
      rc = fail(KDBGPU_ERR_INVALID, "quantizer training saw non-finite values");
      break;
    }
    prefix |= b << shift;
  }
  hist.release();
  if (rc != KDBGPU_OK) return rc;
  float v;
  memcpy(&v, &prefix, sizeof v);
  h->abs_max = v;
  if (abs_max) *abs_max = v;
  return KDBGPU_OK;
  KDB_NOTHROW_END
}
}  // namespace

int kdbgpu_train_quantizer(kdbgpu_index *h, const float *rows, uint32_t n, float *abs_max) {
  if (!h || (!rows && n)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (h->precision != KDBGPU_PRECISION_INT8) return fail(KDBGPU_ERR_INVALID, "the quantizer belongs to int8 indexes");
  if (n == 0) return KDBGPU_OK;  // "Empty dataset provided for training. Skipping." (:52-55)
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  uint32_t step, count;
  train_sample(n, &step, &count);
  // only the sampled rows travel; they land densely, so the select runs over them with step 1
  CUDA_TRY(h->conv_tmp.reserve((size_t)count * h->dim));
  CUDA_TRY(cudaMemcpy2DAsync(h->conv_tmp.p, (size_t)h->dim * sizeof(float), rows, (size_t)step * h->dim * sizeof(float),
                             (size_t)h->dim * sizeof(float), count, cudaMemcpyHostToDevice, h->stream));
  return train_select(h, h->conv_tmp.p, (size_t)h->dim, count, 1, abs_max);
}

int kdbgpu_train_quantizer_device(kdbgpu_index *h, const float *d_rows, size_t row_stride, uint32_t n, float *abs_max) {
  if (!h || (!d_rows && n)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (h->precision != KDBGPU_PRECISION_INT8) return fail(KDBGPU_ERR_INVALID, "the quantizer belongs to int8 indexes");
  if (row_stride < (size_t)h->dim) return fail(KDBGPU_ERR_INVALID, "bad stride");
  if (n == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  uint32_t step, count;
  train_sample(n, &step, &count);
  return train_select(h, d_rows, row_stride, count, step, abs_max);
}

namespace {
uint32_t elem_bytes(const kdbgpu_index *h) {
  return h->precision == KDBGPU_PRECISION_F32 ? 4u : (h->precision == KDBGPU_PRECISION_F16 ? 2u : 1u);
}
// which chunks hold a slot of ids 1..last_id (host scan of the slot table; NULL = sequential slots)
std::vector<char> chunks_in_use(const uint32_t *slot_table, uint32_t last_id, uint32_t vpc) {
  std::vector<char> used;
  auto mark = [&](uint32_t p) {
    const uint32_t c = p / vpc;
    if (c >= used.size()) used.resize((size_t)c + 1, 0);
    used[c] = 1;
  };
  if (!slot_table) {
    if (last_id >= 1) {
      mark(0);
      mark(last_id - 1);
      for (size_t c = 0; c < used.size(); ++c) used[c] = 1;
    }
  } else {
    for (uint32_t id = 1; id <= last_id; ++id)
      if (slot_table[id] != 0xffffffffu) mark(slot_table[id]);
  }
  return used;
}
}  // namespace

// Host-only: what a directory of arena chunks holds (loadExistingChunks + the header fields addChunk
// validates, arena.go:283-305, :346-364).  No device needed.
int kdbgpu_arena_probe(const char *dir, uint32_t *dim, int *precision, uint32_t *n_chunks, uint32_t *vecs_per_chunk) {
  if (!dir) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  std::string err;
  const int max_chunk = arena_max_chunk(dir, &err);
  if (max_chunk < 0) return fail(KDBGPU_ERR_INVALID, "no arena_%%04d.bin chunk in %s", dir);
  unsigned char hdr[64], dummy[16];
  uint32_t d0 = 0;
  int p0 = -1;
  for (int c = 0; c <= max_chunk; ++c) {
    if (arena_read_chunk(dir, c, hdr, dummy, 0, &err) < 0) continue;  // a dropped chunk (DeferDropChunk) leaves a hole
    const uint32_t magic = (uint32_t)hdr[0] | ((uint32_t)hdr[1] << 8) | ((uint32_t)hdr[2] << 16) | ((uint32_t)hdr[3] << 24);
    const uint32_t version = (uint32_t)hdr[4] | ((uint32_t)hdr[5] << 8) | ((uint32_t)hdr[6] << 16) | ((uint32_t)hdr[7] << 24);
    const uint32_t fdim = (uint32_t)hdr[8] | ((uint32_t)hdr[9] << 8) | ((uint32_t)hdr[10] << 16) | ((uint32_t)hdr[11] << 24);
    if (magic != 0x4B414F4Eu) return fail(KDBGPU_ERR_INVALID, "file arena_%04d.bin is not a valid arena (magic mismatch)", c);
    if (version != 1u) return fail(KDBGPU_ERR_INVALID, "file arena_%04d.bin unsupported version %u", c, version);
    if (p0 < 0) {
      d0 = fdim;
      p0 = (int)hdr[12];
    } else if (fdim != d0 || (int)hdr[12] != p0) {
      return fail(KDBGPU_ERR_INVALID, "file arena_%04d.bin disagrees with chunk 0 on dimension / precision", c);
    }
  }
  if (p0 < 0 || p0 > 2 || d0 == 0) return fail(KDBGPU_ERR_INVALID, "no readable chunk header in %s", dir);
  const uint32_t vb = d0 * (p0 == 0 ? 4u : (p0 == 1 ? 2u : 1u));
  if (dim) *dim = d0;
  if (precision) *precision = p0;
  if (n_chunks) *n_chunks = (uint32_t)max_chunk + 1;
  if (vecs_per_chunk) *vecs_per_chunk = arena_vecs_per_chunk(vb);
  return KDBGPU_OK;
}

int kdbgpu_arena_stage_chunk(kdbgpu_index *h, uint32_t chunk_id, const void *chunk, size_t chunk_bytes,
                             const uint32_t *slot_table, uint32_t table_len, uint32_t *rows_staged) {
  if (rows_staged) *rows_staged = 0;
  if (!h || !chunk) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (chunk_bytes < arena_header_size()) return fail(KDBGPU_ERR_INVALID, "chunk shorter than its header");
  if (table_len < 2) return KDBGPU_OK;
  const unsigned char *bytes = static_cast<const unsigned char *>(chunk);
  int rc = arena_check_header(bytes, (uint32_t)h->dim, h->precision, "chunk");
  if (rc) return rc;
  const uint32_t vb = (uint32_t)h->dim * elem_bytes(h), vpc = arena_vecs_per_chunk(vb);
  if (vpc == 0) return fail(KDBGPU_ERR_INVALID, "vector size %u exceeds chunk payload capacity", vb);
  const size_t payload = (size_t)vpc * vb;
  size_t have = chunk_bytes - arena_header_size();  // a short buffer: the rest reads as zero pages, as in the mmap
  if (have > payload) have = payload;
  uint32_t last_id = table_len - 1;
  if (last_id > h->capacity) last_id = h->capacity;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  h->tc_valid = false;
  DevBuf<unsigned char> d_stage;
  DevBuf<uint32_t> d_slot;
  DevBuf<unsigned int> d_cnt;
  struct Free {
    DevBuf<unsigned char> &a;
    DevBuf<uint32_t> &b;
    DevBuf<unsigned int> &c;
    ~Free() { a.release(); b.release(); c.release(); }
  } freer{d_stage, d_slot, d_cnt};
  CUDA_TRY(d_stage.reserve(payload + 16));
  CUDA_TRY(d_cnt.reserve(1, true));
  cudaStream_t s = h->stream;
  if (slot_table) {
    CUDA_TRY(d_slot.reserve(table_len));
    CUDA_TRY(cudaMemcpyAsync(d_slot.p, slot_table, (size_t)table_len * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  }
  if (have < payload) CUDA_TRY(cudaMemsetAsync(d_stage.p + have, 0, payload - have, s));
  // The chunk is the host's own mapping of arena_%04d.bin (VectorArena keeps every chunk mmap'ed, arena.go:307-376).
  // Page-lock it in place for the duration of the copy, so the DMA engine reads the mapping directly instead of
  // the driver bouncing it through its own pinned staging; a mapping that cannot be registered (not page
  // aligned, read-only registration unsupported) is copied the ordinary way.
  const uintptr_t page = 4096;
  const uintptr_t lo = reinterpret_cast<uintptr_t>(bytes);
  const size_t span = (arena_header_size() + have + page - 1) & ~(size_t)(page - 1);  // whole pages of the mapping
  bool registered = false;
  if (have >= (1u << 20) && (lo & (page - 1)) == 0 && getenv("KDBGPU_ARENA_NO_REGISTER") == nullptr) {
    cudaError_t re = cudaHostRegister(reinterpret_cast<void *>(lo), span, cudaHostRegisterReadOnly);
    if (re != cudaSuccess) {
      (void)cudaGetLastError();
      re = cudaHostRegister(reinterpret_cast<void *>(lo), span, cudaHostRegisterDefault);  // a writable mapping
    }
    if (re == cudaSuccess)
      registered = true;
    else
      (void)cudaGetLastError();
  }
  struct Unregister {
    void *p;
    bool on;
    ~Unregister() {
      if (on) {
        cudaHostUnregister(p);
        (void)cudaGetLastError();
      }
    }
  } unreg{reinterpret_cast<void *>(lo), registered};
  h->arena_chunks_registered += registered ? 1 : 0;
  CUDA_TRY(cudaMemcpyAsync(d_stage.p, bytes + arena_header_size(), have, cudaMemcpyHostToDevice, s));
  CUDA_TRY(launch_arena_scatter(d_stage.p, chunk_id, vpc, vb, slot_table ? d_slot.p : nullptr, 1, last_id, h->vecs.p,
                                h->row_words, d_cnt.p, s));
  if (h->precision == KDBGPU_PRECISION_INT8)
    CUDA_TRY(launch_int8_norms(h->vecs.p + (size_t)h->row_words, h->row_words, last_id, (uint32_t)h->dim, h->norms.p + 1, s));
  unsigned int cnt = 0;
  CUDA_TRY(cudaMemcpyAsync(&cnt, d_cnt.p, sizeof cnt, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (rows_staged) *rows_staged = cnt;
  return KDBGPU_OK;
}

int kdbgpu_arena_load_dir(kdbgpu_index *h, const char *dir, const uint32_t *slot_table, uint32_t table_len,
                          uint64_t *rows_staged) {
  KDB_NOTHROW_BEGIN
  if (rows_staged) *rows_staged = 0;
  if (!h || !dir) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (table_len < 2) return KDBGPU_OK;
  const uint32_t vb = (uint32_t)h->dim * elem_bytes(h), vpc = arena_vecs_per_chunk(vb);
  if (vpc == 0) return fail(KDBGPU_ERR_INVALID, "vector size %u exceeds chunk payload capacity", vb);
  const size_t payload = (size_t)vpc * vb;
  uint32_t last_id = table_len - 1;
  if (last_id > h->capacity) last_id = h->capacity;
  std::string err;
  const int max_chunk = arena_max_chunk(dir, &err);
  if (max_chunk < 0) return fail(KDBGPU_ERR_INVALID, "no arena_%%04d.bin chunk in %s", dir);
  const std::vector<char> used = chunks_in_use(slot_table, last_id, vpc);
  for (size_t c = 0; c < used.size(); ++c)
    if (used[c] && (int)c > max_chunk)
      return fail(KDBGPU_ERR_INVALID, "slot table refers to chunk %zu but %s holds chunks 0..%d", c, dir, max_chunk);
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  h->tc_valid = false;
  DevBuf<unsigned char> d_stage[2];
  DevBuf<uint32_t> d_slot;
  DevBuf<unsigned int> d_cnt;
  unsigned char *pinned[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  struct Free {
    DevBuf<unsigned char> *st;
    DevBuf<uint32_t> &b;
    DevBuf<unsigned int> &c;
    unsigned char **pin;
    cudaEvent_t *ev;
    ~Free() {
      st[0].release(); st[1].release(); b.release(); c.release();
      for (int i = 0; i < 2; ++i) {
        if (pin[i]) cudaFreeHost(pin[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
      }
    }
  } freer{d_stage, d_slot, d_cnt, pinned, ev};
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(d_stage[i].reserve(payload + 16));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&pinned[i]), payload, cudaHostAllocDefault));
    CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  CUDA_TRY(d_cnt.reserve(1, true));
  cudaStream_t s = h->stream;
  if (slot_table) {
    CUDA_TRY(d_slot.reserve(table_len));
    CUDA_TRY(cudaMemcpyAsync(d_slot.p, slot_table, (size_t)table_len * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  }
  int nb = 0;
  for (size_t c = 0; c < used.size(); ++c) {
    if (!used[c]) continue;
    const int b = nb++ & 1;
    CUDA_TRY(cudaEventSynchronize(ev[b]));  // the copy that last read this pinned buffer is done
    unsigned char hdr[64];
    if (arena_read_chunk(dir, (int)c, hdr, pinned[b], payload, &err) < 0) return fail(KDBGPU_ERR_INVALID, "%s", err.c_str());
    char what[64];
    snprintf(what, sizeof what, "file arena_%04d.bin", (int)c);
    int rc = arena_check_header(hdr, (uint32_t)h->dim, h->precision, what);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_stage[b].p, pinned[b], payload, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(ev[b], s));
    CUDA_TRY(launch_arena_scatter(d_stage[b].p, (uint32_t)c, vpc, vb, slot_table ? d_slot.p : nullptr, 1, last_id,
                                  h->vecs.p, h->row_words, d_cnt.p, s));
  }
  if (h->precision == KDBGPU_PRECISION_INT8)
    CUDA_TRY(launch_int8_norms(h->vecs.p + (size_t)h->row_words, h->row_words, last_id, (uint32_t)h->dim, h->norms.p + 1, s));
  unsigned int cnt = 0;
  CUDA_TRY(cudaMemcpyAsync(&cnt, d_cnt.p, sizeof cnt, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (rows_staged) *rows_staged = cnt;
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

int kdbgpu_index_precision(const kdbgpu_index *h) { return h ? h->precision : -1; }
uint64_t kdbgpu_arena_chunks_registered(const kdbgpu_index *h) { return h ? h->arena_chunks_registered : 0; }

int kdbgpu_upload_vectors_device(kdbgpu_index *h, uint32_t first_id, uint32_t count, const float *d_rows,
                                 size_t row_stride) {
  if (!h || (!d_rows && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (first_id == 0 || (uint64_t)first_id + count - 1 > h->capacity || row_stride < (size_t)h->dim)
    return fail(KDBGPU_ERR_INVALID, "ids %u..+%u outside 1..%u or bad stride", first_id, count, h->capacity);
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());  // drain searches queued through the asynchronous entry point
  h->tc_valid = false;
  if (h->precision != KDBGPU_PRECISION_F32) {
    if (h->precision == KDBGPU_PRECISION_INT8 && h->abs_max == 0.f)
      return fail(KDBGPU_ERR_STATE, "int8 index without a trained quantizer (kdbgpu_set_quantizer)");
    CUDA_TRY(launch_convert_rows(d_rows, row_stride, h->vecs.p + (size_t)first_id * h->row_words, h->row_words, count,
                                 (uint32_t)h->dim, h->kind, false, h->abs_max,
                                 h->norms.p ? h->norms.p + first_id : nullptr, false, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return KDBGPU_OK;
  }
  CUDA_TRY(cudaMemcpy2DAsync(h->vecs.p + (size_t)first_id * h->stride, (size_t)h->stride * sizeof(float), d_rows,
                             row_stride * sizeof(float), (size_t)h->dim * sizeof(float), count,
                             cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return KDBGPU_OK;
}

namespace {

// One slice of whole nodes [a, b) of the caller's CSR, as it travels to the device in one copy:
// [node_row[a..b]] [row_off[row(a)..row(b)]] [nbrs[edge(a)..edge(b))]
struct GraphSlice {
  uint32_t a, b;
  uint64_t r0, r1, e0, e1;
  size_t bytes() const { return ((size_t)(b - a) + 1 + (size_t)(r1 - r0) + 1) * sizeof(uint64_t) + (size_t)(e1 - e0) * sizeof(uint32_t); }
};
constexpr uint64_t kSliceRows = 1ull << 21;
// <= 92 MB per slice; KDBGPU_GRAPH_SLICE_NODES / _EDGES shrink the slices (tests walk the multi-slice path on small graphs)
uint64_t slice_limit(const char *env, uint64_t dflt) {
  const char *v = getenv(env);
  const long long x = v ? atoll(v) : 0;
  return x >= 1 && (uint64_t)x < dflt ? (uint64_t)x : dflt;
}

// The topology is staged ON THE DEVICE: the host walks the per-node arrays once (levels, row counts: O(n)), the
// O(edges) work — dropping nil neighbours, padding every row to its fixed degree — is graph_scatter_kernel's, fed
// slice by slice through two pinned buffers so that filling one overlaps the DMA of the other.  (The first version
// built the padded rows on the host: 4.9 s for the 2.76 GB sidecar of a 10 M-node graph, profiles/r2_cold_start.jsonl.)
int set_graph_impl(kdbgpu_index *h, uint32_t n, const int32_t *levels, const uint64_t *node_row, const uint64_t *row_off,
                   const uint32_t *nbrs, uint32_t entry, int max_level) {
  const uint32_t deg0 = (uint32_t)(2 * h->m), degu = (uint32_t)h->m;
  const size_t n1 = (size_t)n + 1;
  std::vector<uint32_t> upper_first(n1, 0u), upper_node;
  std::vector<uint8_t> upper_level;
  std::vector<int8_t> lv(n1, (int8_t)-1);
  size_t upper_rows = 0;
  for (uint32_t id = 1; id <= n; ++id) {
    const int32_t L = levels[id];
    if (L < 0) continue;
    if (L > 120) return fail(KDBGPU_ERR_INVALID, "node %u has level %d", id, L);
    if (node_row[id + 1] - node_row[id] != (uint64_t)(L + 1))
      return fail(KDBGPU_ERR_INVALID, "node %u: %llu rows for level %d", id,
                  (unsigned long long)(node_row[id + 1] - node_row[id]), L);
    lv[id] = (int8_t)L;
    upper_first[id] = (uint32_t)upper_rows;
    for (int32_t l = 1; l <= L; ++l) {
      upper_node.push_back(id);
      upper_level.push_back((uint8_t)l);
    }
    upper_rows += (size_t)L;
  }
  // slices of whole nodes within the limits above (every measure is monotone in b for a well-formed CSR)
  std::vector<GraphSlice> slices;
  size_t slice_bytes = 0;
  const uint64_t kSliceNodes = slice_limit("KDBGPU_GRAPH_SLICE_NODES", 1ull << 20);
  const uint64_t kSliceEdges = slice_limit("KDBGPU_GRAPH_SLICE_EDGES", 1ull << 24);
  for (uint32_t a = 1; a <= n;) {
    const uint64_t r0 = node_row[a], e0 = row_off[r0];
    auto fits = [&](uint32_t b) {
      const uint64_t r1 = node_row[b];
      return r1 >= r0 && r1 - r0 <= kSliceRows && row_off[r1] >= e0 && row_off[r1] - e0 <= kSliceEdges;
    };
    uint32_t lo = a, hi = (uint64_t)a + kSliceNodes < (uint64_t)n + 1 ? a + (uint32_t)kSliceNodes : n + 1;  // b in (lo, hi]
    if (!fits(a + 1))
      return fail(KDBGPU_ERR_INVALID, "node %u: row / neighbour offsets out of order or beyond %llu neighbours", a,
                  (unsigned long long)kSliceEdges);
    lo = a + 1;
    while (lo < hi) {  // largest b that fits
      const uint32_t mid = lo + (hi - lo + 1) / 2;
      if (fits(mid)) lo = mid;
      else hi = mid - 1;
    }
    GraphSlice sl{a, lo, r0, node_row[lo], e0, row_off[node_row[lo]]};
    if (sl.bytes() > slice_bytes) slice_bytes = sl.bytes();
    slices.push_back(sl);
    a = lo;
  }
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());  // drain searches queued through the asynchronous entry point
  h->has_graph = false;               // from here on the mirror is being rewritten: no graph until every row is in place
  if ((upper_rows + 1) * degu > h->upper_adj.n) {
    const size_t rows = (upper_rows + 1) + (upper_rows + 1) / 4 + 1024;
    h->upper_adj.release();
    h->upper_node.release();
    h->upper_level.release();
    CUDA_TRY(h->upper_adj.reserve(rows * degu, true));
    CUDA_TRY(h->upper_node.reserve(rows, true));
    CUDA_TRY(h->upper_level.reserve(rows, true));
  }
  DevBuf<unsigned char> d_stage[2];
  DevBuf<int> d_err;
  unsigned char *pinned[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  struct Free {
    DevBuf<unsigned char> *st;
    DevBuf<int> &er;
    unsigned char **pin;
    cudaEvent_t *ev;
    ~Free() {
      st[0].release(); st[1].release(); er.release();
      for (int i = 0; i < 2; ++i) {
        if (pin[i]) cudaFreeHost(pin[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
      }
    }
  } freer{d_stage, d_err, pinned, ev};
  const int n_buf = slices.size() > 1 ? 2 : (slices.empty() ? 0 : 1);
  for (int i = 0; i < n_buf; ++i) {
    CUDA_TRY(d_stage[i].reserve(slice_bytes + 16));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&pinned[i]), slice_bytes + 16, cudaHostAllocDefault));
    CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  CUDA_TRY(d_err.reserve(4, true));
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaMemsetAsync(h->adj0.p, 0, h->adj0.bytes(), s));
  CUDA_TRY(cudaMemsetAsync(h->upper_adj.p, 0, h->upper_adj.bytes(), s));
  if (upper_rows) {
    CUDA_TRY(cudaMemcpyAsync(h->upper_node.p, upper_node.data(), upper_rows * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->upper_level.p, upper_level.data(), upper_rows, cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaMemsetAsync(h->upper_first.p, 0, h->upper_first.bytes(), s));
  CUDA_TRY(cudaMemcpyAsync(h->upper_first.p, upper_first.data(), n1 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(h->levels.p, 0xff, (size_t)h->capacity + 1, s));
  CUDA_TRY(cudaMemcpyAsync(h->levels.p, lv.data(), n1, cudaMemcpyHostToDevice, s));
  for (size_t c = 0; c < slices.size(); ++c) {
    const GraphSlice &sl = slices[c];
    const int b = (int)(c & 1);
    CUDA_TRY(cudaEventSynchronize(ev[b]));  // the copy that last read this pinned buffer is done
    const size_t nn = (size_t)(sl.b - sl.a) + 1, nr = (size_t)(sl.r1 - sl.r0) + 1, ne = (size_t)(sl.e1 - sl.e0);
    unsigned char *p = pinned[b];
    memcpy(p, node_row + sl.a, nn * sizeof(uint64_t));
    memcpy(p + nn * sizeof(uint64_t), row_off + sl.r0, nr * sizeof(uint64_t));
    if (ne) memcpy(p + (nn + nr) * sizeof(uint64_t), nbrs + sl.e0, ne * sizeof(uint32_t));
    CUDA_TRY(cudaMemcpyAsync(d_stage[b].p, p, sl.bytes(), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(ev[b], s));
    const uint64_t *d_node_row = reinterpret_cast<const uint64_t *>(d_stage[b].p);
    CUDA_TRY(launch_graph_scatter(d_node_row, d_node_row + nn, reinterpret_cast<const uint32_t *>(d_node_row + nn + nr), sl.a,
                                  sl.b - sl.a, sl.r0, sl.e0, sl.e1, n, h->levels.p, h->upper_first.p, deg0, degu,
                                  h->adj0.p, h->upper_adj.p, d_err.p, s));
  }
  int err[3] = {0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(err, d_err.p, sizeof err, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  std::fill(h->h_levels.begin(), h->h_levels.end(), (int8_t)-1);
  std::fill(h->h_upper_first.begin(), h->h_upper_first.end(), 0u);
  std::copy(lv.begin(), lv.end(), h->h_levels.begin());
  std::copy(upper_first.begin(), upper_first.end(), h->h_upper_first.begin());
  h->upper_rows_used = (uint32_t)upper_rows;
  h->n = n;
  h->entry = entry;
  h->max_level = max_level;
  if (err[0] == 1)
    return fail(KDBGPU_ERR_INVALID, "node %d level %d has more than %u neighbours", err[1], err[2], err[2] == 0 ? deg0 : degu);
  if (err[0])
    return fail(KDBGPU_ERR_INVALID, "node %d level %d: neighbour offsets out of order", err[1], err[2]);
  h->has_graph = true;
  return KDBGPU_OK;
}

}  // namespace

int kdbgpu_set_graph(kdbgpu_index *h, uint32_t n, const int32_t *levels, const uint64_t *node_row,
                     const uint64_t *row_off, const uint32_t *nbrs, uint32_t entry, int max_level) {
  if (!h || !levels || !node_row || !row_off) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (n > h->capacity) return fail(KDBGPU_ERR_INVALID, "n %u exceeds capacity %u", n, h->capacity);
  if (max_level >= 0 && (entry == 0 || entry > n || levels[entry] < 0))
    return fail(KDBGPU_ERR_INVALID, "entry point %u is not a live node", entry);
  try {
    return set_graph_impl(h, n, levels, node_row, row_off, nbrs, entry, max_level);
  } catch (...) {  // std::bad_alloc of the per-node arrays: nothing unwinds across the C boundary
    return fail(KDBGPU_ERR_NOMEM, "out of host memory staging the topology");
  }
}

int kdbgpu_set_deleted(kdbgpu_index *h, const uint64_t *bitset, size_t words) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());  // drain searches queued through the asynchronous entry point
  const size_t need32 = h->deleted.n;
  CUDA_TRY(cudaMemsetAsync(h->deleted.p, 0, need32 * sizeof(uint32_t), h->stream));
  h->has_deleted = false;
  if (bitset && words) {
    size_t copy32 = words * 2;
    if (copy32 > need32) copy32 = need32;
    CUDA_TRY(cudaMemcpyAsync(h->deleted.p, bitset, copy32 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    for (size_t w = 0; w < words; ++w)
      if (bitset[w]) {
        h->has_deleted = true;
        break;
      }
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return KDBGPU_OK;
}

int kdbgpu_search_batch(kdbgpu_index *h, const float *queries, uint32_t nq, int k, int ef_search,
                        const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                        uint32_t *out_counts, kdbgpu_stats *stats) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (stats) memset(stats, 0, sizeof *stats);
  if (nq == 0) return KDBGPU_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (k <= 0 || k > 10000) return fail(KDBGPU_ERR_INVALID, "k %d outside 1..10000", k);  // http_handlers.go:37
  const int ef = ef_search < k ? k : ef_search;  // hnsw_index.go:2377-2380
  std::shared_lock<std::shared_mutex> lk(h->mu);
  if (!h->has_graph) return fail(KDBGPU_ERR_STATE, "kdbgpu_set_graph has not been called");
  DeviceGuard g(h->device);
  // an empty index or an empty non-nil allow-list returns [] for every query (:383-385, :443-445)
  bool allow_found = true;
  uint32_t allow_first = 0;
  if (allow) allow_first = first_set_bit(allow, allow_words, &allow_found);
  if (h->max_level < 0 || !allow_found || (allow && allow_first > h->n)) {
    memset(out_ids, 0, (size_t)nq * k * sizeof(uint32_t));
    memset(out_scores, 0, (size_t)nq * k * sizeof(double));
    memset(out_counts, 0, (size_t)nq * sizeof(uint32_t));
    return KDBGPU_OK;
  }
  const int wi = acquire_ws(h);
  kdbgpu_index::SearchWs &w = h->sws[wi];
  struct Release {
    kdbgpu_index *h;
    int wi;
    ~Release() { release_ws(h, wi); }
  } releaser{h, wi};
  cudaStream_t s = w.stream;
  CUDA_TRY(cudaStreamWaitEvent(s, w.done, 0));
  CUDA_TRY(w.q_raw.reserve((size_t)nq * h->dim));
  // results, counters and the error flag live in one device blob -> a single D2H copy
  const size_t nk = (size_t)nq * k;
  const BlobLayout BL = blob_layout(nq, k);
  const size_t o_ids = BL.o_ids, o_counts = BL.o_counts, o_stats = BL.o_stats, o_err = BL.o_err, blob_bytes = BL.bytes;
  CUDA_TRY(w.out_blob.reserve(blob_bytes));
  if (w.h_out_bytes < blob_bytes) {
    if (w.h_out) cudaFreeHost(w.h_out);
    w.h_out = nullptr;
    w.h_out_bytes = 0;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&w.h_out), blob_bytes + blob_bytes / 4, cudaHostAllocDefault));
    w.h_out_bytes = blob_bytes + blob_bytes / 4;
  }
  unsigned char *blob = w.out_blob.p;
  CUDA_TRY(cudaEventRecord(w.ev[0], s));
  CUDA_TRY(cudaMemcpyAsync(w.q_raw.p, queries, (size_t)nq * h->dim * sizeof(float), cudaMemcpyHostToDevice, s));
  const uint32_t *d_allow = nullptr;
  if (allow) {
    int rc = stage_allow(h, w.allow, allow, allow_words, s);
    if (rc) return rc;
    d_allow = w.allow.p;
  }
  {
    int prc = prepare_queries(h, w, w.q_raw.p, nq, s);
    if (prc) return prc;
  }
  CUDA_TRY(cudaEventRecord(w.ev[1], s));
  int rc = enqueue_search(h, w, w.q_prep.p, nq, k, ef, d_allow, allow_first, reinterpret_cast<uint32_t *>(blob + o_ids),
                          reinterpret_cast<double *>(blob), reinterpret_cast<uint32_t *>(blob + o_counts), s,
                          reinterpret_cast<unsigned long long *>(blob + o_stats), reinterpret_cast<int *>(blob + o_err));
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(w.ev[2], s));
  CUDA_TRY(cudaMemcpyAsync(w.h_out, blob, blob_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaEventRecord(w.ev[3], s));
  CUDA_TRY(cudaEventRecord(w.done, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(out_scores, w.h_out, nk * sizeof(double));
  memcpy(out_ids, w.h_out + o_ids, nk * sizeof(uint32_t));
  memcpy(out_counts, w.h_out + o_counts, (size_t)nq * sizeof(uint32_t));
  unsigned long long st[4];
  int err = 0;
  memcpy(st, w.h_out + o_stats, sizeof st);
  memcpy(&err, w.h_out + o_err, sizeof err);
  if (stats) {
    stats->dist_evals = st[0];
    stats->hops = st[1];
    stats->hops_l0 = st[2];
    stats->heap_pass_queries = (uint32_t)st[3];
    cudaEventElapsedTime(&stats->kernel_ms, w.ev[1], w.ev[2]);
    cudaEventElapsedTime(&stats->total_ms, w.ev[0], w.ev[3]);
  }
  if (err == KDBGPU_ERR_OVERFLOW)
    return fail(KDBGPU_ERR_OVERFLOW, "candidate heap exceeded %u entries for at least one query",
                h->ovf_cap + (uint32_t)h->tuning.cand_smem);
  return KDBGPU_OK;
}

// Reserve every launch workspace for batches of up to nq queries now, so that no search pays a
// cudaMalloc / cudaHostAlloc (both synchronise the device) on its first trip through a workspace.
int kdbgpu_prepare_search(kdbgpu_index *h, uint32_t nq, int k, int ef_search) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (nq == 0 || k <= 0 || k > 10000) return fail(KDBGPU_ERR_INVALID, "bad shape");
  const int ef = ef_search < k ? k : ef_search;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  DevIndex ix = h->dev();
  const int occ = search_occupancy(ix, ef, h->tuning);
  if (occ <= 0) return fail(KDBGPU_ERR_INVALID, "search configuration does not fit shared memory (dim=%d ef=%d)", h->dim, ef);
  const size_t blob_bytes = blob_layout(nq, k).bytes;
  const int occ_fast = search_fast_eligible(ix, ef, h->tuning) ? search_fast_occupancy(ix, ef, h->tuning) : 0;
  for (auto &w : h->sws) {
    int rc = ensure_ws(h, w, (occ_fast > occ ? occ_fast : occ) * h->num_sms);
    if (rc) return rc;
    CUDA_TRY(w.redo.reserve(nq));
    CUDA_TRY(w.q_raw.reserve((size_t)nq * h->dim));
    CUDA_TRY(w.q_prep.reserve((size_t)nq * h->stride));
    if (h->precision == KDBGPU_PRECISION_INT8) CUDA_TRY(w.qnorms.reserve(nq));
    CUDA_TRY(w.allow.reserve(((size_t)h->capacity + 1 + 31) / 32 + 2));
    CUDA_TRY(w.out_blob.reserve(blob_bytes));
    if (w.h_out_bytes < blob_bytes) {
      if (w.h_out) cudaFreeHost(w.h_out);
      w.h_out = nullptr;
      w.h_out_bytes = 0;
      CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&w.h_out), blob_bytes + blob_bytes / 4, cudaHostAllocDefault));
      w.h_out_bytes = blob_bytes + blob_bytes / 4;
    }
  }
  CUDA_TRY(cudaDeviceSynchronize());
  return KDBGPU_OK;
}

int kdbgpu_search_batch_device(kdbgpu_index *h, const float *d_queries, uint32_t nq, int k, int ef_search,
                               const uint64_t *d_allow, size_t allow_words, uint32_t allow_first_id,
                               uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (nq == 0) return KDBGPU_OK;
  if (!d_queries || !d_out_ids || !d_out_scores || !d_out_counts) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (k <= 0 || k > 10000) return fail(KDBGPU_ERR_INVALID, "k %d outside 1..10000", k);
  const int ef = ef_search < k ? k : ef_search;
  std::shared_lock<std::shared_mutex> lk(h->mu);
  if (!h->has_graph) return fail(KDBGPU_ERR_STATE, "kdbgpu_set_graph has not been called");
  if (d_allow && allow_words * 64 < (size_t)h->n + 1)
    return fail(KDBGPU_ERR_INVALID, "device allow-list must cover ids 0..%u", h->n);
  DeviceGuard g(h->device);
  // asynchronous path: the launch takes a free workspace under the same busy protocol as the host path (so
  // that concurrent callers never share visited bitsets / counters), records w.done and hands the workspace
  // back; the next user orders itself behind w.done on the device.
  const int wi = acquire_ws(h);
  kdbgpu_index::SearchWs &w = h->sws[wi];
  struct Release {
    kdbgpu_index *h;
    int wi;
    ~Release() { release_ws(h, wi); }
  } releaser{h, wi};
  {
    std::lock_guard<std::mutex> wl(h->ws_mu);
    h->last_ws = wi;
  }
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : w.stream;
  if (h->max_level < 0 || (d_allow && (allow_first_id == 0 || allow_first_id > h->n))) {
    CUDA_TRY(cudaMemsetAsync(d_out_ids, 0, (size_t)nq * k * sizeof(uint32_t), s));
    CUDA_TRY(cudaMemsetAsync(d_out_scores, 0, (size_t)nq * k * sizeof(double), s));
    CUDA_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)nq * sizeof(uint32_t), s));
    return KDBGPU_OK;
  }
  CUDA_TRY(cudaStreamWaitEvent(s, w.done, 0));
  int rc = prepare_queries(h, w, d_queries, nq, s);
  if (rc == KDBGPU_OK)
    rc = enqueue_search(h, w, w.q_prep.p, nq, k, ef, reinterpret_cast<const uint32_t *>(d_allow), allow_first_id,
                        d_out_ids, d_out_scores, d_out_counts, s);
  // whatever was queued (possibly nothing) is what the next user of the workspace must wait for
  cudaError_t re = cudaEventRecord(w.done, s);
  if (rc) return rc;
  CUDA_TRY(re);
  return KDBGPU_OK;
}

int kdbgpu_distance_batch(kdbgpu_index *h, const float *query, const uint32_t *ids, uint32_t n, double *out) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (n == 0) return KDBGPU_OK;
  if (!query || !ids || !out) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t s = h->stream;
  CUDA_TRY(h->q_prep.reserve((size_t)h->stride));
  CUDA_TRY(h->ids_tmp.reserve(n));
  CUDA_TRY(h->dist_tmp.reserve(n));
  CUDA_TRY(cudaMemsetAsync(h->q_prep.p, 0, (size_t)h->stride * sizeof(float), s));
  if (h->precision == KDBGPU_PRECISION_F32) {
    CUDA_TRY(cudaMemcpyAsync(h->q_prep.p, query, (size_t)h->dim * sizeof(float), cudaMemcpyHostToDevice, s));
  } else {
    // the prepared float32 query goes through the precision adaptation of searchInternal (:417-434)
    if (h->precision == KDBGPU_PRECISION_INT8 && h->abs_max == 0.f)
      return fail(KDBGPU_ERR_STATE, "int8 index without a trained quantizer (kdbgpu_set_quantizer)");
    CUDA_TRY(h->q_raw.reserve((size_t)h->dim));
    CUDA_TRY(h->qnorm1.reserve(1));
    CUDA_TRY(cudaMemcpyAsync(h->q_raw.p, query, (size_t)h->dim * sizeof(float), cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_convert_rows(h->q_raw.p, (size_t)h->dim, h->q_prep.p, h->stride, 1, (uint32_t)h->dim, h->kind, false,
                                 h->abs_max, h->qnorm1.p, true, s));
  }
  CUDA_TRY(cudaMemcpyAsync(h->ids_tmp.p, ids, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  DevIndex ix = h->dev();
  ix.n = h->capacity;  // any staged row may be addressed, graph or not
  CUDA_TRY(launch_distance_batch(ix, h->q_prep.p, h->qnorm1.p, h->ids_tmp.p, n, h->dist_tmp.p, s));
  CUDA_TRY(cudaMemcpyAsync(out, h->dist_tmp.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return KDBGPU_OK;
}

int kdbgpu_flat_search_batch(kdbgpu_index *h, const float *queries, uint32_t nq, int k, int mode,
                             const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                             uint32_t *out_counts, kdbgpu_stats *stats) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (stats) memset(stats, 0, sizeof *stats);
  if (nq == 0) return KDBGPU_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (k <= 0 || k > 1024) return fail(KDBGPU_ERR_INVALID, "flat k %d outside 1..1024", k);
  if (h->precision != KDBGPU_PRECISION_F32)  // BruteForceIndex holds []float32 only (vector_index.go:104-162)
    return fail(KDBGPU_ERR_INVALID, "the flat scan exists for float32 indexes only");
  const bool prefilter = (mode & KDBGPU_FLAT_PREFILTER) != 0;
  mode &= ~KDBGPU_FLAT_PREFILTER;
  if (mode != 0 && mode != 1) return fail(KDBGPU_ERR_INVALID, "mode %d", mode);
  std::shared_lock<std::shared_mutex> lk(h->mu);  // two flat calls (and the traversals) may be in flight
  if (!h->has_graph) return fail(KDBGPU_ERR_STATE, "kdbgpu_set_graph has not been called (it defines the live rows)");
  DeviceGuard g(h->device);
  if (h->n == 0) {
    memset(out_ids, 0, (size_t)nq * k * sizeof(uint32_t));
    memset(out_scores, 0, (size_t)nq * k * sizeof(double));
    memset(out_counts, 0, (size_t)nq * sizeof(uint32_t));
    return KDBGPU_OK;
  }
  const int wi = acquire_fws(h);
  kdbgpu_index::FlatWs &w = h->fws[wi];
  struct Release {
    kdbgpu_index *h;
    int wi;
    ~Release() { release_fws(h, wi); }
  } releaser{h, wi};
  // BruteForceIndex treats an empty allow-list as unfiltered (vector_index.go:132)
  const uint32_t *d_allow = nullptr;
  if (allow) {
    bool found = false;
    (void)first_set_bit(allow, allow_words, &found);
    if (found) {
      int rc = stage_allow(h, w.allow, allow, allow_words, w.stream);
      if (rc) return rc;
      d_allow = w.allow.p;
    }
  }
  return flat_search_ws(h, w, queries, nq, k, mode, prefilter, d_allow, out_ids, out_scores, out_counts, stats);
}

int kdbgpu_flat_prefilter_scores(kdbgpu_index *h, const float *queries, uint32_t nq, int mode, float *out_scores,
                                 float *out_bound) {
  if (!h || !queries || !out_scores) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (mode != 0 && mode != 1) return fail(KDBGPU_ERR_INVALID, "mode %d", mode);
  if (h->precision != KDBGPU_PRECISION_F32) return fail(KDBGPU_ERR_INVALID, "the flat scan exists for float32 indexes only");
  if (!h->has_graph || h->n == 0) return fail(KDBGPU_ERR_STATE, "no rows staged");
  if (nq == 0) return KDBGPU_OK;
  if ((uint64_t)nq * h->n > (1ull << 30)) return fail(KDBGPU_ERR_INVALID, "validation hook: nq * n too large");
  std::shared_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const int wi = acquire_fws(h);
  kdbgpu_index::FlatWs &w = h->fws[wi];
  struct Release {
    kdbgpu_index *h;
    int wi;
    ~Release() { release_fws(h, wi); }
  } releaser{h, wi};
  cudaStream_t s = w.stream;
  TcPlan P;
  int rc = tc_prepare(h, w, mode, nullptr, &P, s);
  if (rc) return rc;
  const uint32_t c_pad = (nq + P.bm - 1) / P.bm * P.bm;
  rc = tc_stage_queries(h, w, queries, nq, c_pad, mode, P, s);
  if (rc) return rc;
  DevBuf<float> S;
  CUDA_TRY(S.reserve((size_t)nq * h->n));
  CUDA_TRY(w.t_gmin.reserve(1));
  CUDA_TRY(w.t_theta.reserve(c_pad));
  CUDA_TRY(w.t_bound.reserve(c_pad));
  FlatTcLaunch L = tc_launch_desc(h, w, P, nq, c_pad);
  L.epi = 0;
  L.S = S.p;
  L.ldS = h->n;
  cudaError_t e = launch_flat_tc(L, s);
  // bound[q] through the threshold kernel on a dummy one-group input
  if (e == cudaSuccess) e = cudaMemsetAsync(w.t_gmin.p, 0, sizeof(float), s);
  if (e == cudaSuccess)
    e = launch_tc_threshold(w.t_gmin.p, 0, nq, 1, w.tq_sumsq.p, w.tq_resid2.p, h->x_max.p, P.alpha, P.use_norm, P.dp,
                            w.t_theta.p, w.t_bound.p, s);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out_scores, S.p, (size_t)nq * h->n * sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && out_bound)
    e = cudaMemcpyAsync(out_bound, w.t_bound.p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  S.release();
  CUDA_TRY(e);
  return KDBGPU_OK;
}

int kdbgpu_merge_topk_device(kdbgpu_index *h, int n_shards, uint32_t nq, int k, const uint32_t *d_ids,
                             const double *d_scores, const uint32_t *d_counts, uint32_t *d_out_ids,
                             double *d_out_scores, uint32_t *d_out_counts, void *stream) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (n_shards <= 0 || n_shards > 16) return fail(KDBGPU_ERR_INVALID, "n_shards %d outside 1..16", n_shards);
  if (k <= 0) return fail(KDBGPU_ERR_INVALID, "k %d", k);
  if (nq == 0) return KDBGPU_OK;
  if (!d_ids || !d_scores || !d_counts || !d_out_ids || !d_out_scores || !d_out_counts)
    return fail(KDBGPU_ERR_INVALID, "NULL argument");
  DeviceGuard g(h->device);
  cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : h->stream;
  CUDA_TRY(launch_merge_topk(n_shards, nq, k, d_ids, d_scores, d_counts, d_out_ids, d_out_scores, d_out_counts, s));
  return KDBGPU_OK;
}

int kdbgpu_last_search_stats(kdbgpu_index *h, kdbgpu_stats *stats) {
  if (!h || !stats) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  memset(stats, 0, sizeof *stats);
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  if (h->last_ws < 0 || !h->sws[h->last_ws].stats.p) return KDBGPU_OK;
  unsigned long long st[4] = {0, 0, 0, 0};
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(st, h->sws[h->last_ws].stats.p, sizeof st, cudaMemcpyDeviceToHost));
  int err = 0;
  if (h->sws[h->last_ws].err_flag.p)
    CUDA_TRY(cudaMemcpy(&err, h->sws[h->last_ws].err_flag.p, sizeof err, cudaMemcpyDeviceToHost));
  stats->dist_evals = st[0];
  stats->hops = st[1];
  stats->hops_l0 = st[2];
  stats->heap_pass_queries = (uint32_t)st[3];
  if (err == KDBGPU_ERR_OVERFLOW)  // such a query returned count 0; the reference's candidate heap is unbounded
    return fail(KDBGPU_ERR_OVERFLOW, "candidate heap exceeded %u entries for at least one query of the last launch",
                h->ovf_cap + (uint32_t)h->tuning.cand_smem);
  return KDBGPU_OK;
}

int kdbgpu_index_device(const kdbgpu_index *h) { return h ? h->device : -1; }
int kdbgpu_index_dim(const kdbgpu_index *h) { return h ? h->dim : -1; }
int kdbgpu_index_m(const kdbgpu_index *h) { return h ? h->m : -1; }
uint32_t kdbgpu_index_count(const kdbgpu_index *h) { return h ? h->n : 0; }
uint64_t kdbgpu_index_device_bytes(const kdbgpu_index *h) {
  if (!h) return 0;
  uint64_t ws = 0;
  for (const auto &w : h->sws)
    ws += w.visited.bytes() + w.cand_overflow.bytes() + w.q_raw.bytes() + w.q_prep.bytes() + w.qnorms.bytes() +
          w.allow.bytes() + w.out_blob.bytes() + w.redo.bytes();
  return h->vecs.bytes() + h->norms.bytes() + h->adj0.bytes() + h->upper_adj.bytes() + h->upper_first.bytes() +
         h->upper_node.bytes() + h->upper_level.bytes() + h->deleted.bytes() + h->levels.bytes() + h->visited.bytes() +
         h->cand_overflow.bytes() + ws + h->flat_dist.bytes() + h->q_raw.bytes() + h->q_prep.bytes() + h->out_ids.bytes() +
         h->out_scores.bytes() + h->allow.bytes() + h->x_bf16.bytes() + h->conv_tmp.bytes() + h->fws[0].bytes() +
         h->fws[1].bytes();
}
int kdbgpu_search_concurrency(kdbgpu_index *h, int k, int ef_search) {
  if (!h) return 0;
  DeviceGuard g(h->device);
  const int ef = ef_search < k ? k : ef_search;
  const DevIndex ix = h->dev();
  int occ = search_occupancy(ix, ef, h->tuning);
  if (search_fast_eligible(ix, ef, h->tuning)) {  // the pass that answers (nearly) every query of such an index
    const int occ_fast = search_fast_occupancy(ix, ef, h->tuning);
    if (occ_fast > occ) occ = occ_fast;
  }
  return occ * h->num_sms;
}

}  // extern "C"

// ---- graph construction on the device ---------------------------------------
namespace {

// randomLevel (hnsw_index.go:2616-2625): floor(-ln(u) * 1/ln(m)), capped at current max + 1
int random_level(double u, int m, int current_max) {
  const double ml = 1.0 / log((double)m);
  const double f = floor(-log(u) * ml);
  int level = (f > 1e9 || f != f) ? 1000000000 : (int)f;
  if (level > current_max + 1) return current_max + 1;
  return level;
}

int grow_upper(kdbgpu_index *h, size_t need_rows, cudaStream_t s) {
  if (need_rows <= h->upper_node.n && need_rows * (size_t)h->m <= h->upper_adj.n) return KDBGPU_OK;
  const size_t rows = need_rows + need_rows / 2 + 4096;
  DevBuf<uint32_t> adj, node;
  DevBuf<uint8_t> lvl;
  CUDA_TRY(adj.reserve(rows * (size_t)h->m, true));
  CUDA_TRY(node.reserve(rows, true));
  CUDA_TRY(lvl.reserve(rows, true));
  const size_t used = h->upper_rows_used;
  if (used) {
    CUDA_TRY(cudaMemcpyAsync(adj.p, h->upper_adj.p, used * (size_t)h->m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(node.p, h->upper_node.p, used * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lvl.p, h->upper_level.p, used, cudaMemcpyDeviceToDevice, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  h->upper_adj.release();
  h->upper_node.release();
  h->upper_level.release();
  h->upper_adj = adj;
  h->upper_node = node;
  h->upper_level = lvl;
  return KDBGPU_OK;
}

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// rows: `count` raw vectors, on the host (row_stride == dim) or on the device
int add_batch_impl(kdbgpu_index *h, uint32_t count, const float *rows, size_t row_stride, bool rows_on_device,
                   const double *level_draws, int ef_const) {
  KDB_NOTHROW_BEGIN
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (count == 0) return KDBGPU_OK;
  if (!rows || !level_draws) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (ef_const <= 0) ef_const = 200;  // hnsw.New default (hnsw_index.go:143-145)
  if (ef_const > 2048) return fail(KDBGPU_ERR_INVALID, "ef_const %d too large (max 2048)", ef_const);
  if (h->m < 2) return fail(KDBGPU_ERR_INVALID, "construction needs m >= 2");
  if (h->precision == KDBGPU_PRECISION_INT8 && h->abs_max == 0.f)
    return fail(KDBGPU_ERR_STATE, "int8 index without a trained quantizer (kdbgpu_set_quantizer / kdbgpu_train_quantizer)");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  if ((uint64_t)h->n + count > h->capacity)
    return fail(KDBGPU_ERR_INVALID, "batch of %u does not fit: %u of %u ids used", count, h->n, h->capacity);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());  // drain searches queued through the asynchronous entry point
  cudaStream_t s = h->stream;
  h->tc_valid = false;
  const uint32_t start_id = h->n + 1;
  const bool sequential = (uint64_t)h->n < (uint64_t)ef_const;  // :1502-1513
  // stage the rows: cosine vectors are normalised exactly as normalize() (:1557-1559, :485-493)
  const float *d_src = rows;
  if (!rows_on_device) {
    CUDA_TRY(h->q_raw.reserve((size_t)count * h->dim));
    CUDA_TRY(cudaMemcpyAsync(h->q_raw.p, rows, (size_t)count * h->dim * sizeof(float), cudaMemcpyHostToDevice, s));
    d_src = h->q_raw.p;
    row_stride = (size_t)h->dim;
  }
  if (h->precision == KDBGPU_PRECISION_F32) {
    CUDA_TRY(launch_prep_queries(d_src, row_stride, h->vecs.p + (size_t)start_id * h->stride, count, (uint32_t)h->dim,
                                 h->stride, h->metric, s));
  } else {
    // float16 / int8 rows are converted, never normalised (:497-520, :1553-1577); int8 norms alongside
    CUDA_TRY(launch_convert_rows(d_src, row_stride, h->vecs.p + (size_t)start_id * h->row_words, h->row_words, count,
                                 (uint32_t)h->dim, h->kind, false, h->abs_max,
                                 h->norms.p ? h->norms.p + start_id : nullptr, false, s));
  }
  // levels (:647, :1738) and upper-row allocation
  const int pre_max = h->max_level;
  const uint32_t pre_entry = h->entry;
  std::vector<int8_t> lv(count);
  std::vector<uint32_t> uf(count, 0u), new_upper_node;
  std::vector<uint8_t> new_upper_level;
  int cur_max = pre_max;
  uint32_t new_entry = pre_entry;
  uint32_t cursor = h->upper_rows_used;
  for (uint32_t i = 0; i < count; ++i) {
    int L = random_level(level_draws[i], h->m, sequential ? cur_max : pre_max);
    if (L > 120) L = 120;
    lv[i] = (int8_t)L;
    uf[i] = cursor;
    for (int l = 1; l <= L; ++l) {
      new_upper_node.push_back(start_id + i);
      new_upper_level.push_back((uint8_t)l);
    }
    cursor += (uint32_t)L;
    if (cur_max == -1) {  // first node of an empty index (:658-670)
      cur_max = L;
      new_entry = start_id + i;
    } else if (L > cur_max) {  // :793-801 / phase 4 :2066-2080
      cur_max = L;
      new_entry = start_id + i;
    }
  }
  int rc = grow_upper(h, (size_t)cursor + 1, s);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->levels.p + start_id, lv.data(), count, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->upper_first.p + start_id, uf.data(), (size_t)count * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(h->adj0.p + (size_t)start_id * 2 * h->m, 0, (size_t)count * 2 * h->m * sizeof(uint32_t), s));
  const uint32_t n_new_upper = cursor - h->upper_rows_used;
  if (n_new_upper) {
    CUDA_TRY(cudaMemsetAsync(h->upper_adj.p + (size_t)h->upper_rows_used * h->m, 0,
                             (size_t)n_new_upper * h->m * sizeof(uint32_t), s));
    CUDA_TRY(cudaMemcpyAsync(h->upper_node.p + h->upper_rows_used, new_upper_node.data(),
                             (size_t)n_new_upper * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->upper_level.p + h->upper_rows_used, new_upper_level.data(), n_new_upper,
                             cudaMemcpyHostToDevice, s));
  }
  // the new nodes are registered (visible to id-range checks) but unlinked
  h->n = start_id + count - 1;
  // a recoverable failure below (no memory for a workspace, a shape that does not fit) must leave the mirror as it
  // was before the call: the ids go back to nil and the id range shrinks again; the host-side levels, upper-row
  // cursor, entry point and max level are only committed at the end
  struct Rollback {
    kdbgpu_index *h;
    uint32_t start, count;
    bool armed;
    ~Rollback() {
      if (!armed) return;
      h->n = start - 1;
      cudaMemsetAsync(h->levels.p + start, 0xff, count, h->stream);  // level -1 = nil node
      cudaStreamSynchronize(h->stream);
      (void)cudaGetLastError();
    }
  } rollback{h, start_id, count, true};
  DevIndex ix = h->dev();
  ix.entry = pre_entry;
  ix.max_level = pre_max;
  SearchTuning bt;  // construction uses its own fixed shape (build.cu)
  bt.cand_smem = h->tuning.cand_smem;
  const int occ = sequential ? 1 : build_search_occupancy(ix, ef_const, (uint32_t)bt.cand_smem);
  if (occ <= 0) return fail(KDBGPU_ERR_INVALID, "construction search does not fit shared memory (ef_const=%d)", ef_const);
  const int full_grid = occ * h->num_sms;
  rc = ensure_search_workspace(h, full_grid > h->ws_grid ? full_grid : h->ws_grid);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(h->work_counter.p, 0, sizeof(uint32_t), s));
  CUDA_TRY(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), s));
  CUDA_TRY(cudaMemsetAsync(h->stats.p, 0, 4 * sizeof(unsigned long long), s));
  SearchArgs a{};
  a.queries = nullptr;
  a.nq = count;
  a.k = ef_const;
  a.ef = ef_const;
  a.allow = nullptr;
  a.allow_entry = 0;
  a.out_ids = nullptr;
  a.out_scores = nullptr;
  a.out_counts = nullptr;
  a.visited = h->visited.p;
  a.vis_words = h->vis_words;
  a.cand_overflow = h->cand_overflow.p;
  a.ovf_cap = h->ovf_cap;
  a.cand_smem = (uint32_t)bt.cand_smem;
  a.stats = h->stats.p;
  a.work_counter = h->work_counter.p;
  a.err_flag = h->err_flag.p;
  CUDA_TRY(h->b_scalars.reserve(8, true));
  if (sequential) {
    if (seq_add_smem_bytes(ix, ef_const, a.cand_smem) > 227 * 1024)
      return fail(KDBGPU_ERR_INVALID, "sequential construction does not fit shared memory");
    const uint32_t io[2] = {pre_entry, (uint32_t)(pre_max + 1)};
    CUDA_TRY(cudaMemcpyAsync(h->b_scalars.p + 2, io, sizeof io, cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_seq_add(ix, a, start_id, count, ef_const, h->adj0.p, h->upper_adj.p, h->b_scalars.p + 2, s));
  } else {
    std::vector<uint32_t> out_off(count), slot_node;
    std::vector<uint8_t> slot_level;
    uint32_t n_slots = 0;
    for (uint32_t i = 0; i < count; ++i) {
      out_off[i] = n_slots;
      const int top = lv[i] < pre_max ? lv[i] : pre_max;
      for (int l = 0; l <= top; ++l) {
        slot_node.push_back(start_id + i);
        slot_level.push_back((uint8_t)l);
      }
      n_slots += (uint32_t)(top + 1);
    }
    const uint32_t up_base = h->capacity + 1;
    const uint32_t n_rows = up_base + cursor;
    CUDA_TRY(h->b_out_off.reserve(count));
    CUDA_TRY(h->b_slot_node.reserve(n_slots));
    CUDA_TRY(h->b_slot_level.reserve(n_slots));
    CUDA_TRY(h->b_cand_ids.reserve((size_t)n_slots * ef_const));
    CUDA_TRY(h->b_cand_cnt.reserve(n_slots));
    CUDA_TRY(h->b_row_cnt.reserve((size_t)n_rows + 2));
    CUDA_TRY(h->b_row_off.reserve((size_t)n_rows + 2));
    CUDA_TRY(h->b_active.reserve((size_t)n_rows + 2));
    CUDA_TRY(h->b_srcs.reserve((size_t)n_slots * ef_const * 2));
    const int commit_grid = h->num_sms * 6;
    const uint32_t scratch_cap = next_pow2(2u * (uint32_t)h->m + count + 1);
    CUDA_TRY(h->b_scratch_d.reserve((size_t)commit_grid * scratch_cap));
    CUDA_TRY(h->b_scratch_ids.reserve((size_t)commit_grid * scratch_cap));
    CUDA_TRY(cudaMemcpyAsync(h->b_out_off.p, out_off.data(), (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->b_slot_node.p, slot_node.data(), (size_t)n_slots * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->b_slot_level.p, slot_level.data(), n_slots, cudaMemcpyHostToDevice, s));
    BuildLaunch L;
    L.start_id = start_id;
    L.count = count;
    L.pre_entry = pre_entry;
    L.pre_max = pre_max;
    L.efc = ef_const;
    L.n_slots = n_slots;
    L.up_base = up_base;
    L.n_rows = n_rows;
    L.out_off = h->b_out_off.p;
    L.slot_node = h->b_slot_node.p;
    L.slot_level = h->b_slot_level.p;
    L.cand_ids = h->b_cand_ids.p;
    L.cand_cnt = h->b_cand_cnt.p;
    L.row_cnt = h->b_row_cnt.p;
    L.row_off = h->b_row_off.p;
    L.srcs = h->b_srcs.p;
    L.active = h->b_active.p;
    L.n_active = h->b_scalars.p;
    L.work_counter = h->b_scalars.p + 1;
    L.adj0 = h->adj0.p;
    L.upper_adj = h->upper_adj.p;
    L.upper_node = h->upper_node.p;
    L.upper_level = h->upper_level.p;
    L.scratch_d = h->b_scratch_d.p;
    L.scratch_ids = h->b_scratch_ids.p;
    L.scratch_cap = scratch_cap;
    L.commit_grid = commit_grid;
    int grid = full_grid;
    if ((uint32_t)grid > count) grid = (int)count;
    CUDA_TRY(launch_add_batch(ix, a, L, grid, s));
  }
  int err = 0;
  uint32_t io_back[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(&err, h->err_flag.p, sizeof err, cudaMemcpyDeviceToHost, s));
  if (sequential) CUDA_TRY(cudaMemcpyAsync(io_back, h->b_scalars.p + 2, sizeof io_back, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  // commit host-side state
  rollback.armed = false;
  for (uint32_t i = 0; i < count; ++i) {
    h->h_levels[start_id + i] = lv[i];
    h->h_upper_first[start_id + i] = uf[i];
  }
  h->upper_rows_used = cursor;
  h->entry = new_entry;
  h->max_level = cur_max;
  h->has_graph = true;
  if (sequential && (io_back[0] != new_entry || (int)io_back[1] - 1 != cur_max))
    return fail(KDBGPU_ERR_STATE, "entry point bookkeeping diverged (device %u/%d, host %u/%d)", io_back[0],
                (int)io_back[1] - 1, new_entry, cur_max);
  if (err == KDBGPU_ERR_OVERFLOW) return fail(KDBGPU_ERR_OVERFLOW, "a construction-time list exceeded its bound");
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

}  // namespace

extern "C" {

int kdbgpu_add_batch(kdbgpu_index *h, uint32_t count, const float *rows, const double *level_draws, int ef_const) {
  return add_batch_impl(h, count, rows, h ? (size_t)h->dim : 0, false, level_draws, ef_const);
}

int kdbgpu_add_batch_device(kdbgpu_index *h, uint32_t count, const float *d_rows, size_t row_stride,
                            const double *level_draws, int ef_const) {
  if (h && row_stride < (size_t)h->dim) return fail(KDBGPU_ERR_INVALID, "row_stride smaller than dim");
  return add_batch_impl(h, count, d_rows, row_stride, true, level_draws, ef_const);
}

// ---- incremental mirror refresh: follow the CPU index's Add / Delete / Vacuum /
// ---- Refine without re-staging the whole topology ------------------------------------------------
int kdbgpu_register_nodes(kdbgpu_index *h, uint32_t first_id, uint32_t count, const int32_t *levels) {
  KDB_NOTHROW_BEGIN
  if (!h || (!levels && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  if (first_id != h->n + 1) return fail(KDBGPU_ERR_INVALID, "nodes are registered in id order: expected %u, got %u", h->n + 1, first_id);
  if ((uint64_t)h->n + count > h->capacity)
    return fail(KDBGPU_ERR_INVALID, "%u more nodes do not fit: %u of %u ids used", count, h->n, h->capacity);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  cudaStream_t s = h->stream;
  std::vector<int8_t> lv(count);
  std::vector<uint32_t> uf(count, 0u), up_node;
  std::vector<uint8_t> up_level;
  uint32_t cursor = h->upper_rows_used;
  for (uint32_t i = 0; i < count; ++i) {
    const int32_t L = levels[i];
    if (L < -1 || L > 120) return fail(KDBGPU_ERR_INVALID, "node %u has level %d", first_id + i, L);
    lv[i] = (int8_t)L;  // -1 registers a nil slot (an id the reference burnt, e.g. a failed insert)
    uf[i] = cursor;
    for (int l = 1; l <= L; ++l) {
      up_node.push_back(first_id + i);
      up_level.push_back((uint8_t)l);
    }
    if (L > 0) cursor += (uint32_t)L;
  }
  int rc = grow_upper(h, (size_t)cursor + 1, s);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->levels.p + first_id, lv.data(), count, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->upper_first.p + first_id, uf.data(), (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(h->adj0.p + (size_t)first_id * 2 * h->m, 0, (size_t)count * 2 * h->m * sizeof(uint32_t), s));
  const uint32_t n_new_upper = cursor - h->upper_rows_used;
  if (n_new_upper) {
    CUDA_TRY(cudaMemsetAsync(h->upper_adj.p + (size_t)h->upper_rows_used * h->m, 0,
                             (size_t)n_new_upper * h->m * sizeof(uint32_t), s));
    CUDA_TRY(cudaMemcpyAsync(h->upper_node.p + h->upper_rows_used, up_node.data(), (size_t)n_new_upper * sizeof(uint32_t),
                             cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->upper_level.p + h->upper_rows_used, up_level.data(), n_new_upper, cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  for (uint32_t i = 0; i < count; ++i) {
    h->h_levels[first_id + i] = lv[i];
    h->h_upper_first[first_id + i] = uf[i];
  }
  h->upper_rows_used = cursor;
  h->n = first_id + count - 1;
  h->has_graph = true;
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

int kdbgpu_patch_rows(kdbgpu_index *h, uint32_t count, const uint32_t *ids, const int32_t *row_levels,
                      const uint64_t *row_off, const uint32_t *nbrs) {
  KDB_NOTHROW_BEGIN
  if (!h || (count && (!ids || !row_levels || !row_off))) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  if (!h->has_graph) return fail(KDBGPU_ERR_STATE, "no topology staged yet");
  const uint32_t deg0 = (uint32_t)(2 * h->m), degu = (uint32_t)h->m;
  // two groups (level-0 rows / upper rows); nil and out-of-range neighbours are dropped exactly as
  // kdbgpu_set_graph does (the reference skips them with no side effect, hnsw_index.go:2553-2561)
  std::vector<uint32_t> dst[2], off[2], cnt[2], src;
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t id = ids[i];
    const int32_t l = row_levels[i];
    if (id == 0 || id > h->n || h->h_levels[id] < 0) return fail(KDBGPU_ERR_INVALID, "patch %u: node %u is not a live node", i, id);
    if (l < 0 || l > h->h_levels[id]) return fail(KDBGPU_ERR_INVALID, "patch %u: node %u has no level %d", i, id, l);
    const uint32_t cap = l == 0 ? deg0 : degu;
    const int gsel = l == 0 ? 0 : 1;
    const uint32_t start = (uint32_t)src.size();
    for (uint64_t e = row_off[i]; e < row_off[i + 1]; ++e) {
      const uint32_t nb = nbrs[e];
      if (nb == 0 || nb > h->n || h->h_levels[nb] < 0) continue;
      if (src.size() - start >= cap) return fail(KDBGPU_ERR_INVALID, "node %u level %d has more than %u neighbours", id, l, cap);
      src.push_back(nb);
    }
    dst[gsel].push_back(l == 0 ? id : h->h_upper_first[id] + (uint32_t)(l - 1));
    off[gsel].push_back(start);
    cnt[gsel].push_back((uint32_t)src.size() - start);
  }
  if (src.empty()) src.push_back(0u);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());  // searches in flight finish on the old rows
  cudaStream_t s = h->stream;
  DevBuf<uint32_t> d_src, d_meta;
  struct Free {
    DevBuf<uint32_t> &a, &b;
    ~Free() { a.release(); b.release(); }
  } freer{d_src, d_meta};
  CUDA_TRY(d_src.reserve(src.size()));
  CUDA_TRY(cudaMemcpyAsync(d_src.p, src.data(), src.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  const size_t n0 = dst[0].size(), n1 = dst[1].size();
  CUDA_TRY(d_meta.reserve(3 * (n0 + n1) + 1));
  uint32_t *m = d_meta.p;
  for (int gsel = 0; gsel < 2; ++gsel) {
    const size_t n = dst[gsel].size();
    if (!n) continue;
    CUDA_TRY(cudaMemcpyAsync(m, dst[gsel].data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m + n, off[gsel].data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(m + 2 * n, cnt[gsel].data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_patch_rows(gsel == 0 ? h->adj0.p : h->upper_adj.p, gsel == 0 ? deg0 : degu, m, m + n, m + 2 * n,
                               d_src.p, (uint32_t)n, s));
    m += 3 * n;
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

int kdbgpu_remove_nodes(kdbgpu_index *h, uint32_t count, const uint32_t *ids) {
  if (!h || (count && !ids)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  for (uint32_t i = 0; i < count; ++i)
    if (ids[i] == 0 || ids[i] > h->n) return fail(KDBGPU_ERR_INVALID, "node %u outside 1..%u", ids[i], h->n);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  cudaStream_t s = h->stream;
  const int8_t nil = -1;
  for (uint32_t i = 0; i < count; ++i) {  // nodes[deadID] = nil (optimizer.go:270)
    const int old_level = h->h_levels[ids[i]];
    if (old_level > 0 && h->upper_adj.p && ids[i] < h->h_upper_first.size())  // its upper rows go with it
      CUDA_TRY(cudaMemsetAsync(h->upper_adj.p + (size_t)h->h_upper_first[ids[i]] * h->m, 0,
                               (size_t)old_level * h->m * sizeof(uint32_t), s));
    h->h_levels[ids[i]] = -1;
    CUDA_TRY(cudaMemcpyAsync(h->levels.p + ids[i], &nil, 1, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(h->adj0.p + (size_t)ids[i] * 2 * h->m, 0, (size_t)2 * h->m * sizeof(uint32_t), s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return KDBGPU_OK;
}

int kdbgpu_set_entry(kdbgpu_index *h, uint32_t entry, int max_level) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  if (max_level >= 0 && (entry == 0 || entry > h->n || h->h_levels[entry] < 0))
    return fail(KDBGPU_ERR_INVALID, "entry point %u is not a live node", entry);
  if (max_level >= 0 && max_level > h->h_levels[entry])
    return fail(KDBGPU_ERR_INVALID, "entry point %u has level %d, not %d", entry, (int)h->h_levels[entry], max_level);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  h->entry = max_level >= 0 ? entry : 0;
  h->max_level = max_level;
  return KDBGPU_OK;
}

int kdbgpu_get_graph_sizes(kdbgpu_index *h, uint32_t *n, uint64_t *n_rows, uint64_t *n_edges, uint32_t *entry,
                           int *max_level) {
  KDB_NOTHROW_BEGIN
  if (!h || !n || !n_rows || !n_edges || !entry || !max_level) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const uint32_t deg0 = (uint32_t)(2 * h->m), degu = (uint32_t)h->m;
  std::vector<uint32_t> adj0((size_t)(h->n + 1) * deg0), upper((size_t)(h->upper_rows_used + 1) * degu);
  CUDA_TRY(cudaMemcpy(adj0.data(), h->adj0.p, adj0.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(upper.data(), h->upper_adj.p, upper.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  uint64_t rows = 0, edges = 0;
  for (uint32_t id = 1; id <= h->n; ++id) {
    const int L = h->h_levels[id];
    if (L < 0) continue;
    rows += (uint64_t)(L + 1);
    for (int l = 0; l <= L; ++l) {
      const uint32_t cap = l == 0 ? deg0 : degu;
      const uint32_t *r = l == 0 ? &adj0[(size_t)id * deg0] : &upper[((size_t)h->h_upper_first[id] + (size_t)(l - 1)) * degu];
      for (uint32_t i = 0; i < cap && r[i]; ++i) edges++;
    }
  }
  *n = h->n;
  *n_rows = rows;
  *n_edges = edges;
  *entry = h->entry;
  *max_level = h->max_level;
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

int kdbgpu_get_graph(kdbgpu_index *h, int32_t *levels, uint64_t *node_row, uint64_t *row_off, uint32_t *nbrs) {
  KDB_NOTHROW_BEGIN
  if (!h || !levels || !node_row || !row_off || !nbrs) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const uint32_t deg0 = (uint32_t)(2 * h->m), degu = (uint32_t)h->m;
  std::vector<uint32_t> adj0((size_t)(h->n + 1) * deg0), upper((size_t)(h->upper_rows_used + 1) * degu);
  CUDA_TRY(cudaMemcpy(adj0.data(), h->adj0.p, adj0.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(upper.data(), h->upper_adj.p, upper.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  uint64_t r = 0, e = 0;
  levels[0] = -1;
  node_row[0] = 0;
  for (uint32_t id = 1; id <= h->n; ++id) {
    const int L = h->h_levels[id];
    levels[id] = L;
    node_row[id] = r;
    for (int l = 0; l <= L; ++l) {
      const uint32_t cap = l == 0 ? deg0 : degu;
      const uint32_t *row = l == 0 ? &adj0[(size_t)id * deg0] : &upper[((size_t)h->h_upper_first[id] + (size_t)(l - 1)) * degu];
      row_off[r++] = e;
      for (uint32_t i = 0; i < cap && row[i]; ++i) nbrs[e++] = row[i];
    }
  }
  node_row[h->n + 1] = r;
  row_off[r] = e;
  return KDBGPU_OK;
  KDB_NOTHROW_END
}

int kdbgpu_download_vectors(kdbgpu_index *h, uint32_t first_id, uint32_t count, float *rows) {
  if (!h || (!rows && count)) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  if ((uint64_t)first_id + count - 1 > h->capacity) return fail(KDBGPU_ERR_INVALID, "id range outside capacity");
  if (h->precision != KDBGPU_PRECISION_F32)
    return fail(KDBGPU_ERR_INVALID, "float32 rows exist for float32 indexes only (use kdbgpu_download_rows_raw)");
  if (count == 0) return KDBGPU_OK;
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaMemcpy2D(rows, (size_t)h->dim * sizeof(float), h->vecs.p + (size_t)first_id * h->stride,
                        (size_t)h->stride * sizeof(float), (size_t)h->dim * sizeof(float), count,
                        cudaMemcpyDeviceToHost));
  return KDBGPU_OK;
}

}  // extern "C"

extern "C" {
// test/tuning hook (not part of the reference-facing surface): CTA shape of the traversal kernel
int kdbgpu_set_fast_path(kdbgpu_index *h, int on) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  h->tuning.fast = on < 0 ? 0 : (on > 2 ? 2 : on);
  return KDBGPU_OK;
}

int kdbgpu_set_candidate_bound(kdbgpu_index *h, uint32_t spill_entries) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (spill_entries == 0 || spill_entries > (1u << 24)) return fail(KDBGPU_ERR_INVALID, "spill_entries outside 1..2^24");
  std::unique_lock<std::shared_mutex> lk(h->mu);
  DeviceGuard g(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  h->ovf_cap = spill_entries;
  h->ws_grid = 0;  // every workspace re-sizes its spill area on next use
  for (auto &w : h->sws) w.grid = 0;
  return KDBGPU_OK;
}

int kdbgpu_set_tuning(kdbgpu_index *h, int slots, int cand_smem, int max_ctas_per_sm) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  SearchTuning t = h->tuning;
  if (slots > 0) t.slots = slots;
  if (cand_smem > 0) t.cand_smem = cand_smem;
  if (max_ctas_per_sm >= 0) t.max_ctas_per_sm = max_ctas_per_sm;
  if (!search_slots_supported(t.slots)) return fail(KDBGPU_ERR_INVALID, "unsupported slot count %d", t.slots);
  std::unique_lock<std::shared_mutex> lk(h->mu);
  h->tuning = t;
  return KDBGPU_OK;
}

int kdbgpu_set_idle_slots(kdbgpu_index *h, int slots_idle) {
  if (!h) return fail(KDBGPU_ERR_INVALID, "NULL handle");
  if (slots_idle != 0 && !search_slots_supported(slots_idle))
    return fail(KDBGPU_ERR_INVALID, "unsupported slot count %d", slots_idle);
  std::unique_lock<std::shared_mutex> lk(h->mu);
  h->tuning.slots_idle = slots_idle;
  return KDBGPU_OK;
}

}  // extern "C"

// ---- host staging memory --------------------------------------------------------------------------
namespace {
std::mutex g_pinned_mu;
std::vector<void *> g_pinned;  // pointers handed out by cudaHostAlloc (the rest came from aligned_alloc)
}  // namespace

extern "C" {

int kdbgpu_host_alloc(void **out, size_t bytes) {
  if (!out) return fail(KDBGPU_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (bytes == 0) bytes = 1;
  void *p = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 &&
      cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess && p) {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned.push_back(p);
    *out = p;
    return KDBGPU_OK;
  }
  (void)cudaGetLastError();
  p = aligned_alloc(64, (bytes + 63) & ~(size_t)63);
  if (!p) return fail(KDBGPU_ERR_NOMEM, "host allocation of %zu bytes failed", bytes);
  *out = p;
  return KDBGPU_OK;
}

void kdbgpu_host_free(void *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    for (size_t i = 0; i < g_pinned.size(); ++i)
      if (g_pinned[i] == p) {
        g_pinned[i] = g_pinned.back();
        g_pinned.pop_back();
        cudaFreeHost(p);
        (void)cudaGetLastError();
        return;
      }
  }
  free(p);
}

}  // extern "C"

