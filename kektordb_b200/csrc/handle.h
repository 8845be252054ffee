// handle.h — host-side internals shared by the translation units behind the C ABI (api.cu, shard.cu):
// the mirror handle, its launch workspaces, and the helpers that queue one traversal / flat scan.
// Host code only; nothing here crosses the C boundary.
#pragma once
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "kdb_internal.cuh"

namespace kdb {
// records the message kdbgpu_last_error() returns on this thread; returns `code`
int set_error(int code, const char *fmt, ...);
}  // namespace kdb

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      (void)cudaGetLastError();                                                                     \
      return kdb::set_error(_e == cudaErrorMemoryAllocation ? KDBGPU_ERR_NOMEM : KDBGPU_ERR_CUDA, "%s: %s", #expr, \
                  cudaGetErrorString(_e));                                                          \
    }                                                                                               \
  } while (0)

namespace kdb {
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t reserve(size_t want, bool zero = false) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T));
    if (e != cudaSuccess) return e;
    n = want;
    if (zero) e = cudaMemset(p, 0, want * sizeof(T));
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};
}  // namespace kdb

struct kdbgpu_index {
  template <typename T>
  using DevBuf = kdb::DevBuf<T>;
  using HeapEntry = kdb::HeapEntry;
  using SearchTuning = kdb::SearchTuning;
  using DevIndex = kdb::DevIndex;
  int device = 0;
  int dim = 0, metric = 0, m = 0;
  int precision = KDBGPU_PRECISION_F32;
  int kind = kdb::KIND_L2_F32;
  uint32_t capacity = 0;
  uint32_t stride = 0;     // 32-bit words per row slot / prepared query (multiple of 128)
  uint32_t row_words = 0;  // 32-bit words between stored rows (= stride for float32)
  float abs_max = 0.f;     // Quantizer.AbsMax (int8)
  uint32_t n = 0;
  uint32_t entry = 0;
  int max_level = -1;
  bool has_graph = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int num_sms = 0;
  SearchTuning tuning;
  // searches hold `mu` shared (kNumSearchWs of them can be in flight, each on its own workspace and
  // stream, so consecutive batches overlap on the device); everything that changes the mirror or
  // uses the handle-level workspace holds it exclusively
  std::shared_mutex mu;
  struct SearchWs {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;  // completion of the last launch that used this workspace
    cudaStream_t launch_stream = nullptr;  // ... and the stream it was queued on (guarded by ws_mu)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf<float> q_raw, q_prep, qnorms;
    DevBuf<uint32_t> out_ids, out_counts, allow, visited, work_counter, redo;
    DevBuf<double> out_scores;
    DevBuf<HeapEntry> cand_overflow;
    DevBuf<unsigned long long> stats;
    DevBuf<int> err_flag;
    DevBuf<unsigned char> out_blob;  // host path: scores | ids | counts | stats | err, one D2H copy
    unsigned char *h_out = nullptr;  // pinned staging for that copy
    size_t h_out_bytes = 0;
    int grid = 0;
    uint32_t vis_words = 0;
    bool busy = false;
    void release() {
      out_blob.release();
      if (h_out) cudaFreeHost(h_out);
      h_out = nullptr;
      h_out_bytes = 0;
      q_raw.release(); q_prep.release(); qnorms.release(); out_ids.release(); out_counts.release(); allow.release();
      visited.release(); work_counter.release(); redo.release(); out_scores.release(); cand_overflow.release();
      stats.release(); err_flag.release();
    }
  };
  // one flat-scan call in flight: its own stream, staging and result buffers (two of them, so that the copies and the
  // host-side work of one call overlap the kernels of the next)
  struct FlatWs {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool busy = false;
    DevBuf<float> q_raw, q_prep, tq_sumsq, tq_resid2, tc_beta, t_gmin, t_theta, t_bound, t_thf;
    DevBuf<uint16_t> tq_bf16;
    DevBuf<uint32_t> out_ids, out_counts, allow, t_cnt, t_subcnt, t_flags, t_fcnt, t_fid;
    DevBuf<double> out_scores, flat_dist;
    DevBuf<uint2> t_sub, t_ovf;
    DevBuf<unsigned long long> t_nres;
    unsigned char *h_out = nullptr;  // pinned staging of one chunk's results
    size_t h_out_bytes = 0;
    void release() {
      if (h_out) cudaFreeHost(h_out);
      h_out = nullptr;
      h_out_bytes = 0;
      q_raw.release(); q_prep.release(); tq_sumsq.release(); tq_resid2.release(); tc_beta.release(); t_gmin.release();
      t_theta.release(); t_bound.release(); t_thf.release(); tq_bf16.release(); out_ids.release(); out_counts.release();
      allow.release(); t_cnt.release(); t_subcnt.release(); t_flags.release(); t_fcnt.release(); t_fid.release();
      out_scores.release(); flat_dist.release(); t_sub.release(); t_ovf.release(); t_nres.release();
    }
    size_t bytes() const {
      return q_raw.bytes() + q_prep.bytes() + tq_bf16.bytes() + tc_beta.bytes() + t_gmin.bytes() + t_sub.bytes() +
             t_ovf.bytes() + t_fid.bytes() + out_ids.bytes() + out_scores.bytes() + flat_dist.bytes() + allow.bytes();
    }
  };
  static constexpr int kNumFlatWs = 2;
  FlatWs fws[kNumFlatWs];
  std::mutex fws_mu;
  std::condition_variable fws_cv;
  std::mutex tc_mu;  // the lazy build of the bf16 mirror
  static constexpr int kNumSearchWs = 4;
  SearchWs sws[kNumSearchWs];
  std::mutex ws_mu;
  std::condition_variable ws_cv;
  unsigned ws_next = 0;
  int last_ws = -1;

  DevBuf<float> vecs, norms, conv_tmp, qnorm1;
  DevBuf<uint32_t> adj0, upper_adj, upper_first, deleted;
  DevBuf<int8_t> levels;
  bool has_deleted = false;
  // per-call workspace
  DevBuf<float> q_raw, q_prep;
  DevBuf<uint32_t> out_ids, out_counts, allow, visited, ids_tmp;
  DevBuf<double> out_scores, flat_dist, dist_tmp;
  DevBuf<HeapEntry> cand_overflow;
  DevBuf<unsigned long long> stats;
  DevBuf<uint32_t> work_counter;
  DevBuf<int> err_flag;
  int ws_grid = 0;
  uint32_t vis_words = 0;
  uint32_t ovf_cap = 1u << 15;
  // construction state: host mirrors of levels / upper-row ownership, device build workspace
  std::vector<int8_t> h_levels;
  std::vector<uint32_t> h_upper_first;
  uint32_t upper_rows_used = 0;
  DevBuf<uint32_t> upper_node;
  DevBuf<uint8_t> upper_level;
  DevBuf<uint32_t> b_out_off, b_slot_node, b_cand_ids, b_cand_cnt, b_row_cnt, b_row_off, b_srcs, b_active,
      b_scalars, b_scratch_ids;
  DevBuf<uint8_t> b_slot_level;
  DevBuf<double> b_scratch_d;
  // tensor-core flat pre-filter: bf16 mirror of the rows (built lazily, dropped when rows change)
  DevBuf<uint16_t> x_bf16;
  DevBuf<float> x_sumsq, x_resid2, x_max;
  bool tc_valid = false;
  uint32_t tc_n = 0;
  uint64_t arena_chunks_registered = 0;  // kdbgpu_arena_stage_chunk calls that page-locked the caller's mapping in place

  DevIndex dev() const {
    DevIndex d;
    d.vecs = vecs.p;
    d.norms = norms.p;
    d.row_words = row_words;
    d.kind = kind;
    d.adj0 = adj0.p;
    d.upper_adj = upper_adj.p;
    d.upper_first = upper_first.p;
    d.levels = levels.p;
    d.deleted = has_deleted ? deleted.p : nullptr;
    d.stride = stride;
    d.dim = (uint32_t)dim;
    d.n = n;
    d.deg0 = (uint32_t)(2 * m);
    d.degu = (uint32_t)m;
    d.entry = entry;
    d.max_level = max_level;
    d.metric = metric;
    return d;
  }
};

namespace kdb {

// Makes the handle's device current for the calling thread.  It is deliberately NOT restored on
// return: restoring a device the thread never used would create a primary context there (CUDA 12
// cudaSetDevice semantics) — costly in multi-process, one-rank-per-GPU deployments.
struct DeviceGuard {
  bool ok = true;
  explicit DeviceGuard(int dev) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) {
      ok = false;
      (void)cudaGetLastError();
      return;
    }
    if (cur != dev && cudaSetDevice(dev) != cudaSuccess) {
      ok = false;
      (void)cudaGetLastError();
    }
  }
};

// ---- helpers defined in api.cu, used by the shard group (shard.cu) ---------------------------------
// blocks until one of the handle's launch workspaces is free / hands it back
int acquire_ws(kdbgpu_index *h);
void release_ws(kdbgpu_index *h, int j);
// query preparation of searchInternal (hnsw_index.go:401-434) into w.q_prep (+ w.qnorms for int8)
int prepare_queries(kdbgpu_index *h, kdbgpu_index::SearchWs &w, const float *d_q_raw, uint32_t nq, cudaStream_t s);
// one traversal launch on `stream` with workspace `w`; ids are written as id + id_base (0 slots stay 0)
int enqueue_search(kdbgpu_index *h, kdbgpu_index::SearchWs &w, const float *d_q_prepared, uint32_t nq, int k, int ef,
                   const uint32_t *d_allow, uint32_t allow_entry, uint32_t *d_ids, double *d_scores,
                   uint32_t *d_counts, cudaStream_t stream, unsigned long long *d_stats = nullptr,
                   int *d_err = nullptr, uint32_t id_base = 0);
uint32_t first_set_bit(const uint64_t *bits, size_t words, bool *found);
int stage_allow(kdbgpu_index *h, DevBuf<uint32_t> &dst, const uint64_t *allow, size_t allow_words,
                cudaStream_t stream);
// flat scan (exhaustive or tensor-core pre-filter) on flat workspace `w` (acquire_fws / release_fws); `queries` and the
// outputs may be host or device pointers.  Caller holds h->mu shared.  Synchronises w.stream before returning.
int acquire_fws(kdbgpu_index *h);
void release_fws(kdbgpu_index *h, int i);
int flat_search_ws(kdbgpu_index *h, kdbgpu_index::FlatWs &w, const float *queries, uint32_t nq, int k, int mode,
                   bool prefilter, const uint32_t *d_allow, uint32_t *out_ids, double *out_scores, uint32_t *out_counts,
                   kdbgpu_stats *stats);

}  // namespace kdb
