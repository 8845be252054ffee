// kdb_internal.cuh — shared device/host internals of libkektordb_gpu (sm_100a only).
//
// Layout of the GPU mirror of one hnsw.Index (reference: pkg/core/hnsw/hnsw_node.go:13-39,
// hnsw_index.go:77) — see DESIGN.md §3:
//   vecs       [(cap+1)][stride] f32, stride = dim rounded up to 128 floats (one float4 column per lane
//              and pass; 512-byte-aligned rows, zero padded)
//   adj0       [(cap+1)][deg0]   u32, deg0 = 2M, zero padded (id 0 is the nil slot)
//   upper_adj  [rows][degu]      u32, degu = M; node i, level l>=1 -> row upper_first[i] + l-1
//   levels     [(cap+1)]         i8,  -1 = nil node
//   deleted    bitset over ids (u32 words)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kektordb_gpu.h"

namespace kdb {

struct DevIndex {
  const float *vecs;
  const uint32_t *adj0;
  const uint32_t *upper_adj;
  const uint32_t *upper_first;
  const int8_t *levels;
  const uint32_t *deleted;  // may be nullptr
  uint32_t stride;          // floats per row
  uint32_t dim;
  uint32_t n;  // highest id
  uint32_t deg0, degu;
  uint32_t entry;
  int max_level;
  int metric;
};

// One heap element, value semantics like types.Candidate{Id uint32; Distance float64}
// (pkg/core/types/types.go:18-21).
struct __align__(16) HeapEntry {
  double d;
  uint32_t id;
  uint32_t pad;
};

struct SearchArgs {
  const float *queries;  // prepared [nq][stride]
  uint32_t nq;
  int k, ef;
  const uint32_t *allow;  // nullptr = nil allow-list
  uint32_t allow_entry;   // smallest member of the allow-list
  uint32_t *out_ids;
  double *out_scores;
  uint32_t *out_counts;
  uint32_t *visited;  // [grid][vis_words]
  uint32_t vis_words; // multiple of 4
  HeapEntry *cand_overflow;  // [grid][ovf_cap]
  uint32_t ovf_cap;
  uint32_t cand_smem;  // candidate-heap entries held in shared memory
  unsigned long long *stats;  // [3] E, H, H0
  uint32_t *work_counter;
  int *err_flag;
};

struct SearchTuning {
  int slots = 4;            // row slots per query-warp (bulk copies in flight per query), power of two
  int cand_smem = 192;      // candidate-heap entries held in shared memory (the rest spills to HBM)
  int max_ctas_per_sm = 0;  // 0 = whatever fits
};
bool search_slots_supported(int slots);

size_t search_smem_bytes(const DevIndex &ix, int ef, const SearchTuning &t);
// returns resident CTAs per SM (0 = configuration does not fit)
int search_occupancy(const DevIndex &ix, int ef, const SearchTuning &t);
cudaError_t launch_search(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                          cudaStream_t stream);
cudaError_t launch_prep_queries(const float *in, size_t in_stride, float *out, uint32_t nq, uint32_t dim,
                                uint32_t stride, int metric, cudaStream_t stream);
cudaError_t launch_distance_batch(const DevIndex &ix, const float *query_prepared, const uint32_t *ids,
                                  uint32_t n, double *out, cudaStream_t stream);
cudaError_t launch_merge_topk(int n_shards, uint32_t nq, int k, const uint32_t *ids, const double *scores,
                              const uint32_t *counts, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, cudaStream_t stream);
// flat scan: dist [nq][n_rows] f64 (row r = id r+1), then exact top-k per query
cudaError_t launch_flat_distances(const DevIndex &ix, const float *queries_raw, const float *queries_prepared,
                                  uint32_t nq, int mode, double *dist, cudaStream_t stream);
cudaError_t launch_flat_select(const DevIndex &ix, const double *dist, uint32_t nq, int k,
                               const uint32_t *allow, uint32_t *out_ids, double *out_scores,
                               uint32_t *out_counts, cudaStream_t stream);

// flat scan, tensor-core pre-filter (flat_tc.cu): S[q][x] = alpha * <q~, x~> + beta[x] over bf16
// copies, used only to nominate candidates that are then re-scored exactly under a certificate
struct FlatTcLaunch {
  const void *q_bf16;  // [nq_pad][dp]
  const void *x_bf16;  // [n_pad][dp]
  uint32_t nq, nq_pad, n, n_pad, dp;
  float alpha;
  const float *beta;  // [n_pad], +inf = row excluded
  int epi;            // 0 store every score (validation), 1 group minima, 2 emit ids below theta
  float *S;
  size_t ldS;
  float *gmin;         // [nq_pad][visited tiles * 8]
  uint32_t ct_stride;  // visit every ct_stride-th corpus tile (0/1 = all); pass A may sample
  const float *theta;  // [nq_pad]
  void *sub;          // [nq_pad][grid][sub slots] {id, score}: nominees per (query, CTA), no atomics
  uint32_t *sub_cnt;  // [nq_pad][grid]
  uint32_t *ovf_cnt;  // [nq_pad] spill beyond the sub slots
  void *ovf;          // [nq_pad][cap]
  uint32_t cap;
  int grid;
};
uint32_t flat_tc_bm();
uint32_t flat_tc_bn();
uint32_t flat_tc_bk();
uint32_t flat_tc_groups_per_tile();
cudaError_t launch_to_bf16(const float *src, size_t src_stride, uint32_t rows, uint32_t dim, void *dst, uint32_t dp,
                           uint32_t rows_pad, float *sumsq, float *resid2, cudaStream_t stream);
cudaError_t launch_flat_tc(const FlatTcLaunch &L, cudaStream_t stream);
cudaError_t launch_tc_beta(const DevIndex &ix, const float *sumsq, const uint32_t *allow, int use_norm, uint32_t n_pad,
                           float *beta, cudaStream_t stream);
cudaError_t launch_tc_max(const float *sumsq, const float *resid2, uint32_t n, float *out2, cudaStream_t stream);
cudaError_t launch_tc_threshold(const float *gmin, uint32_t n_groups, uint32_t nq, int k, const float *qsumsq,
                                const float *qresid2, const float *xmax, float alpha, int use_norm, uint32_t dp,
                                float *theta, float *bound, cudaStream_t stream);
uint32_t flat_tc_sub_slots();
uint32_t flat_tc_max_queries();
cudaError_t launch_tc_refine(const void *sub, const uint32_t *sub_cnt, uint32_t grid, const uint32_t *ovf_cnt,
                             const void *ovf, uint32_t cap, uint32_t nq, int k, const float *theta, const float *bound,
                             uint32_t *fcnt, uint32_t *fid, uint32_t fcap, float *theta_final, uint32_t *flags,
                             cudaStream_t stream);
size_t tc_rescore_smem(uint32_t dim, uint32_t fcap);
cudaError_t launch_tc_rescore(const DevIndex &ix, int mode, const float *queries, size_t q_stride, uint32_t nq, int k,
                              const uint32_t *fcnt, const uint32_t *fid, uint32_t fcap, const float *theta_final,
                              const float *bound, const float *qsumsq, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, uint32_t *flags, unsigned long long *n_rescored,
                              cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// The distance arithmetic ("kernel order"; oracle KDBO_ARITH_KERNEL restates it):
// lane l owns the float4 columns c = l, l+32, ...; component j of column c feeds accumulator
// (4l + j) by FMA in increasing c; lane partial (a0+a1)+(a2+a3); xor-butterfly 16,8,4,2,1.
// Replaces dotProductAsDistanceGonum / squaredEuclideanDistanceGo
// (pkg/core/distance/distance_go.go:122-128, :57-68) and the Rust kernels behind
// native/compute/include/kektordb_compute.h:8-9.
template <int METRIC>
__device__ __forceinline__ float warp_reduce_row(const float4 *__restrict__ q4, const float4 *__restrict__ r4,
                                                 uint32_t nchunks, int lane) {
  float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
#pragma unroll 4
  for (uint32_t c = lane; c < nchunks; c += 32) {
    const float4 a = q4[c];
    const float4 b = r4[c];
    if (METRIC == KDBGPU_METRIC_COSINE) {
      ax = __fmaf_rn(a.x, b.x, ax);
      ay = __fmaf_rn(a.y, b.y, ay);
      az = __fmaf_rn(a.z, b.z, az);
      aw = __fmaf_rn(a.w, b.w, aw);
    } else {
      const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z),
                  dw = __fsub_rn(a.w, b.w);
      ax = __fmaf_rn(dx, dx, ax);
      ay = __fmaf_rn(dy, dy, ay);
      az = __fmaf_rn(dz, dz, az);
      aw = __fmaf_rn(dw, dw, aw);
    }
  }
  float s = __fadd_rn(__fadd_rn(ax, ay), __fadd_rn(az, aw));
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}
// DistanceFuncF32 result: float64(sum) for L2 (distance_go.go:67), 1.0 - float64(dot) for cosine (:127)
template <int METRIC>
__device__ __forceinline__ double to_distance(float s) {
  return METRIC == KDBGPU_METRIC_COSINE ? 1.0 - static_cast<double>(s) : static_cast<double>(s);
}
#endif  // __CUDACC__

}  // namespace kdb
