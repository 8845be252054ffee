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
// float16 / int8 indexes (distance.Float16 / distance.Int8, reference distance_go.go:41-46) keep the
// same arrays; `vecs` then holds the rows in their stored form (float16 bits / int8) with a pitch of
// row_words 32-bit words (row bytes rounded up to 128), and int8 indexes add norms[(cap+1)] f32
// (quantizedNorms, hnsw_index.go:87).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kektordb_gpu.h"

namespace kdb {

// distance kind = (metric, precision) pair the reference offers (float32Funcs / float16Funcs /
// int8Funcs, distance_go.go:133-146).  The first two values equal KDBGPU_METRIC_*.
enum : int { KIND_L2_F32 = 0, KIND_COS_F32 = 1, KIND_L2_F16 = 2, KIND_COS_I8 = 3 };

struct DevIndex {
  const float *vecs;        // rows as 32-bit words (f32 values, or packed float16 bits / int8)
  const float *norms;       // int8 only: computeInt8Norm of every stored row (hnsw_index.go:3371)
  uint32_t row_words;       // 32-bit words between consecutive rows (= stride for float32)
  int kind;                 // KIND_*
  const uint32_t *adj0;
  const uint32_t *upper_adj;
  const uint32_t *upper_first;
  const int8_t *levels;
  const uint32_t *deleted;  // may be nullptr
  uint32_t stride;          // 32-bit words per shared-memory row slot / prepared query (multiple of 128)
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
  const float *queries;  // prepared [nq][stride] (32-bit words: f32, or packed float16 / int8)
  const float *qnorms;   // int8 only: query-side norm per query (hnsw_index.go:2405-2413)
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
  // fast kernel: queries that met a distance tie are appended here for the exact kernel, which then
  // runs over query_list[0 .. *query_count) instead of 0 .. nq
  uint32_t *redo_list;
  uint32_t *redo_count;
  const uint32_t *query_list;
  const uint32_t *query_count;
  // id-range shards: returned ids are local id + id_base, so that the per-shard results can be
  // gathered and merged without another pass (empty slots stay 0)
  uint32_t id_base;
  int rows_evict_first;  // row copies carry an L2 evict_first hint (SearchTuning.rows_evict_first)
};

struct SearchTuning {
  int fast = 1;             // sorted-list fast pass with exact re-run on ties (searcher.cuh): 0 = heaps only,
                            // 1 = where ties are practically absent (int8), 2 = every kind
  int slots = 4;            // row slots per query-warp (two groups of slots / 2 rows in flight per query): 2, 4, 8, 16
  int slots_idle = 0;       // > 0: shape of a launch that finds no other batch of the handle in flight (the latency
                            // shape: more rows in flight per query, fewer resident queries); 0 = always `slots`
  int cand_smem = 192;      // candidate-heap entries held in shared memory (the rest spills to HBM)
  int max_ctas_per_sm = 0;  // 0 = whatever fits
  int rows_evict_first = 0; // row bulk copies carry an L2 evict_first hint
};
bool search_slots_supported(int slots);

size_t search_smem_bytes(const DevIndex &ix, int ef, const SearchTuning &t);
// returns resident CTAs per SM (0 = configuration does not fit)
int search_occupancy(const DevIndex &ix, int ef, const SearchTuning &t);
cudaError_t launch_search(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                          cudaStream_t stream);
// fast path (ef <= 128, no soft-deleted nodes): shape, occupancy, launch
bool search_fast_eligible(const DevIndex &ix, int ef, const SearchTuning &t);
int search_fast_occupancy(const DevIndex &ix, int ef, const SearchTuning &t);
bool search_fast_hands_over(const DevIndex &ix);  // ties go to a second (heap) launch instead of being re-run in place
cudaError_t launch_search_fast(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                               cudaStream_t stream);
cudaError_t launch_prep_queries(const float *in, size_t in_stride, float *out, uint32_t nq, uint32_t dim,
                                uint32_t stride, int metric, cudaStream_t stream);
cudaError_t launch_distance_batch(const DevIndex &ix, const float *query_prepared, const float *qnorm,
                                  const uint32_t *ids, uint32_t n, double *out, cudaStream_t stream);
// float16 / int8 indexes: f32 rows -> stored form.  normalise = the cosine query preparation of
// searchInternal (hnsw_index.go:406-414); conversion = float16.Fromfloat32 (:427-430) or
// Quantizer.Quantize (quantizer.go:135-160).  out rows have a pitch of out_words 32-bit words and are
// zero padded; norms (int8, may be nullptr) receives computeInt8Norm per row, with 0 mapped to 1 when
// query_side is set (:2410-2413).
cudaError_t launch_convert_rows(const float *in, size_t in_stride, float *out, size_t out_words, uint32_t rows,
                                uint32_t dim, int kind, bool normalise, float abs_max, float *norms,
                                bool query_side, cudaStream_t stream);
// adjacency rows rewritten in place (incremental mirror refresh)
cudaError_t launch_patch_rows(uint32_t *base, uint32_t deg, const uint32_t *dst_row, const uint32_t *src_off,
                              const uint32_t *src_cnt, const uint32_t *src, uint32_t n_patches, cudaStream_t stream);
// one radix-select pass of Quantizer.Train's quantile (see abs_hist_kernel)
cudaError_t launch_abs_hist(const float *rows, size_t row_stride, uint32_t n_sample, uint32_t step, uint32_t dim,
                            uint32_t prefix, int hi_shift, int shift, int bits, unsigned long long *hist,
                            cudaStream_t stream);
// arena chunk payload (device) -> mirror rows of the ids whose physical slot lives in that chunk
// CSR slice of whole nodes -> fixed-degree adjacency rows of the mirror (search.cu: graph_scatter_kernel)
cudaError_t launch_graph_scatter(const uint64_t *node_row, const uint64_t *row_off, const uint32_t *nbrs,
                                 uint32_t first_node, uint32_t n_nodes, uint64_t row0, uint64_t edge0, uint64_t edge1,
                                 uint32_t n, const int8_t *levels, const uint32_t *upper_first, uint32_t deg0,
                                 uint32_t degu, uint32_t *adj0, uint32_t *upper, int *err, cudaStream_t stream);
cudaError_t launch_arena_scatter(const unsigned char *chunk, uint32_t chunk_id, uint32_t vecs_per_chunk,
                                 uint32_t vector_bytes, const uint32_t *slot_table, uint32_t first_id, uint32_t last_id,
                                 float *vecs, size_t row_words, unsigned int *n_staged, cudaStream_t stream);
// int8 rows already in stored form -> norms[row] = computeInt8Norm
cudaError_t launch_int8_norms(const float *rows, size_t row_words, uint32_t count, uint32_t dim, float *norms,
                              cudaStream_t stream);
cudaError_t launch_merge_topk(int n_shards, uint32_t nq, int k, const uint32_t *ids, const double *scores,
                              const uint32_t *counts, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, cudaStream_t stream);
// One batch's results as ONE device blob, so that a single copy (D2H) or a single collective (all-gather of
// per-shard results) moves everything: scores [nq][k] f64 | ids [nq][k] u32 | counts [nq] u32 | stats [4] u64 |
// err i32 (+ pad).  `bytes` is a multiple of 16.
struct BlobLayout {
  size_t o_scores, o_ids, o_counts, o_stats, o_err, bytes;
};
inline BlobLayout blob_layout(uint32_t nq, int k) {
  BlobLayout L;
  const size_t nk = (size_t)nq * (size_t)k;
  L.o_scores = 0;
  L.o_ids = nk * sizeof(double);
  L.o_counts = L.o_ids + nk * sizeof(uint32_t);
  L.o_stats = (L.o_counts + (size_t)nq * sizeof(uint32_t) + 7) & ~(size_t)7;
  L.o_err = L.o_stats + 4 * sizeof(unsigned long long);
  L.bytes = (L.o_err + sizeof(long long) + 15) & ~(size_t)15;
  return L;
}
// merge of per-shard blobs `gather` [n_shards] x shard_stride bytes (each laid out as L) into `out` (same
// layout): per query the k best by (distance, id); stats summed, err = first non-zero.  ids are global already.
cudaError_t launch_merge_packed(int n_shards, uint32_t nq, int k, const unsigned char *gather, size_t shard_stride,
                                const BlobLayout &L, unsigned char *out, cudaStream_t stream);
// local ids -> global ids of an id-range shard (empty slots stay 0); the traversal does this in its epilogue
// (SearchArgs.id_base), the flat scan's results go through this kernel
cudaError_t launch_add_id_base(uint32_t *ids, size_t n, uint32_t base, cudaStream_t stream);
// local bit i of dst (words [0, n_words)) = global bit (base + i) of src (src_words 32-bit words; bits beyond are 0);
// local bit 0 (the nil id) is cleared.  Slices a global allow-list for the shard owning ids base+1 ..
cudaError_t launch_slice_bits(const uint32_t *src, size_t src_words, uint32_t base, uint32_t *dst, size_t n_words,
                              cudaStream_t stream);
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
  uint32_t *tile_counter;  // pair kernel: one zeroed word per launch = dynamic tile schedule; null = static
};
uint32_t flat_tc_bm();
uint32_t flat_tc_bn();
uint32_t flat_tc_bk();
uint32_t flat_tc_groups_per_tile();
cudaError_t launch_to_bf16(const float *src, size_t src_stride, uint32_t rows, uint32_t dim, void *dst, uint32_t dp,
                           uint32_t rows_pad, float *sumsq, float *resid2, cudaStream_t stream);
cudaError_t launch_flat_tc(FlatTcLaunch &L, cudaStream_t stream);  // L.grid <- CTAs actually launched
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
// raw shared-address forms (the traversal keeps its slot / barrier bases as 32-bit shared addresses,
// so issuing a row costs a handful of instructions on lane 0's serial path)
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 eviction-priority hint for data that is read once (the corpus rows of a traversal: 99 % of the traffic, no reuse),
// so that what IS re-read — adjacency rows, visited words, levels, norms — survives in L2 longer
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_u32_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// (a suspend-time hint on try_wait was measured and rejected: the coarser wake-up cost 3-5 % of throughput)
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// true on exactly one lane of the (converged) warp.  A bulk copy issued under this predicate compiles to ONE
// UBLKCP; under `if (lane == 0)` ptxas cannot see that a single lane is active and wraps every copy in an
// elect-one-lane-at-a-time loop (ELECT / UBLKCP / PLOP3 / BRA.U.ANY).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred)
      :
      : "memory");
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// The distance arithmetic ("kernel order"; oracle KDBO_ARITH_KERNEL restates it).
// A lane owns the 16-byte columns c = l, l+32, ... of the row.
//   float32: column = 4 elements; component j of column c feeds accumulator (4l + j) by FMA in
//            increasing c; lane partial (a0+a1)+(a2+a3).
//   float16: column = 8 elements (4 words x {lo, hi}); element i of the column feeds accumulator
//            (8l + i); lane partial ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)).  Differences are taken in
//            f32 on the widened halves, as squaredEuclideanGoFloat16 does (distance_go.go:93-105).
//   int8:    column = 16 elements; exact int32 dot (dotProductGoInt8, distance_go.go:108-118) — any
//            order gives the same bits; the partial travels bit-cast in a float.
// Then the xor-butterfly 16,8,4,2,1 (integer add for int8).
// Replaces dotProductAsDistanceGonum / squaredEuclideanDistanceGo / squaredEuclideanGoFloat16 /
// dotProductGoInt8 and the Rust kernels behind native/compute/include/kektordb_compute.h:8-11.
// Packed float32 pairs (sm_100 FFMA2 / FADD2): one instruction, two independent IEEE operations — the same
// bits as two scalar __fmaf_rn / __fsub_rn, half the issue slots.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 sub2_rn(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
template <int KIND>
struct LaneAcc {
  f32x2 axy = 0ull, azw = 0ull;  // accumulators (4l, 4l+1), (4l+2, 4l+3)
  __device__ __forceinline__ void add(const float4 &a, const float4 &b) {
    if (KIND == KIND_COS_F32) {
      axy = fma2_rn(pack2(a.x, a.y), pack2(b.x, b.y), axy);
      azw = fma2_rn(pack2(a.z, a.w), pack2(b.z, b.w), azw);
    } else {
      const f32x2 dxy = sub2_rn(pack2(a.x, a.y), pack2(b.x, b.y)), dzw = sub2_rn(pack2(a.z, a.w), pack2(b.z, b.w));
      axy = fma2_rn(dxy, dxy, axy);
      azw = fma2_rn(dzw, dzw, azw);
    }
  }
  __device__ __forceinline__ float lane_sum() const {
    float ax, ay, az, aw;
    unpack2(axy, ax, ay);
    unpack2(azw, az, aw);
    return __fadd_rn(__fadd_rn(ax, ay), __fadd_rn(az, aw));
  }
};
__device__ __forceinline__ float2 f16x2_widen(float word) {
  const uint32_t u = __float_as_uint(word);
  const __half2 h = *reinterpret_cast<const __half2 *>(&u);
  return __half22float2(h);  // .x = low half = even element
}
// a 16-byte column of 8 packed halves widened to float32: lo = elements 0..3, hi = elements 4..7
__device__ __forceinline__ void f16x8_widen(const float4 &c, float4 &lo, float4 &hi) {
  const float2 w0 = f16x2_widen(c.x), w1 = f16x2_widen(c.y), w2 = f16x2_widen(c.z), w3 = f16x2_widen(c.w);
  lo = make_float4(w0.x, w0.y, w1.x, w1.y);
  hi = make_float4(w2.x, w2.y, w3.x, w3.y);
}
template <>
struct LaneAcc<KIND_L2_F16> {
  f32x2 a[4] = {0ull, 0ull, 0ull, 0ull};  // a[w] = accumulators (8l + 2w, 8l + 2w + 1)
  __device__ __forceinline__ void add2(const float2 &q, float xb, int w) {
    const float2 x = f16x2_widen(xb);
    const f32x2 d = sub2_rn(pack2(q.x, q.y), pack2(x.x, x.y));
    a[w] = fma2_rn(d, d, a[w]);
  }
  __device__ __forceinline__ void add(const float4 &q, const float4 &b) {  // q: 8 packed halves
    add2(f16x2_widen(q.x), b.x, 0);
    add2(f16x2_widen(q.y), b.y, 1);
    add2(f16x2_widen(q.z), b.z, 2);
    add2(f16x2_widen(q.w), b.w, 3);
  }
  // the query column already widened (f16x8_widen): the traversal keeps it in registers in this form
  __device__ __forceinline__ void add_wide(const float4 &qlo, const float4 &qhi, const float4 &b) {
    add2(make_float2(qlo.x, qlo.y), b.x, 0);
    add2(make_float2(qlo.z, qlo.w), b.y, 1);
    add2(make_float2(qhi.x, qhi.y), b.z, 2);
    add2(make_float2(qhi.z, qhi.w), b.w, 3);
  }
  __device__ __forceinline__ float lane_sum() const {
    float s[8];
#pragma unroll
    for (int w = 0; w < 4; ++w) unpack2(a[w], s[2 * w], s[2 * w + 1]);
    return __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3])),
                     __fadd_rn(__fadd_rn(s[4], s[5]), __fadd_rn(s[6], s[7])));
  }
};
template <>
struct LaneAcc<KIND_COS_I8> {
  int a0 = 0, a1 = 0;
  __device__ __forceinline__ void add(const float4 &q, const float4 &b) {
    a0 = __dp4a(__float_as_int(q.x), __float_as_int(b.x), a0);
    a1 = __dp4a(__float_as_int(q.y), __float_as_int(b.y), a1);
    a0 = __dp4a(__float_as_int(q.z), __float_as_int(b.z), a0);
    a1 = __dp4a(__float_as_int(q.w), __float_as_int(b.w), a1);
  }
  __device__ __forceinline__ float lane_sum() const { return __int_as_float(a0 + a1); }
};
template <int KIND>
__device__ __forceinline__ float warp_sum(float s) {
  if (KIND == KIND_COS_I8) return __int_as_float(__reduce_add_sync(0xffffffffu, __float_as_int(s)));
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}
template <int KIND>
__device__ __forceinline__ float warp_reduce_row(const float4 *__restrict__ q4, const float4 *__restrict__ r4,
                                                 uint32_t nchunks, int lane) {
  LaneAcc<KIND> acc;
#pragma unroll 4
  for (uint32_t c = lane; c < nchunks; c += 32) acc.add(q4[c], r4[c]);
  return warp_sum<KIND>(acc.lane_sum());
}
// int8 cosine distance from the exact dot and the two norms (searchLayerUnlocked distFn,
// hnsw_index.go:2421-2449; distanceBetweenNodes :319-336): float64 divide, clamp, 1 - similarity
__device__ __forceinline__ double int8_distance(int dot, float qnorm, float stored_norm) {
  if (stored_norm == 0.f) return 1.0;
  double sim = __ddiv_rn(static_cast<double>(dot), __dmul_rn(static_cast<double>(qnorm), static_cast<double>(stored_norm)));
  if (sim > 1.0) sim = 1.0;
  if (sim < -1.0) sim = -1.0;
  return __dsub_rn(1.0, sim);
}
// DistanceFunc result from the reduced value `s`: float64(sum) for L2 (distance_go.go:67, :104),
// 1.0 - float64(dot) for cosine f32 (:127), the norm-scaled form for int8
template <int KIND>
__device__ __forceinline__ double to_distance(float s, float qnorm = 0.f, float stored_norm = 0.f) {
  if (KIND == KIND_COS_I8) return int8_distance(__float_as_int(s), qnorm, stored_norm);
  return KIND == KIND_COS_F32 ? 1.0 - static_cast<double>(s) : static_cast<double>(s);
}
#endif  // __CUDACC__

}  // namespace kdb
