// search.cu — the query-side kernels on sm_100a: the persistent HNSW traversal (one warp per
// query, see searcher.cuh), exact query preparation, and the batched distance hook.
//
// hnsw_search_kernel stands in for (*Index).searchInternal / searchLayerUnlocked (reference
// pkg/core/hnsw/hnsw_index.go:369-468, :2351-2611).
#include "searcher.cuh"

namespace kdb {

using namespace dev;

namespace {

// One warp per CTA, persistent over the batch: queries are claimed from a global counter.
template <int SLOTS, int METRIC, int CPL>
__global__ void __launch_bounds__(32) hnsw_search_kernel(const DevIndex ix, const SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  Searcher<SLOTS, METRIC, CPL> s(ix, a, smem);
  s.init_barriers();
  for (;;) {
    uint32_t q = 0;
    if (s.lane == 0) q = atomicAdd(a.work_counter, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= a.nq) break;
    s.run_query(q);
  }
  if (s.lane == 0) {
    atomicAdd(&a.stats[0], s.st_e);
    atomicAdd(&a.stats[1], s.st_h);
    atomicAdd(&a.stats[2], s.st_h0);
  }
}

// normalize() of hnsw_index.go:3030-3045, bit-exact: sequential f32 sum of squares without FMA,
// one f64 sqrt, f32 reciprocal, f32 scale; zero vectors untouched.  One warp per query; also pads
// the row to `stride` with zeros.  L2 queries are only copied and padded (:412-414).
__global__ void prep_queries_kernel(const float *__restrict__ in, size_t in_stride, float *__restrict__ out,
                                    uint32_t nq, uint32_t dim, uint32_t stride, int metric) {
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const float *src = in + (size_t)q * in_stride;
  float *dst = out + (size_t)q * stride;
  float inv = 1.0f;
  bool scale = false;
  if (metric == KDBGPU_METRIC_COSINE) {
    float norm_sq = 0.f;
    if (lane == 0) {
#pragma unroll 16
      for (uint32_t i = 0; i < dim; ++i) {
        const float v = __ldg(src + i);
        norm_sq = __fadd_rn(norm_sq, __fmul_rn(v, v));
      }
    }
    norm_sq = __shfl_sync(0xffffffffu, norm_sq, 0);
    if (norm_sq > 0.f) {
      inv = __fdiv_rn(1.0f, static_cast<float>(sqrt(static_cast<double>(norm_sq))));
      scale = true;
    }
  }
  for (uint32_t i = lane; i < stride; i += 32) {
    float v = i < dim ? src[i] : 0.f;
    if (scale) v = __fmul_rn(v, inv);
    dst[i] = v;
  }
}

// The literal inner-loop hook: out[i] = distFn(query, ids[i]) (hnsw_index.go:2393-2396) — one warp
// per candidate row, 128-bit loads straight from HBM, same kernel-order reduction.
template <int METRIC>
__global__ void distance_batch_kernel(const DevIndex ix, const float *__restrict__ query,
                                      const uint32_t *__restrict__ ids, uint32_t n, double *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4 *q4 = reinterpret_cast<float4 *>(smem);
  const uint32_t nchunks = ix.stride >> 2;
  for (uint32_t c = threadIdx.x; c < nchunks; c += blockDim.x) q4[c] = reinterpret_cast<const float4 *>(query)[c];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t wpb = blockDim.x >> 5;
  for (uint32_t i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const uint32_t id = ids[i];
    if (id == 0 || id > ix.n) {
      if (lane == 0) out[i] = __longlong_as_double(0x7ff8000000000000LL);  // nil node: NaN
      continue;
    }
    const float s = warp_reduce_row<METRIC>(q4, reinterpret_cast<const float4 *>(ix.vecs + (size_t)id * ix.stride),
                                            nchunks, lane);
    if (lane == 0) out[i] = to_distance<METRIC>(s);
  }
}

// CPL (float4 columns per lane) instantiated at compile time; other row lengths use the generic path
__host__ inline int cpl_of(const DevIndex &ix) {
  const uint32_t c = ix.stride / 128;
  return (c == 1 || c == 2 || c == 3 || c == 4 || c == 6 || c == 8 || c == 12) ? (int)c : 0;
}

template <int SL, int METRIC, int CPL>
cudaError_t launch_one(const DevIndex &ix, const SearchArgs &a, int grid, size_t smem, cudaStream_t stream) {
  auto kern = hnsw_search_kernel<SL, METRIC, CPL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 32, smem, stream>>>(ix, a);
  return cudaGetLastError();
}
template <int SL, int METRIC, int CPL>
int occupancy_one(size_t smem) {
  auto kern = hnsw_search_kernel<SL, METRIC, CPL>;
  int nb = 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return nb;
}

#define KDB_CASE_CPL(SLv, METRICv, CPLv, EXPR) \
  case CPLv: {                                 \
    constexpr int SL = SLv;                    \
    constexpr int MT = METRICv;                \
    constexpr int CP = CPLv;                   \
    EXPR;                                      \
  } break;
#define KDB_SWITCH_CPL(SLv, METRICv, EXPR) \
  switch (cpl) {                           \
    KDB_CASE_CPL(SLv, METRICv, 1, EXPR)    \
    KDB_CASE_CPL(SLv, METRICv, 2, EXPR)    \
    KDB_CASE_CPL(SLv, METRICv, 3, EXPR)    \
    KDB_CASE_CPL(SLv, METRICv, 4, EXPR)    \
    KDB_CASE_CPL(SLv, METRICv, 6, EXPR)    \
    KDB_CASE_CPL(SLv, METRICv, 8, EXPR)    \
    KDB_CASE_CPL(SLv, METRICv, 12, EXPR)   \
    default: {                             \
      constexpr int SL = SLv;              \
      constexpr int MT = METRICv;          \
      constexpr int CP = 0;                \
      EXPR;                                \
    } break;                               \
  }
#define KDB_SWITCH_METRIC(SLv, EXPR)                   \
  if (ix.metric == KDBGPU_METRIC_COSINE) {             \
    KDB_SWITCH_CPL(SLv, KDBGPU_METRIC_COSINE, EXPR)    \
  } else {                                             \
    KDB_SWITCH_CPL(SLv, KDBGPU_METRIC_L2, EXPR)        \
  }
#define KDB_DISPATCH(EXPR)                       \
  switch (t.slots) {                             \
    case 2: { KDB_SWITCH_METRIC(2, EXPR) } break;   \
    case 4: { KDB_SWITCH_METRIC(4, EXPR) } break;   \
    case 8: { KDB_SWITCH_METRIC(8, EXPR) } break;   \
    case 16: { KDB_SWITCH_METRIC(16, EXPR) } break; \
    default: break;                              \
  }

}  // namespace

bool search_slots_supported(int slots) { return slots == 2 || slots == 4 || slots == 8 || slots == 16; }

size_t search_smem_bytes(const DevIndex &ix, int ef, const SearchTuning &t) {
  return smem_layout(ix.stride, ef, t.slots, (uint32_t)t.cand_smem, ix.deg0 > ix.degu ? ix.deg0 : ix.degu,
                     cpl_of(ix) == 0, nullptr, nullptr);
}

int search_occupancy(const DevIndex &ix, int ef, const SearchTuning &t) {
  const size_t smem = search_smem_bytes(ix, ef, t);
  if (smem > 227 * 1024) return 0;
  const int cpl = cpl_of(ix);
  int nb = -1;
  KDB_DISPATCH(nb = (occupancy_one<SL, MT, CP>(smem)))
  if (nb < 0) return 0;
  if (t.max_ctas_per_sm > 0 && nb > t.max_ctas_per_sm) nb = t.max_ctas_per_sm;
  return nb;
}

cudaError_t launch_search(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                          cudaStream_t stream) {
  const size_t smem = search_smem_bytes(ix, a.ef, t);
  const int cpl = cpl_of(ix);
  cudaError_t e = cudaErrorInvalidConfiguration;
  KDB_DISPATCH(e = (launch_one<SL, MT, CP>(ix, a, grid, smem, stream)))
  return e;
}

cudaError_t launch_prep_queries(const float *in, size_t in_stride, float *out, uint32_t nq, uint32_t dim,
                                uint32_t stride, int metric, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  const int wpb = 4;
  prep_queries_kernel<<<(nq + wpb - 1) / wpb, wpb * 32, 0, stream>>>(in, in_stride, out, nq, dim, stride, metric);
  return cudaGetLastError();
}

cudaError_t launch_distance_batch(const DevIndex &ix, const float *query_prepared, const uint32_t *ids,
                                  uint32_t n, double *out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int threads = 256;
  const uint32_t wpb = threads / 32;
  int grid = (int)((n + wpb - 1) / wpb);
  if (grid > 148 * 8) grid = 148 * 8;
  const size_t smem = (size_t)ix.stride * sizeof(float);
  if (ix.metric == KDBGPU_METRIC_COSINE)
    distance_batch_kernel<KDBGPU_METRIC_COSINE><<<grid, threads, smem, stream>>>(ix, query_prepared, ids, n, out);
  else
    distance_batch_kernel<KDBGPU_METRIC_L2><<<grid, threads, smem, stream>>>(ix, query_prepared, ids, n, out);
  return cudaGetLastError();
}

}  // namespace kdb
