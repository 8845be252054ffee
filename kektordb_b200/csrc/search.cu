// search.cu — HNSW traversal on sm_100a: one CTA per query, persistent over the batch.
//
// Stands in for (*Index).searchInternal / searchLayerUnlocked
// (reference pkg/core/hnsw/hnsw_index.go:369-468, :2351-2611) with the heaps of
// hnsw_heap.go:18-156 and the visited BitSet of bitset.go.  Per hop the CTA
//   A. (lane 0)   pops the nearest candidate, tests the early exit (:2497-2506)
//   B. (warp 0)   reads the adjacency row, test-and-sets the visited bitset, applies the
//                 allow-list BEFORE any distance (:2537-2549), compacts the survivors in row order
//   C. (all warps) gathers the survivors' rows HBM -> shared memory with 1-D bulk copies
//                 (cp.async.bulk + mbarrier, per-warp slots) and reduces query x row in the
//                 fixed "kernel order" (kdb_internal.cuh)
//   D. (lane 0)   applies the result/candidate heap updates sequentially in row order (:2571-2591)
// Distances of one hop are computed in parallel but applied in the reference's order, and the two
// binary heaps are the reference's own algorithms, so ids, order and scores are bit-identical to
// the oracle in KDBO_ARITH_KERNEL mode — ties included.
#include "searcher.cuh"

namespace kdb {

using namespace dev;

namespace {

template <int NWARPS, int SLOTS, int METRIC>
__global__ void __launch_bounds__(NWARPS * 32) hnsw_search_kernel(const DevIndex ix, const SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  Searcher<NWARPS, SLOTS, METRIC> s(ix, a, smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NWARPS * SLOTS; ++i) mbar_init(&s.sm.bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) s.sm.ctl->q = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t q = s.sm.ctl->q;
    if (q >= a.nq) break;
    s.run_query(q);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
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

template <int NW, int SL>
cudaError_t launch_cfg(const DevIndex &ix, const SearchArgs &a, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e;
  if (ix.metric == KDBGPU_METRIC_COSINE) {
    auto kern = hnsw_search_kernel<NW, SL, KDBGPU_METRIC_COSINE>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, stream>>>(ix, a);
  } else {
    auto kern = hnsw_search_kernel<NW, SL, KDBGPU_METRIC_L2>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, stream>>>(ix, a);
  }
  return cudaGetLastError();
}

template <int NW, int SL>
int occupancy_cfg(const DevIndex &ix, size_t smem) {
  int nb = 0;
  cudaError_t e;
  if (ix.metric == KDBGPU_METRIC_COSINE) {
    auto kern = hnsw_search_kernel<NW, SL, KDBGPU_METRIC_COSINE>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NW * 32, smem);
  } else {
    auto kern = hnsw_search_kernel<NW, SL, KDBGPU_METRIC_L2>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NW * 32, smem);
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return nb;
}

#define KDB_DISPATCH_CFG(NWv, SLv, EXPR)           \
  if (t.nwarps == NWv && t.slots == SLv) {         \
    constexpr int NW = NWv;                        \
    constexpr int SL = SLv;                        \
    EXPR;                                          \
  }

}  // namespace

size_t search_smem_bytes(const DevIndex &ix, int ef, const SearchTuning &t) {
  return smem_layout(ix.stride, ef, t.nwarps, t.slots, (uint32_t)t.cand_smem, nullptr, nullptr);
}

int search_occupancy(const DevIndex &ix, int ef, const SearchTuning &t) {
  const size_t smem = search_smem_bytes(ix, ef, t);
  if (smem > 227 * 1024) return 0;
  int nb = -1;
  KDB_DISPATCH_CFG(2, 2, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  KDB_DISPATCH_CFG(2, 4, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  KDB_DISPATCH_CFG(4, 1, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  KDB_DISPATCH_CFG(4, 2, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  KDB_DISPATCH_CFG(4, 4, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  KDB_DISPATCH_CFG(8, 1, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  KDB_DISPATCH_CFG(8, 2, nb = (occupancy_cfg<NW, SL>(ix, smem)))
  if (nb < 0) return 0;
  if (t.max_ctas_per_sm > 0 && nb > t.max_ctas_per_sm) nb = t.max_ctas_per_sm;
  return nb;
}

cudaError_t launch_search(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                          cudaStream_t stream) {
  const size_t smem = search_smem_bytes(ix, a.ef, t);
  cudaError_t e = cudaErrorInvalidConfiguration;
  KDB_DISPATCH_CFG(2, 2, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
  KDB_DISPATCH_CFG(2, 4, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
  KDB_DISPATCH_CFG(4, 1, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
  KDB_DISPATCH_CFG(4, 2, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
  KDB_DISPATCH_CFG(4, 4, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
  KDB_DISPATCH_CFG(8, 1, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
  KDB_DISPATCH_CFG(8, 2, e = (launch_cfg<NW, SL>(ix, a, grid, smem, stream)))
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
