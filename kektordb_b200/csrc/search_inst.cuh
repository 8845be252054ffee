// search_inst.cuh — the traversal kernels and their per-shape launchers, instantiated for ONE distance
// kind per translation unit (search_k0.cu .. search_k3.cu) so that the four sets compile in parallel.
// search.cu routes a call to the unit of the index's kind through search_kind_dispatch.
#pragma once
#include "searcher.cuh"

namespace kdb {

using namespace dev;

// op of search_kind_dispatch
enum : int { SEARCH_OP_LAUNCH = 0, SEARCH_OP_OCCUPANCY = 1, SEARCH_OP_LAUNCH_FAST = 2, SEARCH_OP_OCCUPANCY_FAST = 3 };

// launches (or sizes) the kernel of (slots, cpl) for this unit's kind; *occ receives resident CTAs per SM
typedef cudaError_t (*search_kind_fn)(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid,
                                      size_t smem, cudaStream_t stream, int *occ);
cudaError_t search_dispatch_k0(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid, size_t smem,
                               cudaStream_t stream, int *occ);
cudaError_t search_dispatch_k1(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid, size_t smem,
                               cudaStream_t stream, int *occ);
cudaError_t search_dispatch_k2(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid, size_t smem,
                               cudaStream_t stream, int *occ);
cudaError_t search_dispatch_k3(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid, size_t smem,
                               cudaStream_t stream, int *occ);

#ifdef KDB_SEARCH_KIND
namespace {

// One warp per CTA, persistent over the batch: queries are claimed from a global counter.
template <int SLOTS, int METRIC, int CPL>
__global__ void __launch_bounds__(32) hnsw_search_kernel(const DevIndex ix, const SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  Searcher<SLOTS, METRIC, CPL> s(ix, a, smem);
  s.init_barriers();
  // all of 0 .. nq, or the queries the fast kernel handed over (distance ties)
  const uint32_t limit = a.query_count ? *a.query_count : a.nq;
  for (;;) {
    uint32_t i = 0;
    if (s.lane == 0) i = atomicAdd(a.work_counter, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= limit) break;
    s.run_query(a.query_list ? a.query_list[i] : i);
  }
  if (s.lane == 0) {
    atomicAdd(&a.stats[0], s.st_e);
    atomicAdd(&a.stats[1], s.st_h);
    atomicAdd(&a.stats[2], s.st_h0);
  }
}

// The fast path of the same search (searcher.cuh, "Fast path"): sorted list in registers, whole-warp
// maintenance; a query that meets two equal distances is appended to redo_list and answered by
// hnsw_search_kernel afterwards, so the output is always the reference's.
template <int SLOTS, int METRIC, int CPL>
// (int8 rows up to 1536-d: the register cap follows the number of query-warps shared memory admits per SM)
__global__ void __launch_bounds__(32, (METRIC == KIND_COS_I8 && CPL >= 1 && CPL <= 3) ? (SLOTS <= 8 ? 24 : 16) : 1)
    hnsw_search_fast_kernel(const DevIndex ix, const SearchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  // two shapes: without heap arrays in shared memory (more resident query-warps; ties are handed to the
  // heap kernel through redo_list — for rows whose distances practically never tie), or with them, the
  // tied query being re-run by the heap path at once in this warp (a second launch would add a whole
  // query's latency to the batch)
  // (int8 rows always hand over — search_fast_hands_over — so the heap path is not even compiled into their kernel)
  const bool hand_over = METRIC == KIND_COS_I8 ? true : a.redo_list != nullptr;
  Searcher<SLOTS, METRIC, CPL> s(ix, a, smem, !hand_over);
  s.init_barriers();
  unsigned int n_tied = 0;
  for (;;) {
    uint32_t q = 0;
    if (s.lane == 0) q = atomicAdd(a.work_counter, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= a.nq) break;
    if (!s.run_query_fast(q)) {
      if (hand_over) {
        if (s.lane == 0) a.redo_list[atomicAdd(a.redo_count, 1u)] = q;
      } else {
        s.run_query(q);
        n_tied++;
      }
    }
    __syncwarp();
  }
  if (s.lane == 0 && n_tied) atomicAdd(a.redo_count, n_tied);
  if (s.lane == 0) {
    atomicAdd(&a.stats[0], s.st_e);
    atomicAdd(&a.stats[1], s.st_h);
    atomicAdd(&a.stats[2], s.st_h0);
  }
}


template <int SL, int KIND, int CPL>
cudaError_t search_shape(int op, const DevIndex &ix, const SearchArgs *a, int grid, size_t smem, cudaStream_t stream,
                         int *occ) {
  if (op == SEARCH_OP_LAUNCH || op == SEARCH_OP_OCCUPANCY) {
    auto kern = hnsw_search_kernel<SL, KIND, CPL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (op == SEARCH_OP_OCCUPANCY) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, 32, smem);
    kern<<<grid, 32, smem, stream>>>(ix, *a);
    return cudaGetLastError();
  }
  auto kern = hnsw_search_fast_kernel<SL, KIND, CPL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (op == SEARCH_OP_OCCUPANCY_FAST) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, 32, smem);
  kern<<<grid, 32, smem, stream>>>(ix, *a);
  return cudaGetLastError();
}

// CPL (16-byte columns per lane) instantiated at compile time; other row lengths use the generic path (0)
template <int SL, int KIND>
cudaError_t search_cpl(int op, const DevIndex &ix, const SearchArgs *a, int cpl, int grid, size_t smem,
                       cudaStream_t stream, int *occ) {
  switch (cpl) {
    case 1: return search_shape<SL, KIND, 1>(op, ix, a, grid, smem, stream, occ);
    case 2: return search_shape<SL, KIND, 2>(op, ix, a, grid, smem, stream, occ);
    case 3: return search_shape<SL, KIND, 3>(op, ix, a, grid, smem, stream, occ);
    case 4: return search_shape<SL, KIND, 4>(op, ix, a, grid, smem, stream, occ);
    case 6: return search_shape<SL, KIND, 6>(op, ix, a, grid, smem, stream, occ);
    case 8: return search_shape<SL, KIND, 8>(op, ix, a, grid, smem, stream, occ);
    case 12: return search_shape<SL, KIND, 12>(op, ix, a, grid, smem, stream, occ);
    default: return search_shape<SL, KIND, 0>(op, ix, a, grid, smem, stream, occ);
  }
}

template <int KIND>
cudaError_t search_kind_dispatch(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid,
                                 size_t smem, cudaStream_t stream, int *occ) {
  switch (slots) {
    case 2: return search_cpl<2, KIND>(op, ix, a, cpl, grid, smem, stream, occ);
    case 4: return search_cpl<4, KIND>(op, ix, a, cpl, grid, smem, stream, occ);
    case 8: return search_cpl<8, KIND>(op, ix, a, cpl, grid, smem, stream, occ);
    case 16: return search_cpl<16, KIND>(op, ix, a, cpl, grid, smem, stream, occ);
    default: return cudaErrorInvalidConfiguration;
  }
}

}  // namespace
#endif  // KDB_SEARCH_KIND

}  // namespace kdb
