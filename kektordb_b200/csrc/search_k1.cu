// search_k1.cu — traversal kernels (heap pass + fast pass) instantiated for KIND_COS_F32; see search_inst.cuh.
#define KDB_SEARCH_KIND 1
#include "search_inst.cuh"

namespace kdb {

cudaError_t search_dispatch_k1(int op, const DevIndex &ix, const SearchArgs *a, int slots, int cpl, int grid, size_t smem,
                               cudaStream_t stream, int *occ) {
  return search_kind_dispatch<KIND_COS_F32>(op, ix, a, slots, cpl, grid, smem, stream, occ);
}

}  // namespace kdb
