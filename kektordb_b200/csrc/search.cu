// search.cu — the query-side kernels on sm_100a: the persistent HNSW traversal (one warp per
// query, see searcher.cuh), exact query preparation, and the batched distance hook.
//
// hnsw_search_kernel stands in for (*Index).searchInternal / searchLayerUnlocked (reference
// pkg/core/hnsw/hnsw_index.go:369-468, :2351-2611).
#include "search_inst.cuh"

namespace kdb {

namespace {

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

// The literal inner-loop hook: out[i] = distFn(query, ids[i]) (hnsw_index.go:2388-2454) — one warp
// per candidate row, 128-bit loads straight from HBM, same kernel-order reduction.
template <int KIND>
__global__ void distance_batch_kernel(const DevIndex ix, const float *__restrict__ query, const float *__restrict__ qnorm,
                                      const uint32_t *__restrict__ ids, uint32_t n, double *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4 *q4 = reinterpret_cast<float4 *>(smem);
  const uint32_t nchunks = ix.row_words >> 2;
  for (uint32_t c = threadIdx.x; c < nchunks; c += blockDim.x) q4[c] = reinterpret_cast<const float4 *>(query)[c];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t wpb = blockDim.x >> 5;
  const float qn = KIND == KIND_COS_I8 ? qnorm[0] : 0.f;
  for (uint32_t i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const uint32_t id = ids[i];
    if (id == 0 || id > ix.n) {
      if (lane == 0) out[i] = __longlong_as_double(0x7ff8000000000000LL);  // nil node: NaN
      continue;
    }
    const float s = warp_reduce_row<KIND>(q4, reinterpret_cast<const float4 *>(ix.vecs + (size_t)id * ix.row_words),
                                          nchunks, lane);
    if (lane == 0) out[i] = to_distance<KIND>(s, qn, KIND == KIND_COS_I8 ? ix.norms[id] : 0.f);
  }
}

// f32 rows -> stored form of a float16 / int8 index (one warp per row), optionally after the cosine
// query normalisation of searchInternal (hnsw_index.go:406-414, same arithmetic as
// prep_queries_kernel).  float16: float16.Fromfloat32 = IEEE round-to-nearest-even (:427-430,
// :501-505).  int8: Quantizer.Quantize (quantizer.go:135-160): f32 divide, f32 multiply by 127,
// clip, round half away from zero in float64; norm = f32(sqrt(f64(sum of squares))) (:3371-3377).
__global__ void convert_rows_kernel(const float *__restrict__ in, size_t in_stride, float *__restrict__ out,
                                    size_t out_words, uint32_t rows, uint32_t dim, int kind, int normalise,
                                    float abs_max, float *__restrict__ norms, int query_side) {
  const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float *src = in + (size_t)r * in_stride;
  uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)r * out_words);
  float inv = 1.0f;
  bool scale = false;
  if (normalise) {
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
  int sumsq = 0;  // <= 127^2 * 8192 fits int32
  for (uint32_t w = lane; w < (uint32_t)out_words; w += 32) {
    uint32_t word = 0u;
    if (kind == KIND_L2_F16) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t e = 2 * w + j;
        float v = e < dim ? src[e] : 0.f;
        if (scale) v = __fmul_rn(v, inv);
        word |= (uint32_t)__half_as_ushort(__float2half_rn(v)) << (16 * j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t e = 4 * w + j;
        int qv = 0;
        if (e < dim && abs_max != 0.f) {
          float v = src[e];
          if (scale) v = __fmul_rn(v, inv);
          float scaled = __fmul_rn(__fdiv_rn(v, abs_max), 127.0f);
          if (scaled > 127.0f)
            scaled = 127.0f;
          else if (scaled < -127.0f)
            scaled = -127.0f;
          qv = static_cast<int>(round(static_cast<double>(scaled)));
        }
        sumsq += qv * qv;
        word |= ((uint32_t)qv & 0xffu) << (8 * j);
      }
    }
    dst[w] = word;
  }
  if (kind == KIND_COS_I8 && norms != nullptr) {
    sumsq = __reduce_add_sync(0xffffffffu, sumsq);
    if (lane == 0) {
      float nrm = static_cast<float>(sqrt(static_cast<double>(sumsq)));
      if (query_side && nrm == 0.f) nrm = 1.f;
      norms[r] = nrm;
    }
  }
}

// Quantizer.Train (quantizer.go:49-125) needs the value at a fixed rank of |x| over a stride sample
// of the rows; the sort + index of the reference becomes an exact 3-pass radix select over the f32 bit
// patterns (non-negative floats order like their bits).  One pass: histogram of `bits` bits at `shift`
// over the sampled values whose higher bits equal `prefix`.
__global__ void abs_hist_kernel(const float *__restrict__ rows, size_t row_stride, uint32_t n_sample, uint32_t step,
                                uint32_t dim, uint32_t prefix, int hi_shift, int shift, int bits,
                                unsigned long long *__restrict__ hist) {
  const uint64_t total = (uint64_t)n_sample * dim;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / dim;
    const uint32_t c = (uint32_t)(i - r * dim);
    const uint32_t key = __float_as_uint(fabsf(rows[r * step * row_stride + c]));
    if (hi_shift >= 32 || (key >> hi_shift) == (prefix >> hi_shift))
      atomicAdd(&hist[(key >> shift) & ((1u << bits) - 1u)], 1ull);
  }
}

// Arena chunk -> logical rows (VectorArena.GetBytes, pkg/storage/mmap/arena.go:378-444): physical slot
// p of logical id i lives in chunk p / vecs_per_chunk at payload offset (p % vecs_per_chunk) *
// vector_bytes.  `chunk` is the chunk's payload (header stripped) already in device memory; one warp
// moves one row into the mirror (pitch row_words, padding untouched = zero).
__global__ void arena_scatter_kernel(const unsigned char *__restrict__ chunk, uint32_t chunk_id, uint32_t vecs_per_chunk,
                                     uint32_t vector_bytes, const uint32_t *__restrict__ slot_table, uint32_t first_id,
                                     uint32_t last_id, float *__restrict__ vecs, size_t row_words,
                                     unsigned int *__restrict__ n_staged) {
  const uint32_t id = first_id + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (id > last_id) return;
  const uint32_t p = slot_table ? slot_table[id] : id - 1u;  // sequential Adds: AllocSlot hands out 0, 1, 2, ... (:121-152)
  if (p == 0xffffffffu || p / vecs_per_chunk != chunk_id) return;
  const unsigned char *src = chunk + (size_t)(p % vecs_per_chunk) * vector_bytes;
  unsigned char *dst = reinterpret_cast<unsigned char *>(vecs + (size_t)id * row_words);
  if ((vector_bytes & 3u) == 0u && (reinterpret_cast<uintptr_t>(src) & 3u) == 0u) {
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t w = lane; w < (vector_bytes >> 2); w += 32) d4[w] = s4[w];
  } else {
    for (uint32_t b = lane; b < vector_bytes; b += 32) dst[b] = src[b];
  }
  if (lane == 0) atomicAdd(n_staged, 1u);
}

// Topology staging (kdbgpu_set_graph / kdbgpu_set_graph_file): the adjacency arrives as the reference-shaped CSR
// (node i owns rows node_row[i] .. node_row[i+1]-1, level 0 first; row r holds nbrs[row_off[r] .. row_off[r+1])) in
// slices of whole nodes [first_node, first_node + n_nodes).  One warp per node writes every row into the fixed-degree
// mirror (level 0: adj0[id][2M], upper levels: upper[upper_first[id] + l - 1][M]), dropping nil / out-of-range
// neighbours in place with an order-preserving warp compaction — the reference skips them at search time with no
// effect on results (hnsw_index.go:2553-2561).  err[0..2] = {code, node, level}: 1 = more live neighbours than the row
// holds, 2 = row offsets outside the slice / not monotone.
__global__ void graph_scatter_kernel(const uint64_t *__restrict__ node_row, const uint64_t *__restrict__ row_off,
                                     const uint32_t *__restrict__ nbrs, uint32_t first_node, uint32_t n_nodes,
                                     uint64_t row0, uint64_t edge0, uint64_t edge1, uint32_t n,
                                     const int8_t *__restrict__ levels, const uint32_t *__restrict__ upper_first,
                                     uint32_t deg0, uint32_t degu, uint32_t *__restrict__ adj0,
                                     uint32_t *__restrict__ upper, int *__restrict__ err) {
  const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n_nodes) return;
  const uint32_t id = first_node + w;
  const int L = levels[id];
  if (L < 0) return;
  const uint64_t r_first = node_row[w] - row0;
  for (int l = 0; l <= L; ++l) {
    const uint64_t eb = row_off[r_first + (uint64_t)l], ee = row_off[r_first + (uint64_t)l + 1];
    if (eb < edge0 || ee > edge1 || eb > ee) {
      if (lane == 0 && atomicCAS(&err[0], 0, 2) == 0) {
        err[1] = (int)id;
        err[2] = l;
      }
      return;
    }
    const uint32_t cap = l == 0 ? deg0 : degu;
    uint32_t *dst = l == 0 ? adj0 + (size_t)id * deg0 : upper + ((size_t)upper_first[id] + (size_t)(l - 1)) * degu;
    uint32_t kept = 0;
    for (uint64_t i0 = eb; i0 < ee; i0 += 32) {
      const uint64_t i = i0 + (uint64_t)lane;
      const uint32_t nb = i < ee ? nbrs[i - edge0] : 0u;
      const bool keep = nb != 0u && nb <= n && levels[nb] >= 0;
      const uint32_t mask = __ballot_sync(0xffffffffu, keep);
      const uint32_t pos = kept + (uint32_t)__popc(mask & ((1u << lane) - 1u));
      if (keep && pos < cap) dst[pos] = nb;
      kept += (uint32_t)__popc(mask);
    }
    if (kept > cap) {
      if (lane == 0 && atomicCAS(&err[0], 0, 1) == 0) {
        err[1] = (int)id;
        err[2] = l;
      }
      return;
    }
  }
}

// Incremental topology refresh: adjacency row dst_row[i] of `base` (deg slots) := src[src_off[i] ..
// +src_cnt[i]), zero padded — the rewritten node.Connections[level] of Add's forward / reverse links
// (hnsw_index.go:717-783), Vacuum's reconnectNode and Refine's commits (optimizer.go).
__global__ void patch_rows_kernel(uint32_t *__restrict__ base, uint32_t deg, const uint32_t *__restrict__ dst_row,
                                  const uint32_t *__restrict__ src_off, const uint32_t *__restrict__ src_cnt,
                                  const uint32_t *__restrict__ src, uint32_t n_patches) {
  const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n_patches) return;
  uint32_t *row = base + (size_t)dst_row[p] * deg;
  const uint32_t off = src_off[p], cnt = src_cnt[p];
  for (uint32_t t = lane; t < deg; t += 32) row[t] = t < cnt ? src[off + t] : 0u;
}

// computeInt8Norm (hnsw_index.go:3371-3377) of rows already in stored form, one warp per row
__global__ void int8_norms_kernel(const float *__restrict__ rows, size_t row_words, uint32_t count, uint32_t dim,
                                  float *__restrict__ norms) {
  const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= count) return;
  const int *src = reinterpret_cast<const int *>(rows + (size_t)r * row_words);
  int sumsq = 0;
  for (uint32_t w = lane; w < (dim + 3) / 4; w += 32) sumsq = __dp4a(src[w], src[w], sumsq);  // padding is zero
  sumsq = __reduce_add_sync(0xffffffffu, sumsq);
  if (lane == 0) norms[r] = static_cast<float>(sqrt(static_cast<double>(sumsq)));
}

// CPL (float4 columns per lane) instantiated at compile time; other row lengths use the generic path
__host__ inline int cpl_of(const DevIndex &ix) {
  const uint32_t c = ix.stride / 128;
  return (c == 1 || c == 2 || c == 3 || c == 4 || c == 6 || c == 8 || c == 12) ? (int)c : 0;
}

}  // namespace

namespace {
search_kind_fn kind_fn(const DevIndex &ix) {
  switch (ix.kind) {
    case KIND_COS_F32: return search_dispatch_k1;
    case KIND_L2_F16: return search_dispatch_k2;
    case KIND_COS_I8: return search_dispatch_k3;
    default: return search_dispatch_k0;
  }
}
}  // namespace

bool search_slots_supported(int slots) { return slots == 2 || slots == 4 || slots == 8 || slots == 16; }

size_t search_smem_bytes(const DevIndex &ix, int ef, const SearchTuning &t) {
  return smem_layout(ix.stride, slot_pitch_words(ix.stride, ix.row_words, ix.kind), ef, t.slots, (uint32_t)t.cand_smem, ix.deg0 > ix.degu ? ix.deg0 : ix.degu,
                     cpl_of(ix) == 0, ix.kind, nullptr, nullptr);
}

int search_occupancy(const DevIndex &ix, int ef, const SearchTuning &t) {
  const size_t smem = search_smem_bytes(ix, ef, t);
  if (smem > 227 * 1024) return 0;
  int nb = 0;
  if (kind_fn(ix)(SEARCH_OP_OCCUPANCY, ix, nullptr, t.slots, cpl_of(ix), 0, smem, nullptr, &nb) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (t.max_ctas_per_sm > 0 && nb > t.max_ctas_per_sm) nb = t.max_ctas_per_sm;
  return nb;
}

cudaError_t launch_search(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                          cudaStream_t stream) {
  const size_t smem = search_smem_bytes(ix, a.ef, t);
  return kind_fn(ix)(SEARCH_OP_LAUNCH, ix, &a, t.slots, cpl_of(ix), grid, smem, stream, nullptr);
}

// hand_over: no heap arrays (ties go to the heap kernel); otherwise the heap kernel's own carve-up
static size_t search_fast_smem_bytes(const DevIndex &ix, int ef, const SearchTuning &t, bool hand_over) {
  return smem_layout(ix.stride, slot_pitch_words(ix.stride, ix.row_words, ix.kind), hand_over ? 0 : ef, t.slots, hand_over ? 0u : (uint32_t)t.cand_smem,
                     ix.deg0 > ix.degu ? ix.deg0 : ix.degu, cpl_of(ix) == 0, ix.kind, nullptr, nullptr);
}
bool search_fast_hands_over(const DevIndex &ix) { return ix.kind == KIND_COS_I8; }

bool search_fast_eligible(const DevIndex &ix, int ef, const SearchTuning &t) {
  // soft-deleted nodes are traversed but never kept (hnsw_index.go:2584): they need the two separate queues.
  // fast = 1 (default) uses the pass where it measures faster: int8 rows (distances are full-precision
  // float64 ratios, no ties, +12 %).  On float32 / float16 rows (float32 sums: 2-3 % of the queries at
  // 1 M x 768 meet a relevant tie) it measures slower than the heaps; fast = 2 forces it.
  const bool kind_ok = t.fast >= 2 || (t.fast == 1 && ix.kind == KIND_COS_I8);
  return kind_ok && ef <= 128 && ix.deleted == nullptr &&
         search_fast_smem_bytes(ix, ef, t, search_fast_hands_over(ix)) <= 227 * 1024;
}

int search_fast_occupancy(const DevIndex &ix, int ef, const SearchTuning &t) {
  const size_t smem = search_fast_smem_bytes(ix, ef, t, search_fast_hands_over(ix));
  int nb = 0;
  if (kind_fn(ix)(SEARCH_OP_OCCUPANCY_FAST, ix, nullptr, t.slots, cpl_of(ix), 0, smem, nullptr, &nb) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (t.max_ctas_per_sm > 0 && nb > t.max_ctas_per_sm) nb = t.max_ctas_per_sm;
  return nb;
}

cudaError_t launch_search_fast(const DevIndex &ix, const SearchArgs &a, const SearchTuning &t, int grid,
                               cudaStream_t stream) {
  const size_t smem = search_fast_smem_bytes(ix, a.ef, t, a.redo_list != nullptr);
  return kind_fn(ix)(SEARCH_OP_LAUNCH_FAST, ix, &a, t.slots, cpl_of(ix), grid, smem, stream, nullptr);
}

cudaError_t launch_prep_queries(const float *in, size_t in_stride, float *out, uint32_t nq, uint32_t dim,
                                uint32_t stride, int metric, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  const int wpb = 4;
  prep_queries_kernel<<<(nq + wpb - 1) / wpb, wpb * 32, 0, stream>>>(in, in_stride, out, nq, dim, stride, metric);
  return cudaGetLastError();
}

cudaError_t launch_distance_batch(const DevIndex &ix, const float *query_prepared, const float *qnorm,
                                  const uint32_t *ids, uint32_t n, double *out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int threads = 256;
  const uint32_t wpb = threads / 32;
  int grid = (int)((n + wpb - 1) / wpb);
  if (grid > 148 * 8) grid = 148 * 8;
  const size_t smem = (size_t)ix.stride * sizeof(float);
  switch (ix.kind) {
    case KIND_COS_F32:
      distance_batch_kernel<KIND_COS_F32><<<grid, threads, smem, stream>>>(ix, query_prepared, qnorm, ids, n, out);
      break;
    case KIND_L2_F16:
      distance_batch_kernel<KIND_L2_F16><<<grid, threads, smem, stream>>>(ix, query_prepared, qnorm, ids, n, out);
      break;
    case KIND_COS_I8:
      distance_batch_kernel<KIND_COS_I8><<<grid, threads, smem, stream>>>(ix, query_prepared, qnorm, ids, n, out);
      break;
    default:
      distance_batch_kernel<KIND_L2_F32><<<grid, threads, smem, stream>>>(ix, query_prepared, qnorm, ids, n, out);
      break;
  }
  return cudaGetLastError();
}

cudaError_t launch_convert_rows(const float *in, size_t in_stride, float *out, size_t out_words, uint32_t rows,
                                uint32_t dim, int kind, bool normalise, float abs_max, float *norms,
                                bool query_side, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  const int wpb = 4;
  convert_rows_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, stream>>>(in, in_stride, out, out_words, rows, dim, kind,
                                                                       normalise ? 1 : 0, abs_max, norms,
                                                                       query_side ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_patch_rows(uint32_t *base, uint32_t deg, const uint32_t *dst_row, const uint32_t *src_off,
                              const uint32_t *src_cnt, const uint32_t *src, uint32_t n_patches, cudaStream_t stream) {
  if (n_patches == 0) return cudaSuccess;
  const uint32_t wpb = 8;
  patch_rows_kernel<<<(n_patches + wpb - 1) / wpb, wpb * 32, 0, stream>>>(base, deg, dst_row, src_off, src_cnt, src,
                                                                       n_patches);
  return cudaGetLastError();
}

cudaError_t launch_abs_hist(const float *rows, size_t row_stride, uint32_t n_sample, uint32_t step, uint32_t dim,
                            uint32_t prefix, int hi_shift, int shift, int bits, unsigned long long *hist,
                            cudaStream_t stream) {
  abs_hist_kernel<<<148 * 4, 256, 0, stream>>>(rows, row_stride, n_sample, step, dim, prefix, hi_shift, shift, bits, hist);
  return cudaGetLastError();
}

cudaError_t launch_arena_scatter(const unsigned char *chunk, uint32_t chunk_id, uint32_t vecs_per_chunk,
                                 uint32_t vector_bytes, const uint32_t *slot_table, uint32_t first_id, uint32_t last_id,
                                 float *vecs, size_t row_words, unsigned int *n_staged, cudaStream_t stream) {
  if (last_id < first_id) return cudaSuccess;
  const uint32_t count = last_id - first_id + 1, wpb = 8;
  arena_scatter_kernel<<<(count + wpb - 1) / wpb, wpb * 32, 0, stream>>>(chunk, chunk_id, vecs_per_chunk, vector_bytes,
                                                                      slot_table, first_id, last_id, vecs, row_words,
                                                                      n_staged);
  return cudaGetLastError();
}

cudaError_t launch_graph_scatter(const uint64_t *node_row, const uint64_t *row_off, const uint32_t *nbrs,
                                 uint32_t first_node, uint32_t n_nodes, uint64_t row0, uint64_t edge0, uint64_t edge1,
                                 uint32_t n, const int8_t *levels, const uint32_t *upper_first, uint32_t deg0,
                                 uint32_t degu, uint32_t *adj0, uint32_t *upper, int *err, cudaStream_t stream) {
  if (n_nodes == 0) return cudaSuccess;
  const uint32_t wpb = 8;
  graph_scatter_kernel<<<(n_nodes + wpb - 1) / wpb, wpb * 32, 0, stream>>>(node_row, row_off, nbrs, first_node, n_nodes, row0,
                                                                      edge0, edge1, n, levels, upper_first, deg0, degu, adj0,
                                                                      upper, err);
  return cudaGetLastError();
}

cudaError_t launch_int8_norms(const float *rows, size_t row_words, uint32_t count, uint32_t dim, float *norms,
                              cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const int wpb = 4;
  int8_norms_kernel<<<(count + wpb - 1) / wpb, wpb * 32, 0, stream>>>(rows, row_words, count, dim, norms);
  return cudaGetLastError();
}

}  // namespace kdb
