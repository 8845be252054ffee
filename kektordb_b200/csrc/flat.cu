// flat.cu — exhaustive scan (BruteForceIndex.SearchWithScores, reference
// pkg/core/vector_index.go:104-162) and the k-way merge of per-shard results.
//
// Exact path: float64 distances in the reference's own arithmetic (sequential in the element
// index, separate multiply and add), materialised as dist[nq][n] in HBM, then an exact per-query
// top-k by radix select on the (distance, id) order.  It is both the product's flat path (mode 0)
// and the ground truth for recall (mode 1).
#include "kdb_internal.cuh"

namespace kdb {
namespace {

constexpr int FT_ROWS = 64;   // corpus rows per CTA tile
constexpr int FT_Q = 16;      // queries per CTA tile
constexpr int FT_E = 32;      // elements per staged chunk
constexpr int FT_THREADS = 256;

// mode 0: sum += f64(q_e - x_e)^2, difference taken in f32 (vector_index.go:150-162), raw query
// mode 1: cosine -> 1 - sum f64(q^_e)*f64(x_e);  L2 -> sum (f64(q_e) - f64(x_e))^2
template <int MODE, int METRIC>
__global__ void __launch_bounds__(FT_THREADS)
    flat_distances_kernel(const DevIndex ix, const float *__restrict__ queries, size_t q_stride, uint32_t nq,
                          double *__restrict__ dist) {
  __shared__ float s_rows[FT_ROWS][FT_E + 1];
  __shared__ float s_q[FT_Q][FT_E];
  const uint32_t row0 = blockIdx.x * FT_ROWS;  // row r <-> id r + 1
  const uint32_t q0 = blockIdx.y * FT_Q;
  const int t = threadIdx.x;
  const int qi = t >> 4;       // 0..15
  const int rg = t & 15;       // rows rg, rg+16, rg+32, rg+48
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (uint32_t e0 = 0; e0 < ix.dim; e0 += FT_E) {
    for (int i = t; i < FT_ROWS * FT_E; i += FT_THREADS) {
      const int r = i / FT_E, e = i % FT_E;
      const uint32_t row = row0 + r;
      float v = 0.f;
      if (row < ix.n && e0 + e < ix.dim) v = ix.vecs[(size_t)(row + 1) * ix.stride + e0 + e];
      s_rows[r][e] = v;
    }
    for (int i = t; i < FT_Q * FT_E; i += FT_THREADS) {
      const int qq = i / FT_E, e = i % FT_E;
      float v = 0.f;
      if (q0 + qq < nq && e0 + e < ix.dim) v = queries[(size_t)(q0 + qq) * q_stride + e0 + e];
      s_q[qq][e] = v;
    }
    __syncthreads();
    const int emax = (ix.dim - e0) < (uint32_t)FT_E ? (int)(ix.dim - e0) : FT_E;
    for (int e = 0; e < emax; ++e) {
      const float qf = s_q[qi][e];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xf = s_rows[rg + 16 * j][e];
        if (MODE == 0) {
          const double diff = static_cast<double>(__fsub_rn(qf, xf));
          acc[j] = __dadd_rn(acc[j], __dmul_rn(diff, diff));
        } else if (METRIC == KDBGPU_METRIC_COSINE) {
          acc[j] = __dadd_rn(acc[j], __dmul_rn(static_cast<double>(qf), static_cast<double>(xf)));
        } else {
          const double diff = __dsub_rn(static_cast<double>(qf), static_cast<double>(xf));
          acc[j] = __dadd_rn(acc[j], __dmul_rn(diff, diff));
        }
      }
    }
    __syncthreads();
  }
  if (q0 + qi < nq) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t row = row0 + rg + 16 * j;
      if (row < ix.n) {
        double d = acc[j];
        if (MODE == 1 && METRIC == KDBGPU_METRIC_COSINE) d = __dsub_rn(1.0, d);
        dist[(size_t)(q0 + qi) * ix.n + row] = d;
      }
    }
  }
}

// order-preserving map double -> uint64
__device__ __forceinline__ unsigned long long dkey(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}

constexpr int SEL_THREADS = 512;
constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr int SEL_MAXK = 1024;

// One CTA per query.  Exact k smallest by (distance, id) among live, allowed rows.
__global__ void __launch_bounds__(SEL_THREADS)
    flat_select_kernel(const DevIndex ix, const double *__restrict__ dist, uint32_t nq, int k,
                       const uint32_t *__restrict__ allow, uint32_t *__restrict__ out_ids,
                       double *__restrict__ out_scores, uint32_t *__restrict__ out_counts) {
  __shared__ uint32_t hist[SEL_BINS];
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_remaining;
  __shared__ unsigned long long s_keys[SEL_MAXK];
  __shared__ uint32_t s_ids[SEL_MAXK];
  __shared__ uint32_t s_count, s_valid;
  const uint32_t q = blockIdx.x;
  if (q >= nq) return;
  const double *row = dist + (size_t)q * ix.n;
  const int t = threadIdx.x;
  auto live = [&](uint32_t r) -> bool {
    const uint32_t id = r + 1;
    if (ix.levels[id] < 0) return false;
    if (ix.deleted && ((ix.deleted[id >> 5] >> (id & 31)) & 1u)) return false;
    if (allow && !((allow[id >> 5] >> (id & 31)) & 1u)) return false;
    return true;
  };
  // count live rows
  if (t == 0) s_valid = 0;
  __syncthreads();
  {
    uint32_t c = 0;
    for (uint32_t r = t; r < ix.n; r += SEL_THREADS) c += live(r) ? 1u : 0u;
    atomicAdd(&s_valid, c);
  }
  __syncthreads();
  const uint32_t kk = s_valid < (uint32_t)k ? s_valid : (uint32_t)k;
  if (kk == 0) {
    for (int i = t; i < k; i += SEL_THREADS) {
      out_ids[(size_t)q * k + i] = 0;
      out_scores[(size_t)q * k + i] = 0.0;
    }
    if (t == 0) out_counts[q] = 0;
    return;
  }
  // radix select of the kk-th smallest key, most significant digit first
  if (t == 0) {
    s_prefix = 0ULL;
    s_remaining = kk;
  }
  __syncthreads();
  int shift = 64;
  while (shift > 0) {
    const int bits = shift >= SEL_BITS ? SEL_BITS : shift;
    shift -= bits;
    for (int i = t; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const int hi_shift = shift + bits;  // bits above the current digit
    for (uint32_t r = t; r < ix.n; r += SEL_THREADS) {
      if (!live(r)) continue;
      const unsigned long long key = dkey(row[r]);
      const bool match = hi_shift >= 64 ? true : ((key >> hi_shift) == (prefix >> hi_shift));
      if (match) atomicAdd(&hist[(key >> shift) & ((1u << bits) - 1u)], 1u);
    }
    __syncthreads();
    if (t == 0) {
      uint32_t rem = s_remaining;
      uint32_t b = 0;
      for (; b < (1u << bits); ++b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      s_prefix = prefix | ((unsigned long long)b << shift);
      s_remaining = rem;
    }
    __syncthreads();
  }
  const unsigned long long kth = s_prefix;  // key of the kk-th smallest
  const uint32_t ties_needed = s_remaining;  // how many rows equal to kth are inside the top kk
  if (t == 0) {
    s_count = 0;
  }
  __syncthreads();
  // strictly smaller keys: all of them
  for (uint32_t r = t; r < ix.n; r += SEL_THREADS) {
    if (!live(r)) continue;
    const unsigned long long key = dkey(row[r]);
    if (key < kth) {
      const uint32_t pos = atomicAdd(&s_count, 1u);
      s_keys[pos] = key;
      s_ids[pos] = r + 1;
    }
  }
  __syncthreads();
  // ties at kth: smallest ids first — sequential scan in id order by chunks
  {
    __shared__ uint32_t s_tie_base;
    if (t == 0) s_tie_base = s_count;
    __syncthreads();
    for (uint32_t r0 = 0; r0 < ix.n; r0 += SEL_THREADS) {
      const uint32_t r = r0 + t;
      bool is_tie = false;
      if (r < ix.n && live(r)) is_tie = dkey(row[r]) == kth;
      // block-wide ordered compaction via warp ballots
      __shared__ uint32_t warp_cnt[SEL_THREADS / 32];
      const uint32_t bal = __ballot_sync(0xffffffffu, is_tie);
      if ((t & 31) == 0) warp_cnt[t >> 5] = __popc(bal);
      __syncthreads();
      uint32_t before = 0;
      for (int w = 0; w < (t >> 5); ++w) before += warp_cnt[w];
      uint32_t total = 0;
      for (int w = 0; w < SEL_THREADS / 32; ++w) total += warp_cnt[w];
      const uint32_t taken = s_tie_base - s_count;  // ties stored so far
      if (is_tie) {
        const uint32_t rank = taken + before + __popc(bal & ((1u << (t & 31)) - 1u));
        if (rank < ties_needed) {
          s_keys[s_count + rank] = kth;
          s_ids[s_count + rank] = r + 1;
        }
      }
      __syncthreads();
      if (t == 0) s_tie_base += total;
      __syncthreads();
      if (s_tie_base - s_count >= ties_needed) break;
    }
  }
  __syncthreads();
  // sort the kk collected by (key, id): bitonic over the next power of two
  uint32_t np2 = 1;
  while (np2 < kk) np2 <<= 1;
  for (uint32_t i = kk + t; i < np2; i += SEL_THREADS) {
    s_keys[i] = ~0ULL;
    s_ids[i] = 0xffffffffu;
  }
  __syncthreads();
  for (uint32_t size = 2; size <= np2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = t; i < np2; i += SEL_THREADS) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const unsigned long long ki = s_keys[i], kj = s_keys[j];
          const uint32_t ii = s_ids[i], ij = s_ids[j];
          const bool gt = ki > kj || (ki == kj && ii > ij);
          if (gt == up) {
            s_keys[i] = kj;
            s_keys[j] = ki;
            s_ids[i] = ij;
            s_ids[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = t; i < k; i += SEL_THREADS) {
    const bool ok = (uint32_t)i < kk;
    out_ids[(size_t)q * k + i] = ok ? s_ids[i] : 0u;
    out_scores[(size_t)q * k + i] = ok ? dkey_inv(s_keys[i]) : 0.0;
  }
  if (t == 0) out_counts[q] = kk;
}

// Merge of per-shard results: the k smallest of the union by (distance, id).  One CTA per query;
// every candidate computes its rank among all candidates (exact under ties, independent of the
// order inside a shard's list) and writes itself to out[rank] if rank < k.
__global__ void __launch_bounds__(128)
    merge_topk_kernel(int n_shards, uint32_t nq, int k, const uint32_t *__restrict__ ids,
                      const double *__restrict__ scores, const uint32_t *__restrict__ counts,
                      uint32_t *__restrict__ out_ids, double *__restrict__ out_scores,
                      uint32_t *__restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  const uint32_t q = blockIdx.x;
  if (q >= nq) return;
  const int total_cap = n_shards * k;
  double *s_d = reinterpret_cast<double *>(merge_smem);
  uint32_t *s_id = reinterpret_cast<uint32_t *>(s_d + total_cap);
  __shared__ uint32_t s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < total_cap; i += blockDim.x) {
    const int s = i / k, j = i % k;
    uint32_t c = counts[(size_t)s * nq + q];
    if (c > (uint32_t)k) c = (uint32_t)k;
    if ((uint32_t)j < c) {
      const size_t o = ((size_t)s * nq + q) * k + j;
      const uint32_t pos = atomicAdd(&s_n, 1u);
      s_d[pos] = scores[o];
      s_id[pos] = ids[o];
    }
  }
  __syncthreads();
  const uint32_t n = s_n;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = s_d[i];
    const uint32_t id = s_id[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) {
      const double dj = s_d[j];
      rank += (dj < d || (dj == d && s_id[j] < id)) ? 1u : 0u;
    }
    if (rank < (uint32_t)k) {
      out_ids[(size_t)q * k + rank] = id;
      out_scores[(size_t)q * k + rank] = d;
    }
  }
  const uint32_t m = n < (uint32_t)k ? n : (uint32_t)k;
  for (uint32_t i = m + threadIdx.x; i < (uint32_t)k; i += blockDim.x) {
    out_ids[(size_t)q * k + i] = 0;
    out_scores[(size_t)q * k + i] = 0.0;
  }
  if (threadIdx.x == 0) out_counts[q] = m;
}

// The same merge over per-shard result blobs (BlobLayout) as they arrive from the all-gather / the peer copies:
// one CTA per query; CTA 0 also folds the shards' counters and error flags.
__global__ void __launch_bounds__(128)
    merge_packed_kernel(int n_shards, uint32_t nq, int k, const unsigned char *__restrict__ gather, size_t shard_stride,
                        BlobLayout L, unsigned char *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  const uint32_t q = blockIdx.x;
  const int total_cap = n_shards * k;
  double *s_d = reinterpret_cast<double *>(merge_smem);
  uint32_t *s_id = reinterpret_cast<uint32_t *>(s_d + total_cap);
  __shared__ uint32_t s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < total_cap; i += blockDim.x) {
    const int s = i / k, j = i % k;
    const unsigned char *b = gather + (size_t)s * shard_stride;
    uint32_t c = reinterpret_cast<const uint32_t *>(b + L.o_counts)[q];
    if (c > (uint32_t)k) c = (uint32_t)k;
    if ((uint32_t)j < c) {
      const size_t o = (size_t)q * k + j;
      const uint32_t pos = atomicAdd(&s_n, 1u);
      s_d[pos] = reinterpret_cast<const double *>(b + L.o_scores)[o];
      s_id[pos] = reinterpret_cast<const uint32_t *>(b + L.o_ids)[o];
    }
  }
  __syncthreads();
  const uint32_t n = s_n;
  double *o_sc = reinterpret_cast<double *>(out + L.o_scores) + (size_t)q * k;
  uint32_t *o_id = reinterpret_cast<uint32_t *>(out + L.o_ids) + (size_t)q * k;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = s_d[i];
    const uint32_t id = s_id[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) {
      const double dj = s_d[j];
      rank += (dj < d || (dj == d && s_id[j] < id)) ? 1u : 0u;
    }
    if (rank < (uint32_t)k) {
      o_id[rank] = id;
      o_sc[rank] = d;
    }
  }
  const uint32_t m = n < (uint32_t)k ? n : (uint32_t)k;
  for (uint32_t i = m + threadIdx.x; i < (uint32_t)k; i += blockDim.x) {
    o_id[i] = 0;
    o_sc[i] = 0.0;
  }
  if (threadIdx.x == 0) reinterpret_cast<uint32_t *>(out + L.o_counts)[q] = m;
  if (q == 0 && threadIdx.x < 5) {
    if (threadIdx.x < 4) {
      unsigned long long t = 0;
      for (int s = 0; s < n_shards; ++s)
        t += reinterpret_cast<const unsigned long long *>(gather + (size_t)s * shard_stride + L.o_stats)[threadIdx.x];
      reinterpret_cast<unsigned long long *>(out + L.o_stats)[threadIdx.x] = t;
    } else {
      int e = 0;
      for (int s = 0; s < n_shards; ++s) {
        const int es = *reinterpret_cast<const int *>(gather + (size_t)s * shard_stride + L.o_err);
        if (e == 0) e = es;
      }
      *reinterpret_cast<int *>(out + L.o_err) = e;
    }
  }
}

__global__ void slice_bits_kernel(const uint32_t *__restrict__ src, size_t src_words, uint32_t base,
                                  uint32_t *__restrict__ dst, size_t n_words) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const uint64_t bit0 = (uint64_t)base + (uint64_t)w * 32u;  // global bit of local bit 32w
  const size_t sw = (size_t)(bit0 >> 5);
  const uint32_t sh = (uint32_t)(bit0 & 31u);
  const uint32_t lo = sw < src_words ? src[sw] : 0u;
  const uint32_t hi = sw + 1 < src_words ? src[sw + 1] : 0u;
  uint32_t v = __funnelshift_r(lo, hi, sh);
  if (w == 0) v &= ~1u;  // local id 0 is the nil slot
  dst[w] = v;
}

// local ids -> global ids of an id-range shard (empty slots, id 0, stay 0)
__global__ void add_id_base_kernel(uint32_t *ids, size_t n, uint32_t base) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ids[i] != 0u) ids[i] += base;
}

}  // namespace

cudaError_t launch_add_id_base(uint32_t *ids, size_t n, uint32_t base, cudaStream_t stream) {
  if (n == 0 || base == 0) return cudaSuccess;
  add_id_base_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ids, n, base);
  return cudaGetLastError();
}

cudaError_t launch_merge_packed(int n_shards, uint32_t nq, int k, const unsigned char *gather, size_t shard_stride,
                                const BlobLayout &L, unsigned char *out, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  const size_t smem = (size_t)n_shards * k * (sizeof(double) + sizeof(uint32_t));
  if (n_shards > 64 || smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  merge_packed_kernel<<<nq, 128, smem, stream>>>(n_shards, nq, k, gather, shard_stride, L, out);
  return cudaGetLastError();
}

cudaError_t launch_slice_bits(const uint32_t *src, size_t src_words, uint32_t base, uint32_t *dst, size_t n_words,
                              cudaStream_t stream) {
  if (n_words == 0) return cudaSuccess;
  slice_bits_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(src, src_words, base, dst, n_words);
  return cudaGetLastError();
}

cudaError_t launch_flat_distances(const DevIndex &ix, const float *queries_raw, const float *queries_prepared,
                                  uint32_t nq, int mode, double *dist, cudaStream_t stream) {
  if (nq == 0 || ix.n == 0) return cudaSuccess;
  dim3 grid((ix.n + FT_ROWS - 1) / FT_ROWS, (nq + FT_Q - 1) / FT_Q);
  if (mode == 0) {
    flat_distances_kernel<0, KDBGPU_METRIC_L2><<<grid, FT_THREADS, 0, stream>>>(ix, queries_raw, ix.dim, nq, dist);
  } else if (ix.metric == KDBGPU_METRIC_COSINE) {
    flat_distances_kernel<1, KDBGPU_METRIC_COSINE>
        <<<grid, FT_THREADS, 0, stream>>>(ix, queries_prepared, ix.stride, nq, dist);
  } else {
    flat_distances_kernel<1, KDBGPU_METRIC_L2><<<grid, FT_THREADS, 0, stream>>>(ix, queries_raw, ix.dim, nq, dist);
  }
  return cudaGetLastError();
}

cudaError_t launch_flat_select(const DevIndex &ix, const double *dist, uint32_t nq, int k, const uint32_t *allow,
                               uint32_t *out_ids, double *out_scores, uint32_t *out_counts, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  if (k > SEL_MAXK) return cudaErrorInvalidValue;
  flat_select_kernel<<<nq, SEL_THREADS, 0, stream>>>(ix, dist, nq, k, allow, out_ids, out_scores, out_counts);
  return cudaGetLastError();
}

cudaError_t launch_merge_topk(int n_shards, uint32_t nq, int k, const uint32_t *ids, const double *scores,
                              const uint32_t *counts, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  const size_t smem = (size_t)n_shards * k * (sizeof(double) + sizeof(uint32_t));
  if (n_shards > 16 || smem > 96 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  merge_topk_kernel<<<nq, 128, smem, stream>>>(n_shards, nq, k, ids, scores, counts, out_ids, out_scores,
                                               out_counts);
  return cudaGetLastError();
}

}  // namespace kdb
