// build.cu — graph construction on the GPU ("next" row f-3 of SURVEY.md §8): a device-side
// restatement of (*Index).AddBatch / addBatchInternal (reference pkg/core/hnsw/hnsw_index.go:1466-2088)
// and, for an index still smaller than efConstruction, of the single Add it falls back to
// (:472-809, :1502-1513).  Deterministic, and bit-identical to the oracle's kdbo_add_batch in
// KDBO_ARITH_KERNEL mode: same searches, same request sets, same (distance, id) sort, same
// selectNeighbors (:2629-2701).
//
// Batch pipeline (one call = one AddBatch):
//   1. build_search_kernel   every new node searches the PRE-batch graph (:1789-1853)
//   2. count / scan / fill   LinkRequests grouped by adjacency row: node <- its candidates and
//                            candidate <- node (:1864-1895)
//   3. commit_kernel         one CTA per touched row: union, sort, dedupe; rows that overflow maxM
//                            are pruned by distance + selectNeighbors (:1926-2056)
#include "searcher.cuh"

namespace kdb {

using namespace dev;

namespace {

constexpr int kCommitThreads = 256;
constexpr int kCommitWarps = kCommitThreads / 32;
constexpr int kListSmem = 1024;  // (d, id) pairs held in shared memory; longer lists use global scratch

__device__ __forceinline__ uint32_t *row_ptr(uint32_t *adj0, uint32_t *upper_adj, const DevIndex &ix, uint32_t id,
                                             int level) {
  if (level == 0) return adj0 + (size_t)id * ix.deg0;
  return upper_adj + ((size_t)ix.upper_first[id] + (uint32_t)(level - 1)) * ix.degu;
}

// distanceBetweenNodes (hnsw_index.go:297-341): a stored row staged in shared memory (q4, int8 norm
// qnorm) against stored node `id`; int8 returns 1.0 when either norm is 0 (:325-327)
template <int METRIC>
__device__ __forceinline__ double dist_to_node(const DevIndex &ix, const float4 *q4, float qnorm, uint32_t id,
                                               int lane) {
  const float s = warp_reduce_row<METRIC>(q4, reinterpret_cast<const float4 *>(ix.vecs + (size_t)id * ix.row_words),
                                          ix.row_words >> 2, lane);
  if (METRIC == KIND_COS_I8) {
    if (qnorm == 0.f) return 1.0;
    return to_distance<METRIC>(s, qnorm, ix.norms[id]);
  }
  return to_distance<METRIC>(s);
}
template <int METRIC>
__device__ __forceinline__ float node_norm(const DevIndex &ix, uint32_t id) {
  return METRIC == KIND_COS_I8 ? ix.norms[id] : 0.f;
}

__device__ __forceinline__ void stage_vector(const DevIndex &ix, float4 *q4, uint32_t id, int tid, int nthreads) {
  const float4 *src = reinterpret_cast<const float4 *>(ix.vecs + (size_t)id * ix.row_words);
  for (uint32_t c = tid; c < (ix.row_words >> 2); c += nthreads) q4[c] = src[c];
}

// In-place bitonic sort of n2 (power of two) pairs by (d, id) or by id only.
template <bool BY_DIST>
__device__ void bitonic_sort(double *d, uint32_t *ids, uint32_t n2, int tid, int nthreads) {
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < n2; i += nthreads) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const uint32_t ii = ids[i], ij = ids[j];
          bool gt;
          if (BY_DIST) {
            const double di = d[i], dj = d[j];
            gt = di > dj || (di == dj && ii > ij);
            if (gt == up) {
              d[i] = dj;
              d[j] = di;
            }
          } else {
            gt = ii > ij;
          }
          if (gt == up) {
            ids[i] = ij;
            ids[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
}

// selectNeighbors (hnsw_index.go:2629-2701) over candidates (ids[i], d[i]), i < n, in the given
// order.  CTA-cooperative; q4 is scratch for one staged vector; sel/disc hold candidate indices.
// Returns the number selected (<= m); sel[0..ret) lists indices into ids[].
template <int METRIC>
__device__ int select_neighbors_dev(const DevIndex &ix, const double *d, const uint32_t *ids, uint32_t n, int m,
                                    float4 *q4, uint32_t *sel, uint32_t *disc, int *flag, int tid, int nthreads) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  if (n <= (uint32_t)m) {  // :2634-2636
    for (uint32_t i = tid; i < n; i += nthreads) sel[i] = i;
    __syncthreads();
    return (int)n;
  }
  int nr = 0, nd = 0;
  for (uint32_t w = 0; w < n && nr < m; ++w) {  // :2642
    if (nr == 0) {
      if (tid == 0) sel[0] = w;
      nr = 1;
      continue;
    }
    __syncthreads();
    stage_vector(ix, q4, ids[w], tid, nthreads);
    if (tid == 0) *flag = 0;
    __syncthreads();
    const double de = d[w];
    const float qn = node_norm<METRIC>(ix, ids[w]);
    for (int r = warp; r < nr; r += nwarps) {  // :2652-2679 (any closer kept neighbour discards e)
      const double dd = dist_to_node<METRIC>(ix, q4, qn, ids[sel[r]], lane);
      if (lane == 0 && dd < de) *flag = 1;
    }
    __syncthreads();
    const bool bad = *flag != 0;
    if (!bad) {
      if (tid == 0) sel[nr] = w;
      nr++;
    } else {
      if (tid == 0 && nd < m) disc[nd] = w;  // only the first m discards can ever be topped up
      nd++;
    }
  }
  __syncthreads();
  for (int i = 0; i < nd && i < m && nr < m; ++i) {  // :2689-2698 top up from the discarded list
    if (tid == 0) sel[nr] = disc[i];
    nr++;
  }
  __syncthreads();
  return nr;
}

struct BuildSearchArgs {
  uint32_t start_id, count;
  uint32_t pre_entry;
  int pre_max;
  int efc;
  const uint32_t *out_off;  // [count]: first output slot of node i; slot + l holds level l
  uint32_t *cand_ids;       // [n_slots][efc]
  uint32_t *cand_cnt;       // [n_slots]
};

// phase 1 (:1789-1853): one warp per new node
template <int SLOTS, int METRIC, int CPL>
__global__ void __launch_bounds__(32)
    build_search_kernel(const DevIndex ix, const SearchArgs a, const BuildSearchArgs b) {
  extern __shared__ __align__(128) unsigned char smem[];
  Searcher<SLOTS, METRIC, CPL> s(ix, a, smem);
  s.init_barriers();
  const int lane = s.lane;
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(a.work_counter, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= b.count) break;
    const uint32_t id = b.start_id + i;
    const int node_level = ix.levels[id];
    s.load_row_as_query(id);  // currObj := storedVector (:674, :1806)
    uint32_t ep = b.pre_entry;
    for (int l = b.pre_max; l > node_level; --l) {  // zoom in (:1819-1824)
      const int n = s.search_layer(l, 1, ep);
      if (n > 0) ep = s.sm.res[0].id;
      s.clear_visited(true);
    }
    const uint32_t slot0 = b.out_off[i];
    for (int l = node_level < b.pre_max ? node_level : b.pre_max; l >= 0; --l) {  // (:1827-1850)
      const int n = s.search_layer(l, b.efc, ep);
      uint32_t *out = b.cand_ids + (size_t)(slot0 + l) * b.efc;
      uint32_t first = 0;
      if (lane == 0) {
        for (int j = n - 1; j >= 0; --j) out[j] = s.res.pop().id;  // ascending, as :2596-2604
        b.cand_cnt[slot0 + l] = n > 0 ? (uint32_t)n : 0u;
        if (n > 0) first = out[0];
      }
      first = __shfl_sync(0xffffffffu, first, 0);
      s.clear_visited(l > 0);
      if (n > 0) ep = first;
    }
    if (lane == 0 && s.overflow) {
      atomicExch(a.err_flag, KDBGPU_ERR_OVERFLOW);
      s.overflow = false;
    }
    __syncwarp();
  }
}

struct RequestArgs {
  const uint32_t *slot_node;
  const uint8_t *slot_level;
  const uint32_t *cand_ids;
  const uint32_t *cand_cnt;
  uint32_t n_slots;
  int efc;
  uint32_t up_base;  // adjacency-row index of upper row 0 (= capacity + 1)
};

__device__ __forceinline__ uint32_t row_index(const DevIndex &ix, uint32_t up_base, uint32_t id, int level) {
  return level == 0 ? id : up_base + ix.upper_first[id] + (uint32_t)(level - 1);
}

// phase 2 (:1864-1895): FILL == false counts requests per row, FILL == true scatters the sources
template <bool FILL>
__global__ void request_kernel(const DevIndex ix, const RequestArgs r, uint32_t *cnt, const uint32_t *off,
                               uint32_t *srcs) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t sr = (uint32_t)(idx / (uint32_t)r.efc);
  const uint32_t j = (uint32_t)(idx % (uint32_t)r.efc);
  if (sr >= r.n_slots || j >= r.cand_cnt[sr]) return;
  const uint32_t node = r.slot_node[sr];
  const int level = r.slot_level[sr];
  const uint32_t c = r.cand_ids[(size_t)sr * r.efc + j];
  const uint32_t row_direct = row_index(ix, r.up_base, node, level);
  const uint32_t row_reverse = row_index(ix, r.up_base, c, level);
  if (!FILL) {
    atomicAdd(&cnt[row_direct], 1u);
    atomicAdd(&cnt[row_reverse], 1u);
  } else {
    srcs[off[row_direct] + atomicAdd(&cnt[row_direct], 1u)] = c;      // node <- candidate
    srcs[off[row_reverse] + atomicAdd(&cnt[row_reverse], 1u)] = node;  // candidate <- node
  }
}

// exclusive scan of cnt[0..n) into off[0..n], single CTA; also lists the rows with cnt > 0
__global__ void __launch_bounds__(1024) scan_rows_kernel(const uint32_t *cnt, uint32_t n, uint32_t *off,
                                                          uint32_t *active, uint32_t *n_active) {
  __shared__ uint32_t part[1024];
  __shared__ uint32_t part_act[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t chunk = (n + 1023) / 1024;
  const uint32_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
  uint32_t s = 0, na = 0;
  for (uint32_t i = b; i < e; ++i) {
    s += cnt[i];
    na += cnt[i] ? 1u : 0u;
  }
  part[t] = s;
  part_act[t] = na;
  __syncthreads();
  for (uint32_t o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    uint32_t v = 0, va = 0;
    if (t >= o) {
      v = part[t - o];
      va = part_act[t - o];
    }
    __syncthreads();
    part[t] += v;
    part_act[t] += va;
    __syncthreads();
  }
  uint32_t run = t ? part[t - 1] : 0u, ra = t ? part_act[t - 1] : 0u;
  for (uint32_t i = b; i < e; ++i) {
    off[i] = run;
    run += cnt[i];
    if (cnt[i]) active[ra++] = i;
  }
  if (t == 1023) {
    off[n] = part[1023];
    *n_active = part_act[1023];
  }
}

struct CommitArgs {
  uint32_t *adj0;
  uint32_t *upper_adj;
  const uint32_t *upper_node;  // owner of upper row u
  const uint8_t *upper_level;  // its level
  uint32_t up_base;
  const uint32_t *off;
  const uint32_t *srcs;
  const uint32_t *active;
  const uint32_t *n_active;
  uint32_t *work_counter;
  double *scratch_d;      // [grid][scratch_cap]
  uint32_t *scratch_ids;  // [grid][scratch_cap]
  uint32_t scratch_cap;   // power of two
  int *err_flag;
};

// phase 3 (:1926-2056)
template <int METRIC>
__global__ void __launch_bounds__(kCommitThreads) commit_kernel(const DevIndex ix, const CommitArgs c) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4 *q4 = reinterpret_cast<float4 *>(smem);
  double *sd = reinterpret_cast<double *>(smem + (size_t)ix.stride * sizeof(float));
  uint32_t *sids = reinterpret_cast<uint32_t *>(sd + kListSmem);
  uint32_t *sel = sids + kListSmem;
  uint32_t *disc = sel + kMaxDeg;
  __shared__ uint32_t s_g, s_cnt;
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_g = atomicAdd(c.work_counter, 1u);
    __syncthreads();
    const uint32_t g = s_g;
    if (g >= *c.n_active) break;
    const uint32_t row = c.active[g];
    uint32_t t;
    int level;
    if (row < c.up_base) {
      t = row;
      level = 0;
    } else {
      t = c.upper_node[row - c.up_base];
      level = c.upper_level[row - c.up_base];
    }
    if (ix.levels[t] < 0 || (ix.deleted && bit_test(ix.deleted, t))) continue;  // :1933-1936
    const int max_m = level == 0 ? (int)ix.deg0 : (int)ix.degu;                // :2010-2013
    uint32_t *rowp = row_ptr(c.adj0, c.upper_adj, ix, t, level);
    // existing connections (:1980-1986) + requested sources
    uint32_t n_exist = 0;
    for (; n_exist < (uint32_t)max_m; ++n_exist)
      if (rowp[n_exist] == 0u) break;
    const uint32_t o0 = c.off[row], n_new = c.off[row + 1] - o0;
    const uint32_t total = n_exist + n_new;
    uint32_t n2 = 1;
    while (n2 < total) n2 <<= 1;
    double *d = sd;
    uint32_t *ids = sids;
    if (n2 > (uint32_t)kListSmem) {
      if (n2 > c.scratch_cap) {
        if (tid == 0) atomicExch(c.err_flag, KDBGPU_ERR_OVERFLOW);
        continue;
      }
      d = c.scratch_d + (size_t)blockIdx.x * c.scratch_cap;
      ids = c.scratch_ids + (size_t)blockIdx.x * c.scratch_cap;
    }
    for (uint32_t i = tid; i < n2; i += kCommitThreads) {
      uint32_t v = 0xffffffffu;
      if (i < n_exist)
        v = rowp[i];
      else if (i < total)
        v = c.srcs[o0 + (i - n_exist)];
      ids[i] = v;
    }
    __syncthreads();
    bitonic_sort<false>(d, ids, n2, tid, kCommitThreads);  // slices.Sort (:1990)
    // drop self + duplicates (:1992-2008), order preserved; warp 0 compacts in place
    if (warp == 0) {
      uint32_t w = 0;
      for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t i = base + lane;
        uint32_t v = 0xffffffffu;
        bool keep = false;
        if (i < total) {
          v = ids[i];
          const uint32_t prev = i > 0 ? ids[i - 1] : 0xffffffffu;
          keep = v != t && v != prev && v != 0xffffffffu;
        }
        __syncwarp();
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        if (keep) ids[w + __popc(km & ((1u << lane) - 1u))] = v;  // w + rank <= i: never ahead of the reads
        w += __popc(km);
        __syncwarp();
      }
      if (lane == 0) s_cnt = w;
    }
    __syncthreads();
    const uint32_t cnt = s_cnt;
    if (cnt <= (uint32_t)max_m) {  // :2016-2018: the ascending-id list replaces the row
      for (uint32_t i = tid; i < (uint32_t)max_m; i += kCommitThreads) rowp[i] = i < cnt ? ids[i] : 0u;
      continue;
    }
    // prune (:2019-2042): distances from t, ascending (distance, id), selectNeighbors
    stage_vector(ix, q4, t, tid, kCommitThreads);
    __syncthreads();
    uint32_t c2 = 1;
    while (c2 < cnt) c2 <<= 1;
    for (uint32_t i = warp; i < c2; i += kCommitWarps) {
      double dv = __longlong_as_double(0x7ff0000000000000LL);  // +inf: padding and deleted targets sort last
      if (i < cnt) {
        const uint32_t e = ids[i];
        const bool dead = ix.levels[e] < 0 || (ix.deleted && bit_test(ix.deleted, e));  // :2024-2026
        if (!dead)
          dv = dist_to_node<METRIC>(ix, q4, node_norm<METRIC>(ix, t), e, lane);
        else if (lane == 0)
          ids[i] = 0xffffffffu;
      } else if (lane == 0) {
        ids[i] = 0xffffffffu;
      }
      if (lane == 0) d[i] = dv;
    }
    __syncthreads();
    bitonic_sort<true>(d, ids, c2, tid, kCommitThreads);  // :2028-2036
    if (tid == 0) {
      uint32_t nc = cnt;
      while (nc > 0 && ids[nc - 1] == 0xffffffffu) nc--;
      s_cnt = nc;
    }
    __syncthreads();
    const uint32_t nc = s_cnt;
    const int ns = select_neighbors_dev<METRIC>(ix, d, ids, nc, max_m, q4, sel, disc, &s_flag, tid, kCommitThreads);
    // node.Connections[lvl] = finalConns (:2051); stage through registers: sel indexes ids[]
    uint32_t mine = 0;
    if (tid < max_m) mine = tid < ns ? ids[sel[tid]] : 0u;
    __syncthreads();
    if (tid < max_m) rowp[tid] = mine;
  }
}

// ---- the single-Add fallback for an index smaller than efConstruction (:1502-1513 -> Add :472-809) ----
struct SeqAddArgs {
  uint32_t start_id, count;
  int efc;
  uint32_t *adj0;
  uint32_t *upper_adj;
  uint32_t *entry_io;  // [0] = entry, [1] = max_level + 1 (0 = empty index)
};

constexpr int kSeqSlots = 4;

template <int METRIC>
__global__ void __launch_bounds__(32) seq_add_kernel(const DevIndex ix_in, const SearchArgs a, const SeqAddArgs b) {
  extern __shared__ __align__(128) unsigned char smem[];
  DevIndex ix = ix_in;
  Searcher<kSeqSlots, METRIC, 0> s(ix, a, smem);
  const size_t base = smem_layout(ix.stride, slot_pitch_words(ix.stride, ix.row_words, ix.kind), a.ef, kSeqSlots, a.cand_smem, ix.deg0 > ix.degu ? ix.deg0 : ix.degu,
                                  true, ix.kind, nullptr, nullptr);
  float4 *q2 = reinterpret_cast<float4 *>(smem + base);
  const int LL = (b.efc + kMaxDeg + 3) & ~1;  // list length, even so the f64 array after it stays aligned
  double *cd = reinterpret_cast<double *>(smem + base + (size_t)ix.stride * sizeof(float));
  uint32_t *cid = reinterpret_cast<uint32_t *>(cd + LL);
  uint32_t *sel = cid + LL;
  uint32_t *disc = sel + LL;
  double *ad = reinterpret_cast<double *>(disc + LL + (LL & 1));
  uint32_t *aid = reinterpret_cast<uint32_t *>(ad + (kMaxDeg + 2));
  uint32_t *fwd = aid + (kMaxDeg + 2);
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = 0;
  s.init_barriers();
  uint32_t entry = b.entry_io[0];
  int cur_max = (int)b.entry_io[1] - 1;
  for (uint32_t it = 0; it < b.count; ++it) {
    const uint32_t id = b.start_id + it;
    const int level = ix.levels[id];
    ix.n = id;  // nodes registered so far (the new one is already in nodes[], :652-655)
    if (cur_max == -1) {  // first node (:658-670)
      entry = id;
      cur_max = level;
      continue;
    }
    __syncthreads();
    s.load_row_as_query(id);
    uint32_t ep = entry;
    for (int l = cur_max; l > level; --l) {  // :685-690
      const int n = s.search_layer(l, 1, ep);
      if (n > 0) ep = s.sm.res[0].id;
      s.clear_visited(true);
    }
    const int top = level < cur_max ? level : cur_max;  // :694-697
    for (int l = top; l >= 0; --l) {                    // :699
      const int n = s.search_layer(l, b.efc, ep);
      if (tid == 0)
        for (int j = n - 1; j >= 0; --j) {
          const HeapEntry e = s.res.pop();
          cd[j] = e.d;
          cid[j] = e.id;
        }
      s.clear_visited(l > 0);
      if (n < 0) continue;  // :702-704
      const int max_m = l == 0 ? (int)ix.deg0 : (int)ix.degu;  // :707-710
      const int ns = select_neighbors_dev<METRIC>(ix, cd, cid, (uint32_t)n, max_m, q2, sel, disc, &s_flag, tid, 32);
      uint32_t *rowp = row_ptr(b.adj0, b.upper_adj, ix, id, l);
      for (int i = tid; i < max_m; i += 32) fwd[i] = i < ns ? cid[sel[i]] : 0u;
      __syncthreads();
      for (int i = tid; i < max_m; i += 32) rowp[i] = fwd[i];  // forward links (:717-722)
      __threadfence_block();
      __syncthreads();
      for (int si = 0; si < ns; ++si) {  // reverse links (:725-783)
        const uint32_t nb = fwd[si];
        if (ix.levels[nb] < 0 || (ix.deleted && bit_test(ix.deleted, nb))) continue;  // :731-734
        if (l > ix.levels[nb]) continue;
        uint32_t *nrow = row_ptr(b.adj0, b.upper_adj, ix, nb, l);
        uint32_t ccount = 0;
        for (; ccount < (uint32_t)max_m; ++ccount)
          if (nrow[ccount] == 0u) break;
        if ((int)ccount < max_m) {  // :748-752 append
          __syncthreads();
          if (tid == 0) nrow[ccount] = id;
          __threadfence_block();
          __syncthreads();
          continue;
        }
        // prune (:753-771): candidates in list order + the new node last, NOT sorted
        __syncthreads();
        stage_vector(ix, q2, nb, tid, 32);
        __syncthreads();
        uint32_t na = 0;
        for (uint32_t j = 0; j < ccount; ++j) {  // uniform over the CTA
          const uint32_t e = nrow[j];
          if (ix.levels[e] >= 0 && !(ix.deleted && bit_test(ix.deleted, e))) {
            if (tid == 0) aid[na] = e;
            na++;
          }
        }
        if (tid == 0) aid[na] = id;
        na++;
        __syncthreads();
        for (uint32_t j = warp; j < na; j += 1) {
          const double dv = dist_to_node<METRIC>(ix, q2, node_norm<METRIC>(ix, nb), aid[j], lane);
          if (lane == 0) ad[j] = dv;
        }
        __syncthreads();
        const int nbest = select_neighbors_dev<METRIC>(ix, ad, aid, na, max_m, q2, sel, disc, &s_flag, tid, 32);
        for (int i = tid; i < max_m; i += 32) fwd[kMaxDeg + i] = i < nbest ? aid[sel[i]] : 0u;
        __syncthreads();
        for (int i = tid; i < max_m; i += 32) nrow[i] = fwd[kMaxDeg + i];  // :774-782
        __threadfence_block();
        __syncthreads();
      }
      if (n > 0) ep = cid[0];  // :786-788
      __syncthreads();
    }
    if (level > cur_max) {  // :793-801
      cur_max = level;
      entry = id;
    }
  }
  __syncthreads();
  if (tid == 0) {
    b.entry_io[0] = entry;
    b.entry_io[1] = (uint32_t)(cur_max + 1);
    if (s.overflow) atomicExch(a.err_flag, KDBGPU_ERR_OVERFLOW);
  }
}

constexpr int kBuildSlots = 8;

__host__ inline int build_cpl_of(const DevIndex &ix) {
  const uint32_t c = ix.stride / 128;
  return (c == 1 || c == 2 || c == 3 || c == 4 || c == 6 || c == 8 || c == 12) ? (int)c : 0;
}

template <int METRIC, int CPL>
cudaError_t launch_build_search_one(const DevIndex &ix, const SearchArgs &a, const BuildSearchArgs &b, int grid,
                                    size_t smem, cudaStream_t stream) {
  auto kern = build_search_kernel<kBuildSlots, METRIC, CPL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, 32, smem, stream>>>(ix, a, b);
  return cudaGetLastError();
}
template <int METRIC, int CPL>
int build_search_occ_one(size_t smem) {
  auto kern = build_search_kernel<kBuildSlots, METRIC, CPL>;
  int nb = 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return nb;
}
#define KDB_BUILD_CPL(EXPR)                                       \
    switch (build_cpl_of(ix)) {                                   \
      case 1: { constexpr int CP = 1; EXPR; } break;              \
      case 2: { constexpr int CP = 2; EXPR; } break;              \
      case 3: { constexpr int CP = 3; EXPR; } break;              \
      case 4: { constexpr int CP = 4; EXPR; } break;              \
      case 6: { constexpr int CP = 6; EXPR; } break;              \
      case 8: { constexpr int CP = 8; EXPR; } break;              \
      case 12: { constexpr int CP = 12; EXPR; } break;            \
      default: { constexpr int CP = 0; EXPR; } break;             \
    }
#define KDB_BUILD_DISPATCH(EXPR)                                                   \
  switch (ix.kind) {                                                               \
    case KIND_COS_F32: { constexpr int MT = KIND_COS_F32; KDB_BUILD_CPL(EXPR) } break; \
    case KIND_L2_F16: { constexpr int MT = KIND_L2_F16; KDB_BUILD_CPL(EXPR) } break;   \
    case KIND_COS_I8: { constexpr int MT = KIND_COS_I8; KDB_BUILD_CPL(EXPR) } break;   \
    default: { constexpr int MT = KIND_L2_F32; KDB_BUILD_CPL(EXPR) } break;            \
  }
#define KDB_KIND_DISPATCH(EXPR)                                          \
  switch (ix.kind) {                                                     \
    case KIND_COS_F32: { constexpr int MT = KIND_COS_F32; EXPR; } break; \
    case KIND_L2_F16: { constexpr int MT = KIND_L2_F16; EXPR; } break;   \
    case KIND_COS_I8: { constexpr int MT = KIND_COS_I8; EXPR; } break;   \
    default: { constexpr int MT = KIND_L2_F32; EXPR; } break;            \
  }

template <int MT>
cudaError_t launch_commit_kind(const DevIndex &ix, const CommitArgs &c, int grid, size_t smem, cudaStream_t stream) {
  auto kern = commit_kernel<MT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kCommitThreads, smem, stream>>>(ix, c);
  return cudaGetLastError();
}
template <int MT>
cudaError_t launch_seq_kind(const DevIndex &ix, const SearchArgs &a, const SeqAddArgs &b, size_t smem,
                            cudaStream_t stream) {
  auto kern = seq_add_kernel<MT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<1, 32, smem, stream>>>(ix, a, b);
  return cudaGetLastError();
}

cudaError_t launch_build_search_cfg(const DevIndex &ix, const SearchArgs &a, const BuildSearchArgs &b, int grid,
                                    size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaErrorInvalidConfiguration;
  KDB_BUILD_DISPATCH(e = (launch_build_search_one<MT, CP>(ix, a, b, grid, smem, stream)))
  return e;
}

}  // namespace

// ---- host-side launchers used by api.cu -------------------------------------------------------
struct BuildLaunch {
  // inputs
  uint32_t start_id, count;
  uint32_t pre_entry;
  int pre_max;
  int efc;
  uint32_t n_slots;
  uint32_t up_base;
  uint32_t n_rows;  // adjacency rows addressable (up_base + upper rows in use)
  // device buffers
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

size_t commit_smem_bytes(const DevIndex &ix) {
  return (size_t)ix.stride * sizeof(float) + (size_t)kListSmem * (sizeof(double) + sizeof(uint32_t)) +
         2 * (size_t)kMaxDeg * sizeof(uint32_t);
}

size_t seq_add_smem_bytes(const DevIndex &ix, int efc, uint32_t cand_smem) {
  const size_t base = smem_layout(ix.stride, slot_pitch_words(ix.stride, ix.row_words, ix.kind), efc, kSeqSlots, cand_smem, ix.deg0 > ix.degu ? ix.deg0 : ix.degu, true, ix.kind,
                                  nullptr, nullptr);
  const size_t lists = (size_t)((efc + kMaxDeg + 3) & ~1) + 2;
  return base + (size_t)ix.stride * sizeof(float) + lists * (sizeof(double) + 3 * sizeof(uint32_t)) +
         (size_t)(kMaxDeg + 2) * (sizeof(double) + sizeof(uint32_t)) + 2 * (size_t)kMaxDeg * sizeof(uint32_t) + 64;
}

size_t build_search_smem(const DevIndex &ix, int efc, uint32_t cand_smem) {
  return smem_layout(ix.stride, slot_pitch_words(ix.stride, ix.row_words, ix.kind), efc, kBuildSlots, cand_smem, ix.deg0 > ix.degu ? ix.deg0 : ix.degu,
                     build_cpl_of(ix) == 0, ix.kind, nullptr, nullptr);
}

int build_search_occupancy(const DevIndex &ix, int efc, uint32_t cand_smem) {
  const size_t smem = build_search_smem(ix, efc, cand_smem);
  if (smem > 227 * 1024) return 0;
  int nb = 0;
  KDB_BUILD_DISPATCH(nb = (build_search_occ_one<MT, CP>(smem)))
  return nb;
}

// AddBatch on an index with >= efConstruction nodes
cudaError_t launch_add_batch(const DevIndex &ix, const SearchArgs &a, const BuildLaunch &L, int search_grid,
                             cudaStream_t stream) {
  cudaError_t e;
  BuildSearchArgs b;
  b.start_id = L.start_id;
  b.count = L.count;
  b.pre_entry = L.pre_entry;
  b.pre_max = L.pre_max;
  b.efc = L.efc;
  b.out_off = L.out_off;
  b.cand_ids = L.cand_ids;
  b.cand_cnt = L.cand_cnt;
  const size_t smem = build_search_smem(ix, L.efc, a.cand_smem);
  e = launch_build_search_cfg(ix, a, b, search_grid, smem, stream);
  if (e != cudaSuccess) return e;
  RequestArgs r;
  r.slot_node = L.slot_node;
  r.slot_level = L.slot_level;
  r.cand_ids = L.cand_ids;
  r.cand_cnt = L.cand_cnt;
  r.n_slots = L.n_slots;
  r.efc = L.efc;
  r.up_base = L.up_base;
  const uint64_t n_req_threads = (uint64_t)L.n_slots * (uint64_t)L.efc;
  const int rb = 256;
  const unsigned rgrid = (unsigned)((n_req_threads + rb - 1) / rb);
  if ((e = cudaMemsetAsync(L.row_cnt, 0, (size_t)(L.n_rows + 1) * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  request_kernel<false><<<rgrid, rb, 0, stream>>>(ix, r, L.row_cnt, nullptr, nullptr);
  scan_rows_kernel<<<1, 1024, 0, stream>>>(L.row_cnt, L.n_rows, L.row_off, L.active, L.n_active);
  if ((e = cudaMemsetAsync(L.row_cnt, 0, (size_t)(L.n_rows + 1) * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  request_kernel<true><<<rgrid, rb, 0, stream>>>(ix, r, L.row_cnt, L.row_off, L.srcs);
  if ((e = cudaMemsetAsync(L.work_counter, 0, sizeof(uint32_t), stream)) != cudaSuccess) return e;
  CommitArgs c;
  c.adj0 = L.adj0;
  c.upper_adj = L.upper_adj;
  c.upper_node = L.upper_node;
  c.upper_level = L.upper_level;
  c.up_base = L.up_base;
  c.off = L.row_off;
  c.srcs = L.srcs;
  c.active = L.active;
  c.n_active = L.n_active;
  c.work_counter = L.work_counter;
  c.scratch_d = L.scratch_d;
  c.scratch_ids = L.scratch_ids;
  c.scratch_cap = L.scratch_cap;
  c.err_flag = a.err_flag;
  const size_t csmem = commit_smem_bytes(ix);
  e = cudaSuccess;
  KDB_KIND_DISPATCH(e = (launch_commit_kind<MT>(ix, c, L.commit_grid, csmem, stream)))
  return e;
}

// the single-Add fallback: `count` sequential Adds by one CTA
cudaError_t launch_seq_add(const DevIndex &ix, const SearchArgs &a, uint32_t start_id, uint32_t count, int efc,
                           uint32_t *adj0, uint32_t *upper_adj, uint32_t *entry_io, cudaStream_t stream) {
  SeqAddArgs b;
  b.start_id = start_id;
  b.count = count;
  b.efc = efc;
  b.adj0 = adj0;
  b.upper_adj = upper_adj;
  b.entry_io = entry_io;
  const size_t smem = seq_add_smem_bytes(ix, efc, a.cand_smem);
  cudaError_t e = cudaSuccess;
  KDB_KIND_DISPATCH(e = (launch_seq_kind<MT>(ix, a, b, smem, stream)))
  return e;
}

}  // namespace kdb
