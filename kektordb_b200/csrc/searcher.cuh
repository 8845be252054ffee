// searcher.cuh — the per-query HNSW layer search shared by the query kernel (search.cu) and the
// graph-construction kernels (build.cu).  Device code, namespace kdb::dev.
//
// ONE WARP PER QUERY.  Stands in for searchLayerUnlocked (reference
// pkg/core/hnsw/hnsw_index.go:2351-2611) with the heaps of hnsw_heap.go:18-156 and the visited
// BitSet of bitset.go.  Per hop the warp
//   A. (lane 0)    pops the nearest candidate and tests the early exit (:2497-2506)
//   B. (all lanes) reads the adjacency row (one neighbour per lane), test-and-sets the visited
//                  bitset, applies the allow-list BEFORE any distance (:2537-2549) and compacts
//                  the survivors in row order
//   C. (all lanes) streams the survivors' rows HBM -> shared memory with 1-D bulk copies
//                  (cp.async.bulk + mbarrier) in GROUPS of G = SLOTS / 2 rows, two groups in flight,
//                  ONE mbarrier per group; the G rows of a group are reduced against the query together
//                  (G independent chains) in the fixed "kernel order" (kdb_internal.cuh), the warp
//                  reduction of the G lane partials is done transposed (G - 1 + log2(32 / G) shuffles
//                  for G rows instead of 5 G; same additions, same order), the group's slots are
//                  refilled, and every row is pre-tested in parallel against the worst kept distance
//   D. (lane 0)    the rows that pass get the reference's heap update (:2571-2591), in row order,
//                  while the next groups of the hop are in flight.
// No CTA-wide barrier exists on this path: a CTA is a single warp, the SM interleaves ~7 of them.
// The two binary heaps are the reference's own algorithms run by lane 0, so ids, order and scores
// are bit-identical to the oracle in KDBO_ARITH_KERNEL mode — ties included.
#pragma once
#include "kdb_internal.cuh"

namespace kdb {
namespace dev {

constexpr int kMaxDeg = 256;   // max neighbours per adjacency row (2M <= 256)
constexpr int kMarkCap = 256;  // visited marks logged per upper-level search before a full clear

struct Ctl {
  uint32_t n_marked;
  uint32_t cur;  // scratch word for callers (build kernels publish the next entry point here)
  uint32_t pad[2];
};

struct SmemPtrs {
  float4 *q4;
  float *slots;
  uint64_t *bars;
  HeapEntry *res;
  HeapEntry *cand;
  uint32_t *eval_id;
  uint32_t *eval_del;
  float *eval_norm;  // int8: stored norm of every neighbour to evaluate (quantizedNorms[id])
  int *eval_thr;     // int8: dot <= eval_thr[j] proves "not admitted" without the float64 divide (see collect_neighbours)
  uint32_t *marked;
  Ctl *ctl;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// shared-memory carve-up of one warp-CTA, identical on host (sizing) and device (pointers)
// kind = the index's distance kind: only int8 rows need the per-neighbour norm / threshold lists
// slot pitch of a row in shared memory (32-bit words).  float rows: the padded stride (the tail of a shorter row is
// zeroed once and must stay zero: 0 x garbage could be NaN).  int8 rows: the row's own pitch — the query's padding
// columns are zero and an integer 0 x anything is 0, so whatever the last column pass reads past the row (the next
// slot, or the regions behind the slots) contributes nothing; 768-byte rows then take 768, not 1024 bytes per slot.
__host__ __device__ inline uint32_t slot_pitch_words(uint32_t stride, uint32_t row_words, int kind) {
  return kind == KIND_COS_I8 ? row_words : stride;
}
__host__ __device__ inline size_t smem_layout(uint32_t stride, uint32_t slot_words, int ef, int slots, uint32_t cand_smem,
                                              uint32_t deg_max, bool q_in_smem, int kind, unsigned char *base, SmemPtrs *p) {
  const uint32_t dm = (deg_max + 31u) & ~31u;
  size_t off = 0;
  const size_t o_slots = off;
  off += (size_t)slots * slot_words * sizeof(float);
  const size_t o_q = off;
  if (q_in_smem) off += (size_t)stride * sizeof(float);
  const size_t o_res = off;
  off += (size_t)(ef + 1) * sizeof(HeapEntry);
  const size_t o_cand = off;
  off += (size_t)cand_smem * sizeof(HeapEntry);
  const size_t o_bars = off;
  off += (size_t)slots * sizeof(uint64_t);
  off = align_up(off, 16);  // the ids of a group are fetched with one vector load
  const size_t o_evalid = off;
  off += (size_t)dm * sizeof(uint32_t);
  const size_t o_evaldel = off;
  off += (size_t)dm * sizeof(uint32_t);
  off = align_up(off, 16);
  const size_t o_evalnorm = off;
  if (kind == KIND_COS_I8) off += (size_t)dm * sizeof(float);
  const size_t o_evalthr = off;
  if (kind == KIND_COS_I8) off += (size_t)dm * sizeof(int);
  const size_t o_marked = off;
  off += (size_t)kMarkCap * sizeof(uint32_t);
  const size_t o_ctl = off;
  off += sizeof(Ctl);
  off = align_up(off, 128);
  if (p) {
    p->slots = reinterpret_cast<float *>(base + o_slots);
    p->q4 = reinterpret_cast<float4 *>(base + o_q);
    p->res = reinterpret_cast<HeapEntry *>(base + o_res);
    p->cand = reinterpret_cast<HeapEntry *>(base + o_cand);
    p->bars = reinterpret_cast<uint64_t *>(base + o_bars);
    p->eval_id = reinterpret_cast<uint32_t *>(base + o_evalid);
    p->eval_del = reinterpret_cast<uint32_t *>(base + o_evaldel);
    p->eval_norm = reinterpret_cast<float *>(base + o_evalnorm);
    p->eval_thr = reinterpret_cast<int *>(base + o_evalthr);
    p->marked = reinterpret_cast<uint32_t *>(base + o_marked);
    p->ctl = reinterpret_cast<Ctl *>(base + o_ctl);
  }
  return off;
}

// ---- the reference's binary heaps, operated by one lane ------------------------------------
// Hole-based sifts: same final arrangement as the swap-based up()/down() of hnsw_heap.go.
struct CandHeap {  // minHeap (hnsw_heap.go:18-89); first `cap_s` entries in smem, rest in global
  HeapEntry *s;
  HeapEntry *g;
  uint32_t cap_s, cap_g;
  uint32_t n;
  __device__ __forceinline__ HeapEntry get(uint32_t i) const { return i < cap_s ? s[i] : g[i - cap_s]; }
  __device__ __forceinline__ void set(uint32_t i, const HeapEntry &e) {
    if (i < cap_s)
      s[i] = e;
    else
      g[i - cap_s] = e;
  }
  __device__ bool push(const HeapEntry &x) {  // Push + up (:33-36, :53-63)
    if (n >= cap_s + cap_g) return false;
    uint32_t j = n++;
    while (j > 0) {
      uint32_t i = (j - 1) >> 1;
      HeapEntry pi = get(i);
      if (!(x.d < pi.d)) break;
      set(j, pi);
      j = i;
    }
    set(j, x);
    return true;
  }
  __device__ HeapEntry pop() {  // Pop + down (:39-51, :65-83)
    HeapEntry top = get(0);
    HeapEntry last = get(n - 1);
    n--;
    if (n > 0) {
      uint32_t i = 0;
      for (;;) {
        uint32_t j1 = 2 * i + 1;
        if (j1 >= n) break;
        uint32_t j = j1;
        HeapEntry cj = get(j1);
        if (j1 + 1 < n) {
          HeapEntry c2 = get(j1 + 1);
          if (c2.d < cj.d) {
            j = j1 + 1;
            cj = c2;
          }
        }
        if (!(cj.d < last.d)) break;
        set(i, cj);
        i = j;
      }
      set(i, last);
    }
    return top;
  }
};

struct ResHeap {  // maxHeap (hnsw_heap.go:91-156), always in shared memory (ef + 1 entries)
  HeapEntry *a;
  int n;
  __device__ void push(const HeapEntry &x) {  // (:105-108, :122-132)
    int j = n++;
    while (j > 0) {
      int i = (j - 1) >> 1;
      HeapEntry pi = a[i];
      if (!(x.d > pi.d)) break;
      a[j] = pi;
      j = i;
    }
    a[j] = x;
  }
  __device__ HeapEntry pop() {  // (:110-120, :134-151)
    HeapEntry top = a[0];
    HeapEntry last = a[n - 1];
    n--;
    if (n > 0) {
      int i = 0;
      for (;;) {
        int j1 = 2 * i + 1;
        if (j1 >= n) break;
        int j = j1;
        HeapEntry cj = a[j1];
        if (j1 + 1 < n) {
          HeapEntry c2 = a[j1 + 1];
          if (c2.d > cj.d) {
            j = j1 + 1;
            cj = c2;
          }
        }
        if (!(cj.d > last.d)) break;
        a[i] = cj;
        i = j;
      }
      a[i] = last;
    }
    return top;
  }
};

__device__ __forceinline__ bool bit_test(const uint32_t *bits, uint32_t id) {
  return (bits[id >> 5] >> (id & 31)) & 1u;
}

// CPL = float4 columns per lane = stride / 128 when known at compile time (query held in
// registers, reduction fully unrolled); CPL == 0 is the generic path (query in shared memory).
template <int SLOTS, int METRIC, int CPL>
struct Searcher {
  static_assert(SLOTS == 2 || SLOTS == 4 || SLOTS == 8 || SLOTS == 16, "SLOTS must be 2, 4, 8 or 16");
  static constexpr int G = SLOTS / 2;  // rows per group; two groups (buffers) in flight
  const DevIndex &ix;
  const SearchArgs &a;
  SmemPtrs sm;
  uint32_t *vis;
  const int lane;
  uint32_t phase_bits;  // parity of each buffer's barrier
  CandHeap cand;        // meaningful on lane 0
  ResHeap res;          // meaningful on lane 0
  unsigned long long st_e, st_h, st_h0;
  bool overflow;
  double worst;  // lane 0: results.Peek().Distance while the result heap is non-empty (kept in a register)
  float qnorm;   // int8: query-side norm (hnsw_index.go:2405-2413)
  // fast path (search_layer_fast): the candidate / result queues as ONE sorted list in registers,
  // entry i in lane i >> 2, slot i & 3, ascending by distance, +inf beyond the live entries
  double ld[4];
  uint32_t lid[4];
  uint32_t lexp;  // bit r: entry r of this lane has been expanded (popped from the candidate queue)
  int ln;         // live entries (uniform)
  bool tie;       // two equal distances met: the heaps' tie order is needed, the exact kernel re-runs the query
  uint32_t slots_u32, bars_u32, slot_bytes, row_bytes, slot_words;
  uint64_t row_policy;  // L2 cache hint of the row copies (0 = none)  // shared-window addresses of the row slots / barriers
  const unsigned char *vec_bytes;
  // the query in registers: CPL 16-byte columns per lane; float16 columns are kept WIDENED (two float4 per column)
  static constexpr int kQRegs = CPL > 0 ? (METRIC == KIND_L2_F16 ? 2 * CPL : CPL) : 1;
  float4 qreg[kQRegs];

  // heaps_in_smem = false: the fast kernel's carve-up (no result / candidate heap arrays)
  __device__ Searcher(const DevIndex &ix_, const SearchArgs &a_, unsigned char *smem, bool heaps_in_smem = true)
      : ix(ix_), a(a_), lane(threadIdx.x & 31), phase_bits(0), st_e(0), st_h(0), st_h0(0), overflow(false), qnorm(1.f), lexp(0), ln(0), tie(false) {
    slot_words = slot_pitch_words(ix.stride, ix.row_words, METRIC);
    smem_layout(ix.stride, slot_words, heaps_in_smem ? a.ef : 0, SLOTS, heaps_in_smem ? a.cand_smem : 0u,
                ix.deg0 > ix.degu ? ix.deg0 : ix.degu, CPL == 0, METRIC, smem, &sm);
    vis = a.visited + (size_t)blockIdx.x * a.vis_words;
    cand.s = sm.cand;
    cand.g = a.cand_overflow + (size_t)blockIdx.x * a.ovf_cap;
    cand.cap_s = a.cand_smem;
    cand.cap_g = a.ovf_cap;
    cand.n = 0;
    res.a = sm.res;
    res.n = 0;
    slots_u32 = smem_u32(sm.slots);
    bars_u32 = smem_u32(sm.bars);
    slot_bytes = slot_words * (uint32_t)sizeof(float);
    row_policy = a.rows_evict_first ? l2_policy_evict_first() : 0ull;
    row_bytes = ix.row_words * (uint32_t)sizeof(float);
    vec_bytes = reinterpret_cast<const unsigned char *>(ix.vecs);
  }

  __device__ __forceinline__ void init_barriers() {
    // rows shorter than a slot (row_words < stride, float16 / int8 pitches) never overwrite the slot
    // tail: zero it once so the padded columns contribute nothing
    if (ix.row_words < slot_words) {
      for (int sl = 0; sl < SLOTS; ++sl)
        for (uint32_t w = ix.row_words + lane; w < slot_words; w += 32) sm.slots[(size_t)sl * slot_words + w] = 0.f;
      fence_proxy_async();
    }
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) mbar_init(&sm.bars[i], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }

  __device__ __forceinline__ void load_query(const float *src_row) {
    const float4 *src = reinterpret_cast<const float4 *>(src_row);
    if (CPL > 0) {
#pragma unroll
      for (int t = 0; t < (CPL > 0 ? CPL : 1); ++t) set_qreg(t, src[lane + 32 * t]);
    } else {
      for (uint32_t c = lane; c < (ix.stride >> 2); c += 32) sm.q4[c] = src[c];
    }
    __syncwarp();
  }
  __device__ __forceinline__ void set_qreg(int t, const float4 &col) {
    if (METRIC == KIND_L2_F16) {
      f16x8_widen(col, qreg[(2 * t) % kQRegs], qreg[(2 * t + 1) % kQRegs]);
    } else {
      qreg[t % kQRegs] = col;
    }
  }
  // accumulate query column t x row column `col`
  template <class Acc>
  __device__ __forceinline__ void acc_col(Acc &acc, int t, const float4 &col) const {
    if constexpr (METRIC == KIND_L2_F16) {
      acc.add_wide(qreg[(2 * t) % kQRegs], qreg[(2 * t + 1) % kQRegs], col);
    } else {
      acc.add(qreg[t % kQRegs], col);
    }
  }

  // a stored row as the query (construction: currObj := storedVector, hnsw_index.go:674, :1806): the
  // row's own bytes, zero beyond its pitch, and for int8 the query-side norm of :2405-2413 (0 -> 1)
  __device__ __forceinline__ void load_row_as_query(uint32_t id) {
    const float4 *src = reinterpret_cast<const float4 *>(ix.vecs + (size_t)id * ix.row_words);
    const uint32_t valid = ix.row_words >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (CPL > 0) {
#pragma unroll
      for (int t = 0; t < (CPL > 0 ? CPL : 1); ++t) set_qreg(t, (uint32_t)(lane + 32 * t) < valid ? src[lane + 32 * t] : z);
    } else {
      for (uint32_t c = lane; c < (ix.stride >> 2); c += 32) sm.q4[c] = c < valid ? src[c] : z;
    }
    if (METRIC == KIND_COS_I8) {
      const float nrm = ix.norms[id];
      qnorm = nrm == 0.f ? 1.f : nrm;
    }
    __syncwarp();
  }

  // ---- row streaming: groups of G rows, two group buffers, one mbarrier per buffer ------------------
  __device__ __forceinline__ void copy_row(uint32_t dst, uint32_t id, uint32_t bar) const {
    const unsigned char *src = vec_bytes + (size_t)id * row_bytes;
    if (row_policy != 0ull)
      bulk_g2s_u32_hint(dst, src, row_bytes, bar, row_policy);
    else
      bulk_g2s_u32(dst, src, row_bytes, bar);
  }
  // (both called by the whole, converged warp; one elected lane issues — see elect_one)
  __device__ __forceinline__ void issue_one(uint32_t buf, uint32_t id) {  // a single row into buffer `buf`
    if (elect_one()) {
      const uint32_t bar = bars_u32 + buf * 8u;
      mbar_expect_tx_u32(bar, row_bytes);
      copy_row(slots_u32 + buf * (uint32_t)G * slot_bytes, id, bar);
    }
  }
  // rows eval_id[j0 .. j0 + cnt) into the slots of buffer (g & 1), all completing on that buffer's barrier
  __device__ __forceinline__ void issue_group(uint32_t g, uint32_t j0, uint32_t cnt) {
    if (elect_one()) {
      const uint32_t buf = g & 1u;
      const uint32_t bar = bars_u32 + buf * 8u;
      uint32_t dst = slots_u32 + buf * (uint32_t)G * slot_bytes;
      mbar_expect_tx_u32(bar, cnt * row_bytes);
      if (cnt == (uint32_t)G) {  // a full group: straight-line code, the ids fetched with one shared-memory load
        uint32_t id[G];
        if (G == 4) {
          const uint4 v = *reinterpret_cast<const uint4 *>(sm.eval_id + j0);
          id[0] = v.x, id[G > 1 ? 1 : 0] = v.y, id[G > 2 ? 2 : 0] = v.z, id[G > 3 ? 3 : 0] = v.w;
        } else if (G == 2) {
          const uint2 v = *reinterpret_cast<const uint2 *>(sm.eval_id + j0);
          id[0] = v.x, id[G > 1 ? 1 : 0] = v.y;
        } else {
#pragma unroll
          for (int r = 0; r < G; ++r) id[r] = sm.eval_id[j0 + r];
        }
        if (row_policy != 0ull) {
#pragma unroll
          for (int r = 0; r < G; ++r)
            bulk_g2s_u32_hint(dst + r * slot_bytes, vec_bytes + (size_t)id[r] * row_bytes, row_bytes, bar, row_policy);
        } else {
#pragma unroll
          for (int r = 0; r < G; ++r) bulk_g2s_u32(dst + r * slot_bytes, vec_bytes + (size_t)id[r] * row_bytes, row_bytes, bar);
        }
      } else {
        for (uint32_t r = 0; r < cnt; ++r, dst += slot_bytes)
          copy_row(dst, sm.eval_id[j0 + r], bar);
      }
    }
  }
  __device__ __forceinline__ void wait_buf(uint32_t buf) {
    const uint32_t bar = bars_u32 + buf * 8u, parity = (phase_bits >> buf) & 1u;
    while (!mbar_try_wait_u32(bar, parity)) {
    }
    phase_bits ^= 1u << buf;
  }

  // lane partials of query x the G rows of buffer `buf`, kernel order (LaneAcc, kdb_internal.cuh); the G chains
  // are independent and interleave
  __device__ __forceinline__ void group_partials(uint32_t buf, float (&p)[G]) const {
    const float4 *r4 = reinterpret_cast<const float4 *>(sm.slots + (size_t)buf * G * slot_words);
    const uint32_t pitch4 = slot_words >> 2;
    LaneAcc<METRIC> acc[G];
    if (CPL > 0) {
#pragma unroll
      for (int t = 0; t < (CPL > 0 ? CPL : 1); ++t) {
#pragma unroll
        for (int r = 0; r < G; ++r) acc_col(acc[r], t, r4[r * pitch4 + lane + 32 * t]);
      }
    } else {
      const uint32_t nchunks = ix.stride >> 2;
#pragma unroll 2
      for (uint32_t c = lane; c < nchunks; c += 32) {
        const float4 q = sm.q4[c];
#pragma unroll
        for (int r = 0; r < G; ++r) acc[r].add(q, r4[r * pitch4 + c]);
      }
    }
#pragma unroll
    for (int r = 0; r < G; ++r) p[r] = acc[r].lane_sum();
  }
  // The xor-butterfly 16, 8, 4, 2, 1 of G rows at once.  While more than one row is left, a step at offset O pairs
  // row i with row i + R/2: the lanes with bit O clear keep row i and hand their partial of row i + R/2 to the lane
  // O away (and vice versa) — lane l still computes  own(l) + own(l ^ O)  for the row it keeps, which is exactly the
  // butterfly's addition, so the totals are bit-identical.  reduce_head = the first step (after it every lane's
  // shared-memory reads of the group have returned); reduce_tail = the rest.  Result: p[0] on lane l is the total
  // of row l >> kRowShift (int8: exact integer totals of every row in p[0..G) on every lane).
  static constexpr int kLog2G = G == 1 ? 0 : (G == 2 ? 1 : (G == 4 ? 2 : 3));
  static constexpr int kRowShift = 5 - kLog2G;
  template <int R, int O>
  __device__ __forceinline__ void reduce_step(float (&p)[G]) const {
    if (R > 1) {
      const bool up = (lane & O) != 0;
#pragma unroll
      for (int i = 0; i < (R > 1 ? R / 2 : 1); ++i) {
        const float keep = up ? p[i + R / 2] : p[i];
        const float send = up ? p[i] : p[i + R / 2];
        p[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, O));
      }
    } else {
      p[0] = __fadd_rn(p[0], __shfl_xor_sync(0xffffffffu, p[0], O));
    }
  }
  __device__ __forceinline__ void reduce_head(float (&p)[G]) const {
    if (METRIC == KIND_COS_I8) {
#pragma unroll
      for (int r = 0; r < G; ++r) p[r] = __int_as_float(__reduce_add_sync(0xffffffffu, __float_as_int(p[r])));
    } else {
      reduce_step<G, 16>(p);
    }
  }
  __device__ __forceinline__ void reduce_tail(float (&p)[G]) const {
    if (METRIC != KIND_COS_I8) {
      reduce_step<(G >= 2 ? G / 2 : 1), 8>(p);
      reduce_step<(G >= 4 ? G / 4 : 1), 4>(p);
      reduce_step<(G >= 8 ? G / 8 : 1), 2>(p);
      reduce_step<1, 1>(p);
    }
  }

  // dist(query, entry) (:2471): one row through buffer 0; the reduced value on every lane
  __device__ __forceinline__ float entry_sum(uint32_t ep) {
    issue_one(0, ep);
    wait_buf(0);
    const float4 *r4 = reinterpret_cast<const float4 *>(sm.slots);
    LaneAcc<METRIC> acc;
    if (CPL > 0) {
#pragma unroll
      for (int t = 0; t < (CPL > 0 ? CPL : 1); ++t) acc_col(acc, t, r4[lane + 32 * t]);
    } else {
      const uint32_t nchunks = ix.stride >> 2;
      for (uint32_t c = lane; c < nchunks; c += 32) acc.add(sm.q4[c], r4[c]);
    }
    const float s0 = warp_sum<METRIC>(acc.lane_sum());
    __syncwarp();  // all lanes are done reading the slot
    return s0;
  }

  // lane 0: the reference's per-neighbour result update (:2571-2591) for neighbour j at distance d
  __device__ __forceinline__ void heap_update(double d, uint32_t j, int ef) {
    HeapEntry e;
    e.d = d;
    e.id = sm.eval_id[j];
    e.pad = 0;
    bool admit = res.n < ef;  // worstDist = MaxFloat64 while results is empty
    if (!admit) admit = e.d < worst;
    if (admit) {
      if (!cand.push(e)) overflow = true;  // :2581
      if (!sm.eval_del[j]) {              // :2584
        res.push(e);
        if (res.n > ef) (void)res.pop();  // :2587-2589
        worst = res.a[0].d;
      }
    }
  }

  // Phases C + D of a hop: the rows of eval_id[0 .. n_eval) through the group pipeline.  LIST = the sorted-list
  // queues of the fast path (all lanes), otherwise the two heaps (lane 0).
  // Pre-test: once the result queue is full its worst kept distance only decreases (:2587-2589), so a row that fails
  // `d < worst` against the value at the START of its group fails the exact test (:2577) too — those rows cost
  // nothing more.  The rows that pass are handed to the exact update one by one, in row order.
  template <bool LIST>
  __device__ __forceinline__ void stream_hop(const uint32_t n_eval, const int ef) {
    if (n_eval == 0) return;
    const uint32_t n_groups = (n_eval + (uint32_t)G - 1u) / (uint32_t)G;
    issue_group(0, 0, n_eval < (uint32_t)G ? n_eval : (uint32_t)G);
    if (n_groups > 1) issue_group(1, G, n_eval - G < (uint32_t)G ? n_eval - G : (uint32_t)G);
    // the admission state every lane pre-tests against: uniform copies of lane 0's (heap path), refreshed after a
    // group that admitted something
    bool full = false;
    double wst = 0.0;
    if (!LIST) {
      full = __shfl_sync(0xffffffffu, res.n >= ef ? 1 : 0, 0) != 0;
      wst = __shfl_sync(0xffffffffu, worst, 0);
    }
    for (uint32_t g = 0, j0 = 0; g < n_groups; ++g, j0 += G) {
      const uint32_t buf = g & 1u;
      const uint32_t cnt = n_eval - j0 < (uint32_t)G ? n_eval - j0 : (uint32_t)G;
      if (LIST) {
        full = ln >= ef;
        wst = worst;
      }
      wait_buf(buf);
      float p[G];
      group_partials(buf, p);
      reduce_head(p);
      // Every lane's reads of the group's slots have returned (the step above consumed every lane's partials): the
      // slots are refilled BEFORE the rest of the reduction and the queue updates.  (No proxy fence: the slots were
      // only READ through the generic proxy; the bulk copies' completion is observed through the mbarrier.)
      __syncwarp();
      if (g + 2 < n_groups) {
        const uint32_t j2 = j0 + 2u * G;
        issue_group(g, j2, n_eval - j2 < (uint32_t)G ? n_eval - j2 : (uint32_t)G);
      }
      reduce_tail(p);
      if (METRIC == KIND_COS_I8) {
        // int8: every lane holds the exact integer dot of every row of the group.  The thresholds of
        // collect_neighbours reject with one integer compare per row; what passes takes the float64 path.
        int thr[G];
        if (G == 4) {  // (16-byte aligned: j0 is a multiple of G and the list starts on a 128-byte boundary)
          const int4 v = *reinterpret_cast<const int4 *>(sm.eval_thr + j0);
          thr[0] = v.x, thr[G > 1 ? 1 : 0] = v.y, thr[G > 2 ? 2 : 0] = v.z, thr[G > 3 ? 3 : 0] = v.w;
        } else if (G == 2) {
          const int2 v = *reinterpret_cast<const int2 *>(sm.eval_thr + j0);
          thr[0] = v.x, thr[G > 1 ? 1 : 0] = v.y;
        } else {
#pragma unroll
          for (int r = 0; r < G; ++r) thr[r] = sm.eval_thr[j0 + r];
        }
        uint32_t m = 0u;
#pragma unroll
        for (int r = 0; r < G; ++r)
          if ((uint32_t)r < cnt && __float_as_int(p[r]) > thr[r]) m |= 1u << r;
        const bool any_i8 = m != 0u;
        if (any_i8) {
#pragma unroll
          for (int r = 0; r < G; ++r) {
            if ((m >> r) & 1u) {
              const double d = int8_distance(__float_as_int(p[r]), qnorm, sm.eval_norm[j0 + r]);
              if (LIST) {
                list_update(d, j0 + r, ef);
              } else if (lane == 0) {
                heap_update(d, j0 + r, ef);
              }
            }
          }
          if (!LIST) {
            full = __shfl_sync(0xffffffffu, res.n >= ef ? 1 : 0, 0) != 0;
            wst = __shfl_sync(0xffffffffu, worst, 0);
          }
        }
        continue;
      }
      // float rows: distance + pre-test on the lane that holds the row
      const uint32_t r_mine = (uint32_t)lane >> kRowShift;
      const bool holder = (lane & ((1 << kRowShift) - 1)) == 0 && r_mine < cnt;
      double d = 0.0;
      bool pass = false;
      if (holder) {
        d = to_distance<METRIC>(p[0]);
        pass = !full || d < wst;
      }
      uint32_t mask = __ballot_sync(0xffffffffu, pass);
      const bool any = mask != 0u;
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1u;
        const double dd = __shfl_sync(0xffffffffu, d, src);
        const uint32_t j = j0 + ((uint32_t)src >> kRowShift);
        if (LIST) {
          list_update(dd, j, ef);
        } else if (lane == 0) {
          heap_update(dd, j, ef);
        }
      }
      if (!LIST && any) {
        full = __shfl_sync(0xffffffffu, res.n >= ef ? 1 : 0, 0) != 0;
        wst = __shfl_sync(0xffffffffu, worst, 0);
      }
    }
  }

  __device__ __forceinline__ void mark_logged(uint32_t id, bool log) {  // lane 0
    atomicOr(&vis[id >> 5], 1u << (id & 31));
    if (log) {
      const uint32_t m = sm.ctl->n_marked;
      if (m < (uint32_t)kMarkCap) sm.marked[m] = id;
      sm.ctl->n_marked = m + 1;
    }
  }

  // visited.Clear() (bitset.go:44-48): only this search's marks if they were logged, else all words
  __device__ __forceinline__ void clear_visited(bool logged) {
    __syncwarp();
    const uint32_t m = sm.ctl->n_marked;
    if (logged && m <= (uint32_t)kMarkCap) {
      for (uint32_t i = lane; i < m; i += 32) vis[sm.marked[i] >> 5] = 0u;
    } else {
      uint4 *v4 = reinterpret_cast<uint4 *>(vis);
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (uint32_t i = lane; i < (a.vis_words >> 2); i += 32) v4[i] = z;
    }
    __syncwarp();
  }

  // Phase B of a hop: the adjacency row of `cur` -> the ordered list of neighbours to evaluate
  // (sm.eval_id / eval_del / eval_norm / eval_thr).  Returns their number (on every lane).
  // full / worst_now (uniform): the result queue is full and its worst kept distance, as of the start of the hop.
  // int8 only: with P = qNorm * storedNorm > 0 and w = the worst kept distance,
  //   dot <= ((1 - w) - 1e-12) * P   ==>   1 - dot / P >= w + 1e-12   ==>   the reference's rounded distance (three
  // roundings, <= 1e-15 off; the clamp only ever yields 2 >= w) is >= w: not admitted (:2577).  The bound is turned
  // into an integer threshold per neighbour HERE, one lane per neighbour, so that the per-evaluation test is one
  // integer compare.  w only decreases during the hop, so a threshold from the start of the hop rejects less than it
  // could, never more: whatever passes it takes the exact float64 path.
  __device__ __forceinline__ uint32_t collect_neighbours(uint32_t cur, int level, bool log_marks, bool &expand, bool full,
                                                         double worst_now) {
    uint32_t n_eval = 0;
    // "level >= len(currentNode.Connections)" -> continue (:2521-2524)
    expand = (level == 0) || (ix.levels[cur] >= level);
    if (expand) {
      const uint32_t *row;
      uint32_t deg;
      if (level == 0) {
        deg = ix.deg0;
        row = ix.adj0 + (size_t)cur * deg;
      } else {
        deg = ix.degu;
        row = ix.upper_adj + ((size_t)ix.upper_first[cur] + (uint32_t)(level - 1)) * deg;
      }
      const double wmargin = __dsub_rn(__dsub_rn(1.0, worst_now), 1e-12);
      // 64 neighbours per pass (two per lane: slot a = idx, slot b = idx + 32, a before b in row order).  Everything
      // that depends only on the ids — nil test, int8 norm, deleted bit, allow-list bit — is loaded for both slots
      // BEFORE the two visited test-and-sets, which are issued back to back: one round trip instead of two.
      for (uint32_t base = 0; base < deg; base += 64) {  // :2537
        uint32_t idv[2];
        bool act[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t idx = base + 32u * h + lane;
          idv[h] = idx < deg ? row[idx] : 0u;  // plain load: build kernels mutate rows
        }
        act[0] = idv[0] != 0u;  // rows are compacted at upload; 0 = padding
        act[1] = idv[1] != 0u;
        if (__ballot_sync(0xffffffffu, act[0]) == 0u) break;
        const bool any_b = __ballot_sync(0xffffffffu, act[1]) != 0u;
        bool leader[2], fresh[2] = {false, false}, allowed[2] = {true, true};
        int8_t lvl[2] = {-1, -1};
        float sn[2] = {0.f, 0.f};
        uint32_t del[2] = {0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // a repeated id inside the row is visited by its first occurrence (:2539-2542)
          const uint32_t same = __match_any_sync(0xffffffffu, idv[h]);
          leader[h] = act[h] && ((__ffs(same) - 1) == lane);
          // nodes[neighborID] == nil is tested at search time (:2553-2561): a row the host has not re-patched yet
          // may still name a node Vacuum removed
          if (leader[h] && idv[h] <= ix.n) {
            lvl[h] = ix.levels[idv[h]];
            if (METRIC == KIND_COS_I8) sn[h] = ix.norms[idv[h]];
            if (ix.deleted != nullptr) del[h] = bit_test(ix.deleted, idv[h]) ? 1u : 0u;
            if (a.allow != nullptr) allowed[h] = bit_test(a.allow, idv[h]);
          }
          if (h == 0 && !any_b) break;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (leader[h]) {
            const uint32_t bit = 1u << (idv[h] & 31);
            const uint32_t old = atomicOr(&vis[idv[h] >> 5], bit);  // visited.Has + visited.Add
            fresh[h] = (old & bit) == 0u;
          }
          if (h == 0 && !any_b) break;
        }
        if (any_b) {
          // An id that occurs in BOTH slots is visited by its a-occurrence.  The two test-and-sets of such a pair
          // are not ordered: if the b-lane won, hand the visit over to the a-lane.  Only an a-lane that found its
          // bit set can be the loser of such a race, so those (few) ids are compared with the fresh b ids.
          uint32_t lost = __ballot_sync(0xffffffffu, leader[0] && !fresh[0]);
          const uint32_t won_b = __ballot_sync(0xffffffffu, fresh[1]);
          while (lost != 0u && won_b != 0u) {
            const int src = __ffs(lost) - 1;
            lost &= lost - 1u;
            const uint32_t x = __shfl_sync(0xffffffffu, idv[0], src);
            const bool hit = fresh[1] && idv[1] == x;
            if (__any_sync(0xffffffffu, hit)) {
              if (hit) fresh[1] = false;
              if (lane == src) fresh[0] = true;
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (log_marks) {
            const uint32_t fm = __ballot_sync(0xffffffffu, fresh[h]);
            const uint32_t m0 = sm.ctl->n_marked;
            if (fresh[h]) {
              const uint32_t pos = m0 + __popc(fm & ((1u << lane) - 1u));
              if (pos < (uint32_t)kMarkCap) sm.marked[pos] = idv[h];
            }
            __syncwarp();
            if (lane == 0) sm.ctl->n_marked = m0 + __popc(fm);
            __syncwarp();
          }
          // allow-list before any distance work (:2545-2549), then the nil test (:2553-2561)
          const bool keep = fresh[h] && allowed[h] && lvl[h] >= 0;
          const uint32_t km = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            const uint32_t pos = n_eval + __popc(km & ((1u << lane) - 1u));
            sm.eval_id[pos] = idv[h];
            sm.eval_del[pos] = del[h];
            if (METRIC == KIND_COS_I8) {
              sm.eval_norm[pos] = sn[h];
              int thr = (int)0x80000000;  // nothing is rejected without the exact test
              if (full && sn[h] != 0.f)
                thr = __double2int_rd(__dmul_rn(wmargin, __dmul_rn(static_cast<double>(qnorm), static_cast<double>(sn[h]))));
              sm.eval_thr[pos] = thr;
            }
          }
          n_eval += __popc(km);
          if (h == 0 && !any_b) break;
        }
      }
      __syncwarp();
    }
    return n_eval;
  }

  // searchLayerUnlocked (hnsw_index.go:2351-2611).  Returns (on every lane) the number of results
  // left in the max-heap `res` (not yet drained), or -1 if the entry node is nil (:2466-2468).
  __device__ int search_layer(const int level, const int ef, const uint32_t ep) {
    const bool log_marks = level > 0;
    if (ep == 0 || ep > ix.n || ix.levels[ep] < 0) return -1;
    if (lane == 0) sm.ctl->n_marked = 0;
    const float s0 = entry_sum(ep);  // dist(query, entry) (:2471)
    if (lane == 0) {
      cand.n = 0;
      res.n = 0;
      HeapEntry e;
      e.d = to_distance<METRIC>(s0, qnorm, METRIC == KIND_COS_I8 ? ix.norms[ep] : 0.f);
      e.id = ep;
      e.pad = 0;
      cand.push(e);                // :2478
      mark_logged(ep, log_marks);  // :2479
      bool ep_valid = true;        // :2481-2485 (an empty allow-list never reaches the kernel)
      if (a.allow != nullptr && !bit_test(a.allow, ep)) ep_valid = false;
      const bool del = ix.deleted != nullptr && bit_test(ix.deleted, ep);
      if (ep_valid && !del) {  // :2487-2489
        res.push(e);
        worst = e.d;
      }
      st_e += 1;
    }
    __syncwarp();
    for (;;) {  // :2495
      // ---- A: pop + early exit (lane 0), broadcast
      uint32_t cur = 0xffffffffu;
      if (lane == 0 && cand.n > 0 && !overflow) {
        const HeapEntry c = cand.pop();
        if (!(res.n >= ef && c.d > worst)) cur = c.id;  // :2501-2506
      }
      cur = __shfl_sync(0xffffffffu, cur, 0);
      if (cur == 0xffffffffu) break;
      // ---- B: adjacency row -> ordered list of neighbours to evaluate
      bool expand;
      bool full = false;
      double worst_now = 0.0;
      if (METRIC == KIND_COS_I8) {  // the cheap-reject thresholds need the result queue's state on every lane
        full = __shfl_sync(0xffffffffu, res.n >= ef ? 1 : 0, 0) != 0;
        worst_now = __shfl_sync(0xffffffffu, worst, 0);
      }
      const uint32_t n_eval = collect_neighbours(cur, level, log_marks, expand, full, worst_now);
      // ---- C + D: stream the rows in groups, apply the heap updates of the rows that can be admitted
      stream_hop<false>(n_eval, ef);
      if (lane == 0) {
        st_e += n_eval;
        if (expand) {
          st_h += 1;
          if (level == 0) st_h0 += 1;
        }
      }
      __syncwarp();
    }
    const int n = __shfl_sync(0xffffffffu, res.n, 0);
    __syncwarp();
    return n;
  }


  // ================================================================================================
  // Fast path.  While every distance a query meets is distinct, ANY correct priority queue pops the
  // same sequence as the reference's binary heaps (hnsw_heap.go): the minimum / maximum is unique at
  // every pop, so the expansions, admissions, evictions and the final ascending order are all forced.
  // Then the two heaps collapse into one sorted list of the ef best admitted entries with an
  // "expanded" bit: an admitted entry that is not among the ef best can never be expanded (it is
  // farther than the worst kept one, which is what stops the search, :2501-2506).  The list lives in
  // registers (4 entries per lane, ef <= 128) and is maintained by the whole warp — ~40 instructions
  // per admission instead of ~150 dependent ones on lane 0.  A tie only matters where it makes an
  // extreme ambiguous, so `tie` is set (and the query re-run by the exact heap path) exactly when
  //   * the nearest unexpanded entry and the next candidate in order are equally far (pop, :2496),
  //   * an evicted entry is as far as the new worst kept one (eviction :2587-2589; such an entry would
  //     also still pass the strict early-exit test :2501-2506 and be expanded by the reference),
  //   * two of the first k (+1) entries of the final list are equally far (output order, :2596-2610);
  // equal pairs that never reach one of those positions cannot change anything the caller can observe.
  // Not used when soft-deleted nodes exist (they are traversed but not kept, :2584) or ef > 128.
  // ================================================================================================
  __device__ __forceinline__ void sl_clear() {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      ld[r] = inf;
      lid[r] = 0u;
    }
    lexp = 0u;
    ln = 0;
  }
  __device__ __forceinline__ double sl_key(int slot) const {
    return slot == 0 ? ld[0] : (slot == 1 ? ld[1] : (slot == 2 ? ld[2] : ld[3]));
  }
  // all lanes; the caller has decided admission.  cap = ef.
  __device__ __forceinline__ void sl_insert(double xd, uint32_t xid, int cap) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) c += ld[r] < xd ? 1 : 0;
    const int p = __reduce_add_sync(0xffffffffu, c);  // keys strictly smaller than x
    const double pd = __shfl_up_sync(0xffffffffu, ld[3], 1);
    const uint32_t pid = __shfl_up_sync(0xffffffffu, lid[3], 1);
    const uint32_t pe = __shfl_up_sync(0xffffffffu, lexp >> 3, 1) & 1u;
    const int base = lane << 2;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    uint32_t ne = 0u;
#pragma unroll
    for (int r = 3; r >= 0; --r) {  // top slot first: each slot reads the still-old value below it
      const int idx = base + r;
      const double sd = r > 0 ? ld[r > 0 ? r - 1 : 0] : pd;
      const uint32_t sid = r > 0 ? lid[r > 0 ? r - 1 : 0] : pid;
      const uint32_t se = r > 0 ? ((lexp >> (r > 0 ? r - 1 : 0)) & 1u) : pe;
      const bool here = idx == p, shift = idx > p, gone = idx >= cap;
      double nd = here ? xd : (shift ? sd : ld[r]);
      uint32_t nid = here ? xid : (shift ? sid : lid[r]);
      uint32_t nb = here ? 0u : (shift ? se : ((lexp >> r) & 1u));
      if (gone) {  // what moved past the capacity is the evicted worst entry
        nd = inf;
        nid = 0u;
        nb = 0u;
      }
      ld[r] = nd;
      lid[r] = nid;
      ne |= nb << r;
    }
    lexp = ne;
    ln = ln < cap ? ln + 1 : cap;
  }
  // worst kept distance = key of entry ef - 1 (valid once ln == ef)
  __device__ __forceinline__ double sl_worst(int cap) const {
    return __shfl_sync(0xffffffffu, sl_key((cap - 1) & 3), (cap - 1) >> 2);
  }
  // nearest entry not expanded yet -> its id (and marks it), or 0xffffffff
  __device__ __forceinline__ uint32_t sl_pop_nearest() {
    const int base = lane << 2;
    uint32_t live = 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r) live |= (base + r < ln ? 1u : 0u) << r;
    const uint32_t m = live & ~lexp;
    const uint32_t b = __ballot_sync(0xffffffffu, m != 0u);
    if (b == 0u) return 0xffffffffu;
    const int L = __ffs(b) - 1;
    const int r0 = m ? __ffs(m) - 1 : 0;
    const uint32_t mine = r0 == 0 ? lid[0] : (r0 == 1 ? lid[1] : (r0 == 2 ? lid[2] : lid[3]));
    const uint32_t id = __shfl_sync(0xffffffffu, mine, L);
    // the candidate queue's minimum must be unique: the next entry in order, if it is still a
    // candidate, must be strictly farther (otherwise the heap's tie order decides who is expanded first)
    const double nd0 = __shfl_down_sync(0xffffffffu, ld[0], 1);
    const uint32_t ne0 = __shfl_down_sync(0xffffffffu, lexp, 1) & 1u;
    const double succ_d = r0 == 0 ? ld[1] : (r0 == 1 ? ld[2] : (r0 == 2 ? ld[3] : nd0));
    const uint32_t succ_exp = r0 < 3 ? ((lexp >> (r0 + 1)) & 1u) : ne0;
    const bool succ_live = base + r0 + 1 < ln;
    const bool bad = lane == L && succ_live && succ_exp == 0u && succ_d == sl_key(r0);
    if (__any_sync(0xffffffffu, bad)) tie = true;
    if (lane == L) lexp |= 1u << r0;
    return id;
  }
  // the per-neighbour result update (:2571-2591) on the sorted list; all lanes
  __device__ __forceinline__ void list_update(double d, uint32_t j, int ef) {
    if (ln < ef || d < worst) {
      const bool evicts = ln >= ef;
      const double evicted = worst;
      sl_insert(d, sm.eval_id[j], ef);
      if (ln >= ef) {
        worst = sl_worst(ef);
        // the result queue's maximum must be unique when it is evicted; an evicted entry as far as the
        // new worst one would also still be expanded by the reference (:2501-2506 is a strict test)
        if (evicts && worst == evicted) tie = true;
      }
    }
  }

  // searchLayerUnlocked on the sorted list.  Returns the number of kept entries (they stay in the
  // list, ascending), or -1 if the entry node is nil.  `tie` may be set: the result is then void.
  __device__ int search_layer_fast(const int level, const int ef, const uint32_t ep) {
    const bool log_marks = level > 0;
    if (ep == 0 || ep > ix.n || ix.levels[ep] < 0) return -1;
    if (lane == 0) sm.ctl->n_marked = 0;
    const float s0 = entry_sum(ep);
    sl_clear();
    sl_insert(to_distance<METRIC>(s0, qnorm, METRIC == KIND_COS_I8 ? ix.norms[ep] : 0.f), ep, ef);  // :2478, :2487
    if (ln >= ef) worst = sl_worst(ef);
    if (lane == 0) {
      mark_logged(ep, log_marks);  // :2479
      st_e += 1;
    }
    __syncwarp();
    for (;;) {
      const uint32_t cur = tie ? 0xffffffffu : sl_pop_nearest();  // :2496-2506
      if (cur == 0xffffffffu) break;
      bool expand;
      const uint32_t n_eval = collect_neighbours(cur, level, log_marks, expand, ln >= ef, worst);
      stream_hop<true>(n_eval, ef);
      if (lane == 0) {
        st_e += n_eval;
        if (expand) {
          st_h += 1;
          if (level == 0) st_h0 += 1;
        }
      }
      __syncwarp();
    }
    return ln;
  }

  // searchInternal on the fast path.  Returns false when a tie was met (nothing written; the exact
  // path must answer the query).  Counters are only kept for queries answered here.
  __device__ bool run_query_fast(uint32_t q) {
    const unsigned long long e0 = st_e, h0 = st_h, h00 = st_h0;
    tie = false;
    load_query(a.queries + (size_t)q * ix.stride);
    if (METRIC == KIND_COS_I8) qnorm = a.qnorms[q];
    uint32_t ep = ix.entry;
    if (a.allow != nullptr && !bit_test(a.allow, ep)) ep = a.allow_entry;
    bool failed = ix.max_level < 0;
    for (int l = ix.max_level; l > 0 && !failed && !tie; --l) {
      const int n = search_layer_fast(l, 1, ep);
      if (n > 0) ep = __shfl_sync(0xffffffffu, lid[0], 0);
      clear_visited(true);
      if (n <= 0) failed = true;
    }
    int count = 0;
    if (!failed && !tie) {
      const int n = search_layer_fast(0, a.ef, ep);
      clear_visited(false);
      if (n > 0) {  // equal distances among the first k (+1) entries: their order is the heap's
        const int base = lane << 2;
        const double nd0 = __shfl_down_sync(0xffffffffu, ld[0], 1);
        const int last_pair = (n - 2 < a.k - 1) ? n - 2 : a.k - 1;  // pairs (i, i+1), i <= last_pair
        bool bad = false;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double nxt = r < 3 ? ld[r < 3 ? r + 1 : 3] : nd0;
          bad = bad || (base + r <= last_pair && ld[r] == nxt);
        }
        if (__any_sync(0xffffffffu, bad)) tie = true;
      }
      if (n > 0 && !tie) {
        count = n < a.k ? n : a.k;
        const int base = lane << 2;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (base + r < count) {  // already ascending (:2596-2610)
            a.out_ids[(size_t)q * a.k + base + r] = lid[r] + a.id_base;
            a.out_scores[(size_t)q * a.k + base + r] = ld[r];
          }
      }
    }
    if (tie) {
      st_e = e0;
      st_h = h0;
      st_h0 = h00;
      return false;
    }
    for (int i = count + lane; i < a.k; i += 32) {
      a.out_ids[(size_t)q * a.k + i] = 0u;
      a.out_scores[(size_t)q * a.k + i] = 0.0;
    }
    if (lane == 0) a.out_counts[q] = (uint32_t)count;
    __syncwarp();
    return true;
  }

  // searchInternal (hnsw_index.go:369-468) for query q
  __device__ void run_query(uint32_t q) {
    load_query(a.queries + (size_t)q * ix.stride);
    if (METRIC == KIND_COS_I8) qnorm = a.qnorms[q];
    uint32_t ep = ix.entry;
    if (a.allow != nullptr && !bit_test(a.allow, ep)) ep = a.allow_entry;  // :436-447
    bool failed = ix.max_level < 0;
    for (int l = ix.max_level; l > 0 && !failed; --l) {  // :450-459, k = 1, efSearch = 0 -> ef = 1
      const int n = search_layer(l, 1, ep);
      if (n > 0) ep = sm.res[0].id;  // nearest[0]
      clear_visited(true);
      if (n <= 0) failed = true;  // error or "search failed at level" -> []
    }
    int count = 0;
    if (!failed) {
      const int n = search_layer(0, a.ef, ep);  // :462
      if (n > 0) {
        if (lane == 0) {  // :2596-2610 drain from the back, keep the first k
          for (int i = n - 1; i >= 0; --i) {
            const HeapEntry e = res.pop();
            if (i < a.k) {
              a.out_ids[(size_t)q * a.k + i] = e.id + a.id_base;
              a.out_scores[(size_t)q * a.k + i] = e.d;
            }
          }
        }
        count = n < a.k ? n : a.k;
      }
      clear_visited(false);
    }
    if (lane == 0) {
      if (overflow) {
        atomicExch(a.err_flag, KDBGPU_ERR_OVERFLOW);
        overflow = false;
        count = 0;
      }
      for (int i = count; i < a.k; ++i) {
        a.out_ids[(size_t)q * a.k + i] = 0u;
        a.out_scores[(size_t)q * a.k + i] = 0.0;
      }
      a.out_counts[q] = (uint32_t)count;
    }
    __syncwarp();
  }
};

}  // namespace dev
}  // namespace kdb
