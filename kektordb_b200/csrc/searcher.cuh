// searcher.cuh — the per-CTA HNSW layer search shared by the query kernel (search.cu) and the
// graph-construction kernels (build.cu).  Everything here is device code in namespace kdb::dev.
//
// Stands in for searchLayerUnlocked (reference pkg/core/hnsw/hnsw_index.go:2351-2611) with the
// heaps of hnsw_heap.go:18-156 and the visited BitSet of bitset.go; see search.cu for the phase
// description.
#pragma once
#include "kdb_internal.cuh"

namespace kdb {
namespace dev {

constexpr int kMaxDeg = 256;    // max neighbours per adjacency row (2M <= 256)
constexpr int kMarkCap = 512;   // visited marks logged per upper-level search before full clear

struct Ctl {
  uint32_t cur;
  uint32_t n_eval;
  int done;
  uint32_t expanded;
  uint32_t q;
  int res_n;
  uint32_t n_marked;
  uint32_t pad;
};

struct SmemPtrs {
  float4 *q4;
  float *slots;
  uint64_t *bars;
  HeapEntry *res;
  HeapEntry *cand;
  double *eval_d;
  uint32_t *eval_id;
  uint32_t *eval_del;
  uint32_t *marked;
  Ctl *ctl;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// shared-memory carve-up, identical on host (sizing) and device (pointers)
__host__ __device__ inline size_t smem_layout(uint32_t stride, int ef, int nwarps, int slots, uint32_t cand_smem,
                                              unsigned char *base, SmemPtrs *p) {
  size_t off = 0;
  size_t o_slots = off;
  off += (size_t)nwarps * slots * stride * sizeof(float);
  size_t o_q = off;
  off += (size_t)stride * sizeof(float);
  size_t o_res = off;
  off += (size_t)(ef + 1) * sizeof(HeapEntry);
  size_t o_cand = off;
  off += (size_t)cand_smem * sizeof(HeapEntry);
  size_t o_evald = off;
  off += (size_t)kMaxDeg * sizeof(double);
  size_t o_bars = off;
  off += (size_t)nwarps * slots * sizeof(uint64_t);
  size_t o_evalid = off;
  off += (size_t)kMaxDeg * sizeof(uint32_t);
  size_t o_evaldel = off;
  off += (size_t)kMaxDeg * sizeof(uint32_t);
  size_t o_marked = off;
  off += (size_t)kMarkCap * sizeof(uint32_t);
  size_t o_ctl = off;
  off += sizeof(Ctl);
  off = align_up(off, 128);
  if (p) {
    p->slots = reinterpret_cast<float *>(base + o_slots);
    p->q4 = reinterpret_cast<float4 *>(base + o_q);
    p->res = reinterpret_cast<HeapEntry *>(base + o_res);
    p->cand = reinterpret_cast<HeapEntry *>(base + o_cand);
    p->eval_d = reinterpret_cast<double *>(base + o_evald);
    p->bars = reinterpret_cast<uint64_t *>(base + o_bars);
    p->eval_id = reinterpret_cast<uint32_t *>(base + o_evalid);
    p->eval_del = reinterpret_cast<uint32_t *>(base + o_evaldel);
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

template <int NWARPS, int SLOTS, int METRIC>
struct Searcher {
  const DevIndex &ix;
  const SearchArgs &a;
  SmemPtrs sm;
  uint32_t *vis;
  const int tid, lane, warp;
  uint32_t phase_bits;  // parity of each of this warp's slot barriers
  CandHeap cand;
  ResHeap res;
  unsigned long long st_e, st_h, st_h0;
  bool overflow;

  __device__ Searcher(const DevIndex &ix_, const SearchArgs &a_, unsigned char *smem)
      : ix(ix_), a(a_), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), phase_bits(0),
        st_e(0), st_h(0), st_h0(0), overflow(false) {
    smem_layout(ix.stride, a.ef, NWARPS, SLOTS, a.cand_smem, smem, &sm);
    vis = a.visited + (size_t)blockIdx.x * a.vis_words;
    cand.s = sm.cand;
    cand.g = a.cand_overflow + (size_t)blockIdx.x * a.ovf_cap;
    cand.cap_s = a.cand_smem;
    cand.cap_g = a.ovf_cap;
    cand.n = 0;
    res.a = sm.res;
    res.n = 0;
  }

  // Phase C: rows eval_id[0..n_eval) -> eval_d.  Warp w owns rows w, w+NWARPS, ...; each of its
  // SLOTS shared-memory slots has an mbarrier; the warp keeps SLOTS bulk copies in flight.
  __device__ __forceinline__ void gather() {
    const uint32_t n_eval = sm.ctl->n_eval;
    const uint32_t nrows = n_eval > (uint32_t)warp ? (n_eval - warp + NWARPS - 1) / NWARPS : 0;
    const uint32_t row_bytes = ix.stride * sizeof(float);
    const uint32_t nchunks = ix.stride >> 2;
    float *wslots = sm.slots + (size_t)warp * SLOTS * ix.stride;
    uint64_t *wbars = sm.bars + warp * SLOTS;
    auto issue = [&](uint32_t j) {
      if (lane == 0) {
        const uint32_t slot = j % SLOTS;
        const uint32_t id = sm.eval_id[warp + j * NWARPS];
        mbar_expect_tx(&wbars[slot], row_bytes);
        bulk_g2s(wslots + (size_t)slot * ix.stride, ix.vecs + (size_t)id * ix.stride, row_bytes, &wbars[slot]);
      }
    };
    const uint32_t pro = nrows < (uint32_t)SLOTS ? nrows : (uint32_t)SLOTS;
    for (uint32_t j = 0; j < pro; ++j) issue(j);
    for (uint32_t j = 0; j < nrows; ++j) {
      const uint32_t slot = j % SLOTS;
      mbar_wait(&wbars[slot], (phase_bits >> slot) & 1u);
      phase_bits ^= 1u << slot;
      const float s = warp_reduce_row<METRIC>(sm.q4, reinterpret_cast<const float4 *>(wslots + (size_t)slot * ix.stride),
                                              nchunks, lane);
      if (lane == 0) sm.eval_d[warp + j * NWARPS] = to_distance<METRIC>(s);
      __syncwarp();
      if (j + SLOTS < nrows) {
        fence_proxy_async();  // generic-proxy reads of the slot precede the async-proxy overwrite
        issue(j + SLOTS);
      }
    }
  }

  __device__ __forceinline__ void mark_logged(uint32_t id, bool log) {  // lane-0 only helper
    atomicOr(&vis[id >> 5], 1u << (id & 31));
    if (log) {
      uint32_t m = sm.ctl->n_marked;
      if (m < (uint32_t)kMarkCap) sm.marked[m] = id;
      sm.ctl->n_marked = m + 1;
    }
  }

  // visited.Clear() (bitset.go:44-48): only this search's marks if they were logged, else all words
  __device__ __forceinline__ void clear_visited(bool logged) {
    __syncthreads();
    const uint32_t m = sm.ctl->n_marked;
    if (logged && m <= (uint32_t)kMarkCap) {
      for (uint32_t i = tid; i < m; i += NWARPS * 32) vis[sm.marked[i] >> 5] = 0u;
    } else {
      uint4 *v4 = reinterpret_cast<uint4 *>(vis);
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (uint32_t i = tid; i < (a.vis_words >> 2); i += NWARPS * 32) v4[i] = z;
    }
    __syncthreads();
  }

  // searchLayerUnlocked (hnsw_index.go:2351-2611).  Returns the number of results left in the
  // max-heap `res` (not yet drained), or -1 if the entry node is nil (:2466-2468).
  __device__ int search_layer(const int level, const int ef, const uint32_t ep) {
    const bool log_marks = level > 0;
    if (ep == 0 || ep > ix.n || ix.levels[ep] < 0) return -1;
    if (tid == 0) {
      sm.ctl->n_eval = 1;
      sm.ctl->n_marked = 0;
      sm.eval_id[0] = ep;
    }
    __syncthreads();
    gather();  // dist(query, entry) (:2471)
    __syncthreads();
    if (tid == 0) {
      cand.n = 0;
      res.n = 0;
      HeapEntry e;
      e.d = sm.eval_d[0];
      e.id = ep;
      e.pad = 0;
      cand.push(e);              // :2478
      mark_logged(ep, log_marks);  // :2479
      bool ep_valid = true;      // :2481-2485 (an empty allow-list never reaches the kernel)
      if (a.allow != nullptr && !bit_test(a.allow, ep)) ep_valid = false;
      const bool del = ix.deleted != nullptr && bit_test(ix.deleted, ep);
      if (ep_valid && !del) res.push(e);  // :2487-2489
      st_e += 1;
    }
    for (;;) {  // :2495
      if (tid == 0) {
        int done = 0;
        if (cand.n == 0 || overflow) {
          done = 1;
        } else {
          const HeapEntry cur = cand.pop();
          if (res.n >= ef && cur.d > res.a[0].d) {  // :2501-2506
            done = 1;
          } else {
            sm.ctl->cur = cur.id;
          }
        }
        sm.ctl->done = done;
      }
      __syncthreads();
      if (sm.ctl->done) break;
      const uint32_t cur = sm.ctl->cur;
      if (warp == 0) {  // phase B
        uint32_t n_eval = 0;
        // "level >= len(currentNode.Connections)" -> continue (:2521-2524)
        const bool expand = (level == 0) || (ix.levels[cur] >= level);
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
          for (uint32_t base = 0; base < deg; base += 32) {  // :2537
            const uint32_t idx = base + lane;
            const uint32_t id = idx < deg ? row[idx] : 0u;  // plain load: build kernels mutate rows
            const bool act = id != 0u;  // rows are compacted at upload; 0 = padding
            if (__ballot_sync(0xffffffffu, act) == 0u) break;
            // a repeated id inside the row is visited by its first occurrence (:2539-2542)
            const uint32_t same = __match_any_sync(0xffffffffu, id);
            const bool leader = act && ((__ffs(same) - 1) == lane);
            bool fresh = false;
            if (leader) {
              const uint32_t bit = 1u << (id & 31);
              const uint32_t old = atomicOr(&vis[id >> 5], bit);  // visited.Has + visited.Add
              fresh = (old & bit) == 0u;
            }
            if (log_marks) {
              const uint32_t fm = __ballot_sync(0xffffffffu, fresh);
              const uint32_t m0 = sm.ctl->n_marked;
              if (fresh) {
                const uint32_t pos = m0 + __popc(fm & ((1u << lane) - 1u));
                if (pos < (uint32_t)kMarkCap) sm.marked[pos] = id;
              }
              __syncwarp();
              if (lane == 0) sm.ctl->n_marked = m0 + __popc(fm);
              __syncwarp();
            }
            // allow-list before any distance work (:2545-2549)
            const bool keep = fresh && (a.allow == nullptr || bit_test(a.allow, id));
            uint32_t del = 0u;
            if (keep && ix.deleted != nullptr) del = bit_test(ix.deleted, id) ? 1u : 0u;
            const uint32_t km = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const uint32_t pos = n_eval + __popc(km & ((1u << lane) - 1u));
              sm.eval_id[pos] = id;
              sm.eval_del[pos] = del;
            }
            n_eval += __popc(km);
          }
        }
        if (lane == 0) {
          sm.ctl->n_eval = n_eval;
          sm.ctl->expanded = expand ? 1u : 0u;
        }
      }
      __syncthreads();
      gather();  // phase C (:2566)
      __syncthreads();
      if (tid == 0) {  // phase D (:2571-2591)
        const uint32_t n_eval = sm.ctl->n_eval;
        for (uint32_t i = 0; i < n_eval; ++i) {
          HeapEntry e;
          e.d = sm.eval_d[i];
          e.id = sm.eval_id[i];
          e.pad = 0;
          bool admit = res.n < ef;  // worstDist = MaxFloat64 while results is empty
          if (!admit) admit = e.d < res.a[0].d;
          if (admit) {
            if (!cand.push(e)) overflow = true;  // :2581
            if (!sm.eval_del[i]) {              // :2584
              res.push(e);
              if (res.n > ef) (void)res.pop();  // :2587-2589
            }
          }
        }
        st_e += n_eval;
        if (sm.ctl->expanded) {
          st_h += 1;
          if (level == 0) st_h0 += 1;
        }
      }
    }
    if (tid == 0) sm.ctl->res_n = res.n;
    __syncthreads();
    return sm.ctl->res_n;
  }

  // searchInternal (hnsw_index.go:369-468) for query q
  __device__ void run_query(uint32_t q) {
    const uint32_t nchunks = ix.stride >> 2;
    const float4 *src = reinterpret_cast<const float4 *>(a.queries + (size_t)q * ix.stride);
    for (uint32_t c = tid; c < nchunks; c += NWARPS * 32) sm.q4[c] = src[c];
    __syncthreads();
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
        if (tid == 0) {  // :2596-2610 drain from the back, keep the first k
          for (int i = n - 1; i >= 0; --i) {
            const HeapEntry e = res.pop();
            if (i < a.k) {
              a.out_ids[(size_t)q * a.k + i] = e.id;
              a.out_scores[(size_t)q * a.k + i] = e.d;
            }
          }
        }
        count = n < a.k ? n : a.k;
      }
      clear_visited(false);
    }
    if (tid == 0) {
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
  }
};


}  // namespace dev
}  // namespace kdb
