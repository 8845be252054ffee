// flat_tc.cu — the tensor-core pre-filter of the flat scan (BASELINE config 3): a persistent,
// warp-specialised tcgen05 GEMM  S[q][x] = alpha * <q~, x~> + beta[x]  over bf16 copies of the
// queries and the corpus, fp32 accumulation in TMEM, operands staged by TMA (128-byte swizzle) through
// a 4-stage mbarrier ring, two accumulator stages so the epilogue of one tile overlaps the MMAs of
// the next.  It never decides a result by itself: it only nominates candidates, which flat.cu /
// rescore re-evaluates in the reference's float64 arithmetic under a certificate (see api.cu).
//
//   tile        128 queries (UMMA M, TMEM lanes) x 256 corpus rows (UMMA N, TMEM columns), K = 64 per block
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2-5   epilogue: tcgen05.ld the accumulator, thread = query, columns = corpus rows
//                 EPI_GROUPMIN min score of every 32-row group per query  (pass A -> per-query threshold)
//                 EPI_EMIT     append id where score < theta[q]           (pass B -> candidates)
//                 EPI_STORE    write every score (validation only)
//
// Replaces the O(N*D) scalar loop of BruteForceIndex.SearchWithScores
// (reference pkg/core/vector_index.go:104-162) as far as candidate nomination goes.
#include <cuda.h>
#include <cuda_bf16.h>

#include "kdb_internal.cuh"

namespace kdb {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // bf16 elements = 128 bytes = one swizzle row
constexpr int UK = 16;  // K of one tcgen05.mma.kind::f16
constexpr int STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr uint32_t B_BYTES = BN * BK * 2;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t TMEM_COLS = 512;  // two accumulator stages of BN columns

enum { EPI_STORE = 0, EPI_GROUPMIN = 1, EPI_EMIT = 2 };
constexpr int GROUP = 32;                 // corpus rows per group minimum (one tcgen05.ld.x32)
constexpr int GROUPS_PER_TILE = BN / GROUP;
constexpr int SUB = 32;                   // nominee slots per (query, CTA)
constexpr int MAX_Q_PER_LAUNCH = 2048;    // per-CTA nominee counters live in shared memory

struct TcArgs {
  uint32_t nq_tiles, n_ctiles, k_blocks;  // n_ctiles = corpus tiles this launch visits
  uint32_t ct_stride;                     // visited tile i is corpus tile i * ct_stride (pass A samples)
  uint32_t nq, n;
  uint32_t nq_rows;  // rows the per-query arrays (theta, gmin, sub_cnt) hold (pair kernel: tiles may run past them)
  float alpha;
  const float *beta;  // [n_ctiles * BN], +inf = row excluded (padding, nil, deleted, not allowed)
  float *S;           // EPI_STORE: [nq_tiles*BM][ldS]
  size_t ldS;
  float *gmin;         // EPI_GROUPMIN: [nq_tiles*BM][n_ctiles*GROUPS_PER_TILE]
  const float *theta;  // EPI_EMIT: [nq_tiles*BM]
  // EPI_EMIT: nominees of query q found by CTA c go to sub[q][c][0..SUB) without any atomic (the CTA
  // counts them in shared memory); what does not fit spills to ovf[q][0..cap) through ovf_cnt[q]
  uint2 *sub;          // [nq_pad][gridDim.x][SUB]  {id, score bits}
  uint32_t *sub_cnt;   // [nq_pad][gridDim.x]
  uint32_t *ovf_cnt;   // [nq_pad]
  uint2 *ovf;          // [nq_pad][cap]
  uint32_t cap;
  uint32_t *tile_counter;  // pair kernel, dynamic schedule: zeroed before the launch (see flat_tc2_kernel)
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc(const void *p) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 128
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Epilogue arithmetic of one 32-column chunk (thread = query): scores S = alpha * acc + beta with packed FFMA2
// (16 instructions for 32 scores), minimum of the 32 with three-input min (16 instead of 31 instructions).
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void epi_scores(const uint32_t (&r)[32], const float *sb, float alpha, float (&sc)[32]) {
  const f32x2 a2 = pack2(alpha, alpha);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {  // beta: 8 broadcast LDS.128 instead of 32 dependent LDS.32
    const float4 b4 = *reinterpret_cast<const float4 *>(sb + 4 * j4);
    const f32x2 s01 = fma2_rn(a2, pack2(__uint_as_float(r[4 * j4 + 0]), __uint_as_float(r[4 * j4 + 1])), pack2(b4.x, b4.y));
    const f32x2 s23 = fma2_rn(a2, pack2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), pack2(b4.z, b4.w));
    unpack2(s01, sc[4 * j4 + 0], sc[4 * j4 + 1]);
    unpack2(s23, sc[4 * j4 + 2], sc[4 * j4 + 3]);
  }
}
__device__ __forceinline__ float epi_min32(const float (&sc)[32]) {
  float m[11];  // 32 -> 11 -> 4 -> 2 -> 1
#pragma unroll
  for (int j = 0; j < 10; ++j) m[j] = min3f(sc[3 * j], sc[3 * j + 1], sc[3 * j + 2]);
  m[10] = fminf(sc[30], sc[31]);
  const float a0 = min3f(m[0], m[1], m[2]), a1 = min3f(m[3], m[4], m[5]), a2 = min3f(m[6], m[7], m[8]);
  return min3f(a0, a1, min3f(a2, m[9], m[10]));
}

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
    flat_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX, const TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *tiles = smem;  // [STAGES][A | B], every tile 1024-byte aligned (SWIZZLE_128B atom)
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *tmem_full = empty + STAGES;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *s_beta = reinterpret_cast<float *>(tmem_ptr + 4);  // [2][BN]
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_beta + 2 * BN);  // [nq_tiles * BM], EPI_EMIT only
  if (EPI == EPI_EMIT)
    for (uint32_t i = threadIdx.x; i < a.nq_tiles * BM; i += TC_THREADS) s_cnt[i] = 0u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t total_tiles = a.nq_tiles * a.n_ctiles;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t stage = 0, phase = 0;
      for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const uint32_t ct = (tile / a.nq_tiles) * a.ct_stride, qt = tile % a.nq_tiles;
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          unsigned char *sa = tiles + (size_t)stage * STAGE_BYTES;
          tma_load_2d(sa, &tmQ, (int)(kb * BK), (int)(qt * BM), &full[stage]);
          tma_load_2d(sa + A_BYTES, &tmX, (int)(kb * BK), (int)(ct * BN), &full[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = make_idesc();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);  // TMA bytes have landed
          tc_fence_after();
          const unsigned char *sa = tiles + (size_t)stage * STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // advancing K inside the 128-byte swizzled row: +32 bytes = +2 in the (addr >> 4) field
            umma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | (uint32_t)k) != 0u);
          }
          tc_commit(&empty[stage]);  // frees the smem stage once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        acc ^= 1u;
        if (acc == 0u) acc_phase ^= 1u;
      }
    }
  } else {  // ===== epilogue warps 2..5 =====
    const uint32_t quarter = (uint32_t)warp & 3u;  // the TMEM lane quarter this warp may read
    const uint32_t et = (uint32_t)(threadIdx.x - 64);  // 0..127
    uint32_t acc = 0, acc_phase = 0;
    // beta of a tile is fetched one tile ahead (registers), so that its global-load latency is not paid between
    // "accumulator ready" and the first tcgen05.ld of every tile
    float nb0 = 0.f, nb1 = 0.f;
    if (blockIdx.x < total_tiles) {
      const uint32_t ct0 = (blockIdx.x / a.nq_tiles) * a.ct_stride;
      nb0 = a.beta[(size_t)ct0 * BN + et];
      nb1 = a.beta[(size_t)ct0 * BN + et + 128];
    }
    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const uint32_t cti = tile / a.nq_tiles, qt = tile % a.nq_tiles;
      const uint32_t ct = cti * a.ct_stride;
      const uint32_t q = qt * BM + quarter * 32u + (uint32_t)lane;
      float *sb = s_beta + acc * BN;
      sb[et] = nb0;
      sb[et + 128] = nb1;
      if (tile + gridDim.x < total_tiles) {
        const uint32_t ctn = ((tile + gridDim.x) / a.nq_tiles) * a.ct_stride;
        nb0 = a.beta[(size_t)ctn * BN + et];
        nb1 = a.beta[(size_t)ctn * BN + et + 128];
      }
      float theta = 0.f;  // fetched before the waits: its latency hides behind them
      uint32_t my_cnt = 0u, my_cnt0 = 0u;
      if (EPI == EPI_EMIT) {
        theta = a.theta[q];
        my_cnt = my_cnt0 = s_cnt[q];  // this thread is the only one in the CTA that serves query q
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * BN;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        float sc[32];
        epi_scores(r, sb + c0, a.alpha, sc);
        if (EPI == EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t col = ct * BN + c0 + (uint32_t)j;
            if (q < a.nq && col < a.n) a.S[(size_t)q * a.ldS + col] = sc[j];
          }
        } else {
          const float best = epi_min32(sc);
          if (EPI == EPI_GROUPMIN) {
            a.gmin[(size_t)q * ((size_t)a.n_ctiles * GROUPS_PER_TILE) + (size_t)cti * GROUPS_PER_TILE + c0 / GROUP] = best;
          } else if (best < theta && q < a.nq) {  // a group holding at least one nominee of this query
            uint32_t mask = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (sc[j] < theta) ? (1u << j) : 0u;  // predicated, no branches
            while (mask) {
              const int j = __ffs((int)mask) - 1;
              mask &= mask - 1u;
              // sc[j] for a run-time j: 5-level select tree over the register-resident scores
              float v16[16], v8[8], v4[4];
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] = (j & 16) ? sc[i + 16] : sc[i];
#pragma unroll
              for (int i = 0; i < 8; ++i) v8[i] = (j & 8) ? v16[i + 8] : v16[i];
#pragma unroll
              for (int i = 0; i < 4; ++i) v4[i] = (j & 4) ? v8[i + 4] : v8[i];
              const float v2a = (j & 2) ? v4[2] : v4[0], v2b = (j & 2) ? v4[3] : v4[1];
              const float v = (j & 1) ? v2b : v2a;
              const uint2 rec = make_uint2(ct * BN + c0 + (uint32_t)j + 1u, __float_as_uint(v));  // internal id
              const uint32_t pos = my_cnt++;
              if (pos < (uint32_t)SUB) {
                a.sub[((size_t)q * gridDim.x + blockIdx.x) * SUB + pos] = rec;
              } else {
                const uint32_t p = atomicAdd(&a.ovf_cnt[q], 1u);
                if (p < a.cap) a.ovf[(size_t)q * a.cap + p] = rec;
              }
            }
          }
        }
      }
      if (EPI == EPI_EMIT && my_cnt != my_cnt0) s_cnt[q] = my_cnt;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1u;
      if (acc == 0u) acc_phase ^= 1u;
    }
    if (EPI == EPI_EMIT) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (uint32_t i = et; i < a.nq_tiles * BM; i += 128) a.sub_cnt[(size_t)i * gridDim.x + blockIdx.x] = s_cnt[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// =====================================================================================================
// The same GEMM on CTA PAIRS (tcgen05 cta_group::2): two CTAs of one cluster (same TPC) compute a 256-query x
// 256-row tile together.  Each CTA stages its own 128 query rows (A) and only HALF of the corpus tile (B, 128
// rows): the tensor core of the leader reads both halves, so the bytes per MMA that go L2 -> shared memory and
// shared memory -> tensor core drop by a third, a stage shrinks from 48 KB to 32 KB and the ring grows from 4 to
// 6 stages — 1.5x the prefetch distance in MMA time, which is what the single-CTA kernel lacked (tensor pipe 69 %
// active with neither L2 nor the crossbar above 52 %, profiles/README.md).
//   rank 0 (leader)  issues every tcgen05.mma.cta_group::2; its `full` barriers collect the bytes of BOTH CTAs'
//                    TMA loads (.cta_group::2 loads signal the leader's barrier)
//   both ranks       TMA producer for their own A rows + B half; epilogue over their own 128 TMEM lanes
//   barriers         full[s]       leader only: its own expect_tx of both CTAs' bytes
//                    empty[s]      per CTA, signalled by the leader's commit (multicast to both CTAs)
//                    tmem_full[a]  per CTA, signalled by the leader's commit (multicast)
//                    tmem_empty[a] leader only, 8 arrivals: the four epilogue warps of both CTAs (remote arrive)
constexpr int STAGES2 = 6;
constexpr uint32_t B2_BYTES = B_BYTES / 2;
constexpr uint32_t STAGE2_BYTES = A_BYTES + B2_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// (default semantics — release at CTA scope: a cluster-scope release is a full fence, ~600 cycles per arrive, and what
// the arrive orders here are tcgen05.ld's already fenced by tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *tm, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {  // arrives on this barrier in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// instruction descriptor of the pair: D = f32, A = B = bf16, both K-major, N = 256, M = 256 (128 per CTA)
__device__ __forceinline__ uint32_t make_idesc_pair() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- dynamic tile schedule of the pair kernel ---------------------------------------------------------
// With a static schedule (tile = cluster + i * clusters) every pair visits the same number of tiles, but pairs do not
// run at the same speed (ncu, profiles/r2_flat_tc2_kernel_raw.csv: sm__cycles_active min / avg / max = 1.60 / 1.71 /
// 1.82 M cycles for the nomination pass): the launch lasts as long as its slowest pair and 6-7 % of the SM time idles
// at the end.  DYN: the leader's producer lane draws the next tile from a global counter (one atomicAdd per tile, issued
// a tile ahead) and publishes it through a TQ-deep ring that exists in both CTAs of the pair:
//   s_tile[slot]     the tile index, written locally and into the peer (st.shared::cluster)
//   tq_full[slot]    per CTA, 1 arrival: the scheduler (release at cluster scope for the peer)
//   tq_empty[slot]   leader only, 10 arrivals: everyone who reads the ring — the peer's producer, the MMA lane and the
//                    four epilogue warps of both CTAs — once they hold the value
// Tiles are still handed out in order, so the pairs working at the same moment keep sharing corpus tiles in L2.
constexpr int TQ = 4;
constexpr uint32_t kNoTile = 0x7fffffffu;
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_acquire(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// "I hold ring value v": the barrier address is made to depend on v (v < 2^31, so v >> 31 is 0), which keeps the
// arrive behind the completion of the shared-memory read it announces
__device__ __forceinline__ void mbar_arrive_cluster_after(uint32_t cluster_addr, uint32_t v) {
  asm volatile(
      "{\n"
      ".reg .b32 t;\n"
      "shr.u32 t, %1, 31;\n"
      "add.u32 t, t, %0;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [t];\n"
      "}\n" ::"r"(cluster_addr),
      "r"(v)
      : "memory");
}
// One reader of the tile sequence: next_lane() from a single lane, next_warp() from a converged warp (lane 0 reads
// and signals, the value is broadcast).
template <bool DYN>
struct TileReader {
  uint32_t t, step, total;  // static schedule
  uint32_t *s_tile;
  uint64_t *tq_full;
  uint32_t empty0;  // shared::cluster address of the leader's tq_empty[0]
  uint32_t slot, ph;
  bool leader;
  __device__ __forceinline__ uint32_t next_lane() {
    if (!DYN) {
      const uint32_t r = t < total ? t : kNoTile;
      t += step;
      return r;
    }
    if (leader) mbar_wait(&tq_full[slot], ph);
    else mbar_wait_cluster_acquire(&tq_full[slot], ph);
    const uint32_t r = *reinterpret_cast<volatile uint32_t *>(&s_tile[slot]);
    mbar_arrive_cluster_after(empty0 + 8u * slot, r);
    if (++slot == (uint32_t)TQ) {
      slot = 0;
      ph ^= 1u;
    }
    return r;
  }
  __device__ __forceinline__ uint32_t next_warp(int lane) {
    if (!DYN) return next_lane();
    uint32_t r = 0u;
    if (lane == 0) r = next_lane();
    return __shfl_sync(0xffffffffu, r, 0);
  }
};

template <int EPI, bool DYN>
__global__ void __launch_bounds__(TC_THREADS, 1)
    flat_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmXh, const TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *tiles = smem;  // [STAGES2][A | B half], every tile 1024-byte aligned (SWIZZLE_128B atom)
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES2 * STAGE2_BYTES);
  uint64_t *empty = full + STAGES2;
  uint64_t *tmem_full = empty + STAGES2;
  uint64_t *tmem_empty = tmem_full + 2;
  uint64_t *tq_full = tmem_empty + 2;   // tile ring (DYN)
  uint64_t *tq_empty = tq_full + TQ;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tq_empty + TQ);
  uint32_t *s_tile = tmem_ptr + 4;                                  // [TQ]
  float *s_beta = reinterpret_cast<float *>(s_tile + TQ);           // [2][BN]
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_beta + 2 * BN);  // [nq_tiles * 2 * BM], EPI_EMIT only
  const uint32_t q_rows = a.nq_tiles * 2u * BM;                     // a.nq_tiles = 256-query tile pairs here
  if (EPI == EPI_EMIT)
    for (uint32_t i = threadIdx.x; i < q_rows; i += TC_THREADS) s_cnt[i] = 0u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0u;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXh) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES2; ++i) {
        mbar_init(&full[i], 1);  // the leader's expect_tx; the peer only contributes transaction bytes
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], 8);
      }
      for (int i = 0; i < TQ; ++i) {
        mbar_init(&tq_full[i], 1);
        mbar_init(&tq_empty[i], 10);
      }
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t total_tiles = a.nq_tiles * a.n_ctiles;
  const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  TileReader<DYN> feed;
  feed.t = cluster_id;
  feed.step = n_clusters;
  feed.total = total_tiles;
  feed.s_tile = s_tile;
  feed.tq_full = tq_full;
  feed.empty0 = mapa_u32(smem_u32(&tq_empty[0]), 0u);
  feed.slot = 0u;
  feed.ph = 0u;
  feed.leader = leader;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs): own A rows + own half of B =====
      uint32_t stage = 0, phase = 0;
      // the scheduler (DYN, leader): draws tiles from the global counter one ahead of their use and publishes them
      uint32_t drawn = 0u, pslot = 0u, pph = 0u;
      if (DYN && leader) drawn = atomicAdd(a.tile_counter, 1u);
      for (;;) {
        uint32_t tile;
        if (DYN && leader) {
          tile = drawn < total_tiles ? drawn : kNoTile;
          mbar_wait(&tq_empty[pslot], pph ^ 1u);  // every reader holds what this slot carried a ring ago
          s_tile[pslot] = tile;
          st_cluster_u32(mapa_u32(smem_u32(&s_tile[pslot]), 1u), tile);
          mbar_arrive(&tq_full[pslot]);
          mbar_arrive_cluster_release(mapa_u32(smem_u32(&tq_full[pslot]), 1u));
          if (++pslot == (uint32_t)TQ) {
            pslot = 0u;
            pph ^= 1u;
          }
          if (tile != kNoTile) drawn = atomicAdd(a.tile_counter, 1u);
        } else {
          tile = feed.next_lane();
        }
        if (tile == kNoTile) break;
        const uint32_t ct = (tile / a.nq_tiles) * a.ct_stride, qt = tile % a.nq_tiles;
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          const uint32_t leader_full = mapa_u32(smem_u32(&full[stage]), 0u);
          // the leader expects the bytes of BOTH CTAs; the peer's loads for this phase cannot be issued before the
          // previous phase was consumed (its own empty barrier), so they never land in an earlier phase
          if (leader) mbar_expect_tx(&full[stage], 2u * STAGE2_BYTES);
          unsigned char *sa = tiles + (size_t)stage * STAGE2_BYTES;
          tma_load_2d_pair(sa, &tmQ, (int)(kb * BK), (int)(qt * 2u * BM + rank * BM), leader_full);
          tma_load_2d_pair(sa + A_BYTES, &tmXh, (int)(kb * BK), (int)(ct * BN + rank * (BN / 2)), leader_full);
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {  // ===== MMA issuer: the leader CTA only =====
      const uint32_t idesc = make_idesc_pair();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      while (feed.next_lane() != kNoTile) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);  // BOTH epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (uint32_t kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);  // both CTAs' TMA bytes have landed
          tc_fence_after();
          const unsigned char *sa = tiles + (size_t)stage * STAGE2_BYTES;
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | (uint32_t)k) != 0u);
          tc_commit_pair(&empty[stage]);  // frees the stage in both CTAs once these MMAs retire
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_pair(&tmem_full[acc]);  // accumulator complete -> both epilogues
        acc ^= 1u;
        if (acc == 0u) acc_phase ^= 1u;
      }
    }
  } else {  // ===== epilogue warps 2..5 (both CTAs, each over its own 128 queries) =====
    const uint32_t quarter = (uint32_t)warp & 3u;
    const uint32_t et = (uint32_t)(threadIdx.x - 64);
    uint32_t acc = 0, acc_phase = 0;
    float nb0 = 0.f, nb1 = 0.f;  // beta one tile ahead, as in the single-CTA kernel
    uint32_t tile = feed.next_warp(lane);
    if (tile != kNoTile) {
      const uint32_t ct0 = (tile / a.nq_tiles) * a.ct_stride;
      nb0 = a.beta[(size_t)ct0 * BN + et];
      nb1 = a.beta[(size_t)ct0 * BN + et + 128];
    }
    // (DYN: the scheduler runs at least the operand ring — half a tile of MMAs — ahead of the accumulator this warp
    // is about to read, so the next tile is normally published already when it is asked for here)
    for (uint32_t tile_next; tile != kNoTile; tile = tile_next) {
      tile_next = feed.next_warp(lane);
      const uint32_t cti = tile / a.nq_tiles, qt = tile % a.nq_tiles;
      const uint32_t ct = cti * a.ct_stride;
      const uint32_t q = qt * 2u * BM + rank * BM + quarter * 32u + (uint32_t)lane;
      const bool q_ok = q < a.nq_rows;  // rows the caller's per-query arrays hold
      float *sb = s_beta + acc * BN;
      sb[et] = nb0;
      sb[et + 128] = nb1;
      if (tile_next != kNoTile) {
        const uint32_t ctn = (tile_next / a.nq_tiles) * a.ct_stride;
        nb0 = a.beta[(size_t)ctn * BN + et];
        nb1 = a.beta[(size_t)ctn * BN + et + 128];
      }
      float theta = 0.f;
      uint32_t my_cnt = 0u, my_cnt0 = 0u;
      if (EPI == EPI_EMIT) {
        theta = q_ok ? a.theta[q] : 0.f;
        my_cnt = my_cnt0 = s_cnt[q];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * BN;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        float sc[32];
        epi_scores(r, sb + c0, a.alpha, sc);
        if (EPI == EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t col = ct * BN + c0 + (uint32_t)j;
            if (q < a.nq && col < a.n) a.S[(size_t)q * a.ldS + col] = sc[j];
          }
        } else {
          const float best = epi_min32(sc);
          if (EPI == EPI_GROUPMIN) {
            if (q_ok)
              a.gmin[(size_t)q * ((size_t)a.n_ctiles * GROUPS_PER_TILE) + (size_t)cti * GROUPS_PER_TILE + c0 / GROUP] = best;
          } else if (best < theta && q < a.nq) {
            uint32_t mask = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (sc[j] < theta) ? (1u << j) : 0u;
            while (mask) {
              const int j = __ffs((int)mask) - 1;
              mask &= mask - 1u;
              float v16[16], v8[8], v4[4];
#pragma unroll
              for (int i = 0; i < 16; ++i) v16[i] = (j & 16) ? sc[i + 16] : sc[i];
#pragma unroll
              for (int i = 0; i < 8; ++i) v8[i] = (j & 8) ? v16[i + 8] : v16[i];
#pragma unroll
              for (int i = 0; i < 4; ++i) v4[i] = (j & 4) ? v8[i + 4] : v8[i];
              const float v2a = (j & 2) ? v4[2] : v4[0], v2b = (j & 2) ? v4[3] : v4[1];
              const float v = (j & 1) ? v2b : v2a;
              const uint2 rec = make_uint2(ct * BN + c0 + (uint32_t)j + 1u, __float_as_uint(v));
              const uint32_t pos = my_cnt++;
              if (pos < (uint32_t)SUB) {
                a.sub[((size_t)q * gridDim.x + blockIdx.x) * SUB + pos] = rec;
              } else {
                const uint32_t p = atomicAdd(&a.ovf_cnt[q], 1u);
                if (p < a.cap) a.ovf[(size_t)q * a.cap + p] = rec;
              }
            }
          }
        }
      }
      if (EPI == EPI_EMIT && my_cnt != my_cnt0) s_cnt[q] = my_cnt;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0u));  // the leader counts all 8 warps
      acc ^= 1u;
      if (acc == 0u) acc_phase ^= 1u;
    }
    if (EPI == EPI_EMIT) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // every query row of the launch is served by exactly one CTA of each pair: the other one reports 0 nominees
      for (uint32_t i = et; i < q_rows; i += 128)
        if (i < a.nq_rows) a.sub_cnt[(size_t)i * gridDim.x + blockIdx.x] = s_cnt[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA of the pair leaves while the other may still signal it or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// f32 rows [rows][src_stride] -> bf16 [rows_pad][dp] (zero padded); per row, rounded UP to f32:
// sumsq = sum v^2 and resid2 = sum (v - bf16(v))^2, both accumulated in f64 (they feed the error
// certificate, so they must never under-estimate)
__global__ void to_bf16_kernel(const float *__restrict__ src, size_t src_stride, uint32_t rows, uint32_t dim,
                               __nv_bfloat16 *__restrict__ dst, uint32_t dp, uint32_t rows_pad,
                               float *__restrict__ sumsq, float *__restrict__ resid2) {
  const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_pad) return;
  double acc = 0.0, racc = 0.0;
  for (uint32_t e = lane; e < dp; e += 32) {
    float v = 0.f;
    if (r < rows && e < dim) v = src[(size_t)r * src_stride + e];
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    dst[(size_t)r * dp + e] = b;
    const double dv = (double)v, dr = dv - (double)__bfloat162float(b);
    acc += dv * dv;
    racc += dr * dr;
  }
  for (int o = 16; o >= 1; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    racc += __shfl_xor_sync(0xffffffffu, racc, o);
  }
  if (lane == 0) {
    if (sumsq) sumsq[r] = __double2float_ru(acc * (1.0 + 1e-12));
    if (resid2) resid2[r] = __double2float_ru(racc * (1.0 + 1e-12));
  }
}

size_t tc_smem_bytes(uint32_t nq_pad_emit) {
  return 1024 + (size_t)STAGES * STAGE_BYTES + (2 * STAGES + 4) * sizeof(uint64_t) + 16 + 2 * BN * sizeof(float) +
         (size_t)nq_pad_emit * sizeof(uint32_t);
}
size_t tc2_smem_bytes(uint32_t q_rows_emit) {
  return 1024 + (size_t)STAGES2 * STAGE2_BYTES + (2 * STAGES2 + 4 + 2 * TQ) * sizeof(uint64_t) + 16 + TQ * sizeof(uint32_t) +
         2 * BN * sizeof(float) + (size_t)q_rows_emit * sizeof(uint32_t);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over [rows][dp] (dp contiguous), box = 64 x box_rows, 128-byte swizzle
bool make_tmap(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t dp, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {dp, rows};
  const cuuint64_t gstride[1] = {dp * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// ---- launchers (api.cu) --------------------------------------------------------------------------
uint32_t flat_tc_bm() { return BM; }
uint32_t flat_tc_bn() { return BN; }
uint32_t flat_tc_bk() { return BK; }
uint32_t flat_tc_groups_per_tile() { return GROUPS_PER_TILE; }
uint32_t flat_tc_sub_slots() { return SUB; }
uint32_t flat_tc_max_queries() { return MAX_Q_PER_LAUNCH; }

cudaError_t launch_to_bf16(const float *src, size_t src_stride, uint32_t rows, uint32_t dim, void *dst, uint32_t dp,
                           uint32_t rows_pad, float *sumsq, float *resid2, cudaStream_t stream) {
  if (rows_pad == 0) return cudaSuccess;
  const int wpb = 8;
  to_bf16_kernel<<<(rows_pad + wpb - 1) / wpb, wpb * 32, 0, stream>>>(src, src_stride, rows, dim,
                                                                      reinterpret_cast<__nv_bfloat16 *>(dst), dp,
                                                                      rows_pad, sumsq, resid2);
  return cudaGetLastError();
}

// CTA pairs (flat_tc2_kernel) whenever the batch has at least one full 256-query tile pair; KDBGPU_FLAT_2CTA=0 keeps
// the single-CTA kernel (same bits; the pair kernel measures 2-3 % faster at 1 M x 768, profiles/README.md).  L.grid is
// updated to the number of CTAs actually launched (the caller sizes / reads the per-CTA nominee lists with it).
cudaError_t launch_flat_tc(FlatTcLaunch &L, cudaStream_t stream) {
  const char *pair_env = getenv("KDBGPU_FLAT_2CTA");  // read per launch: tests switch it
  const bool pair_ok = !(pair_env && pair_env[0] == '0');
  const bool pair = pair_ok && L.nq_pad >= 2u * BM;
  CUtensorMap tmQ, tmX;
  if (!make_tmap(&tmQ, L.q_bf16, L.nq_pad, L.dp, BM) || !make_tmap(&tmX, L.x_bf16, L.n_pad, L.dp, pair ? BN / 2 : BN))
    return cudaErrorInvalidValue;
  TcArgs a;
  a.nq_tiles = pair ? (L.nq_pad + 2 * BM - 1) / (2 * BM) : L.nq_pad / BM;
  a.ct_stride = L.ct_stride ? L.ct_stride : 1u;
  a.n_ctiles = (L.n_pad / BN + a.ct_stride - 1) / a.ct_stride;
  a.k_blocks = L.dp / BK;
  a.nq = L.nq;
  a.n = L.n;
  a.nq_rows = L.nq_pad;
  a.alpha = L.alpha;
  a.beta = L.beta;
  a.S = L.S;
  a.ldS = L.ldS;
  a.gmin = L.gmin;
  a.theta = L.theta;
  a.sub = reinterpret_cast<uint2 *>(L.sub);
  a.sub_cnt = L.sub_cnt;
  a.ovf_cnt = L.ovf_cnt;
  a.ovf = reinterpret_cast<uint2 *>(L.ovf);
  a.cap = L.cap;
  const char *dyn_env = getenv("KDBGPU_FLAT_DYNAMIC");  // 0 = static tile schedule (same bits)
  a.tile_counter = (dyn_env && dyn_env[0] == '0') ? nullptr : L.tile_counter;
  if (L.epi == EPI_EMIT && (L.nq_pad > MAX_Q_PER_LAUNCH || L.grid > 256)) return cudaErrorInvalidValue;
  cudaError_t e;
  if (pair) {
    const uint32_t q_rows = a.nq_tiles * 2u * BM;
    const size_t smem = tc2_smem_bytes(L.epi == EPI_EMIT ? q_rows : 0);
    const uint64_t tiles = (uint64_t)a.nq_tiles * a.n_ctiles;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t clusters = (uint32_t)(sms / 2);
    if ((uint64_t)clusters > tiles) clusters = (uint32_t)tiles;
    if ((int)(2 * clusters) > L.grid && L.grid >= 2) clusters = (uint32_t)L.grid / 2;
    if (clusters == 0) clusters = 1;
#define KDB_TC2_LAUNCH(EPIv, DYNv)                                                                       \
  {                                                                                                      \
    auto kern = flat_tc2_kernel<EPIv, DYNv>;                                                             \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return e;                                                                      \
    cudaLaunchConfig_t cfg = {};                                                                         \
    cfg.gridDim = dim3(2 * clusters, 1, 1);                                                              \
    cfg.blockDim = dim3(TC_THREADS, 1, 1);                                                               \
    cfg.dynamicSmemBytes = smem;                                                                         \
    cfg.stream = stream;                                                                                 \
    cudaLaunchAttribute at[1];                                                                           \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                      \
    at[0].val.clusterDim.x = 2;                                                                          \
    at[0].val.clusterDim.y = 1;                                                                          \
    at[0].val.clusterDim.z = 1;                                                                          \
    cfg.attrs = at;                                                                                      \
    cfg.numAttrs = 1;                                                                                    \
    int max_clusters = 0;                                                                                \
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 &&  \
        (uint32_t)max_clusters < clusters) {                                                             \
      clusters = (uint32_t)max_clusters;                                                                 \
      cfg.gridDim = dim3(2 * clusters, 1, 1);                                                            \
    }                                                                                                    \
    (void)cudaGetLastError();                                                                            \
    e = cudaLaunchKernelEx(&cfg, kern, tmQ, tmX, a);                                                     \
  }
    // the caller zeroes *tile_counter on this stream before the launch
    if (L.epi == EPI_STORE) KDB_TC2_LAUNCH(EPI_STORE, false)
    else if (L.epi == EPI_GROUPMIN && a.tile_counter) KDB_TC2_LAUNCH(EPI_GROUPMIN, true)
    else if (L.epi == EPI_GROUPMIN) KDB_TC2_LAUNCH(EPI_GROUPMIN, false)
    else if (a.tile_counter) KDB_TC2_LAUNCH(EPI_EMIT, true)
    else KDB_TC2_LAUNCH(EPI_EMIT, false)
#undef KDB_TC2_LAUNCH
    L.grid = (int)(2 * clusters);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
  }
  const size_t smem = tc_smem_bytes(L.epi == EPI_EMIT ? L.nq_pad : 0);
#define KDB_TC_LAUNCH(EPIv)                                                                              \
  {                                                                                                      \
    auto kern = flat_tc_kernel<EPIv>;                                                                    \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return e;                                                                      \
    kern<<<L.grid, TC_THREADS, smem, stream>>>(tmQ, tmX, a);                                             \
  }
  if (L.epi == EPI_STORE) KDB_TC_LAUNCH(EPI_STORE)
  else if (L.epi == EPI_GROUPMIN) KDB_TC_LAUNCH(EPI_GROUPMIN)
  else KDB_TC_LAUNCH(EPI_EMIT)
#undef KDB_TC_LAUNCH
  return cudaGetLastError();
}

// ---- pass plumbing: beta vector, per-query thresholds, exact re-score + certificate ---------------
namespace {

constexpr float kInf = __builtin_huge_valf();

// beta[r] for corpus row r (id r+1): sumsq (L2) or 0 (cosine); +inf when the row must not be nominated
__global__ void tc_beta_kernel(const DevIndex ix, const float *__restrict__ sumsq, const uint32_t *__restrict__ allow,
                               int use_norm, uint32_t n_pad, float *__restrict__ beta) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_pad) return;
  const uint32_t id = r + 1;
  bool ok = r < ix.n && ix.levels[id] >= 0;
  if (ok && ix.deleted) ok = !((ix.deleted[id >> 5] >> (id & 31)) & 1u);
  if (ok && allow) ok = (allow[id >> 5] >> (id & 31)) & 1u;
  beta[r] = ok ? (use_norm ? sumsq[r] : 0.f) : kInf;
}

// out[0] = max sumsq, out[1] = max resid2 over rows < n (non-negative floats order like their bits)
__global__ void tc_max_kernel(const float *__restrict__ sumsq, const float *__restrict__ resid2, uint32_t n,
                              float *__restrict__ out) {
  float m0 = 0.f, m1 = 0.f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    m0 = fmaxf(m0, sumsq[i]);
    m1 = fmaxf(m1, resid2[i]);
  }
  for (int o = 16; o >= 1; o >>= 1) {
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(m0));
    atomicMax(reinterpret_cast<unsigned int *>(out) + 1, __float_as_uint(m1));
  }
}

__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k >> 31) ? (k & 0x7fffffffu) : ~k); }

constexpr int TH_THREADS = 256;
constexpr int TH_BINS = 2048;
// One CTA per query.  tau = k-th smallest group minimum (radix select, 11+11+10 bits): k distinct
// groups hold a row scoring <= tau, so at least k rows score <= tau.  With e[q] bounding
// |approximate - true| score for every row of the corpus and `slack` the float64 path's own
// rounding, theta = tau + 2(e + slack) makes every row left out of {score < theta} provably farther
// than the k-th nominated row.  bound[q] = e + slack is handed to the certificate.
__global__ void __launch_bounds__(TH_THREADS)
    tc_threshold_kernel(const float *__restrict__ gmin, uint32_t n_groups, uint32_t nq, int k,
                        const float *__restrict__ qsumsq, const float *__restrict__ qresid2,
                        const float *__restrict__ xmax, float alpha, int use_norm, uint32_t dp,
                        float *__restrict__ theta, float *__restrict__ bound) {
  __shared__ uint32_t hist[TH_BINS];
  __shared__ uint32_t part[TH_THREADS];
  __shared__ uint32_t s_prefix, s_rem;
  const uint32_t q = blockIdx.x;
  if (q >= nq) return;
  const float *row = gmin + (size_t)q * n_groups;
  const int t = threadIdx.x;
  if (t == 0) {
    s_prefix = 0u;
    s_rem = (uint32_t)k;
  }
  __syncthreads();
  bool enough = true;
  int shift = 32;
  for (int pass = 0; pass < 3 && enough; ++pass) {
    const int bits = pass < 2 ? 11 : 10;
    shift -= bits;
    for (int i = t; i < TH_BINS; i += TH_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const int hi = shift + bits;
    for (uint32_t i = t; i < n_groups; i += TH_THREADS) {
      const uint32_t key = fkey(row[i]);
      if (hi >= 32 || (key >> hi) == (prefix >> hi)) atomicAdd(&hist[(key >> shift) & ((1u << bits) - 1u)], 1u);
    }
    __syncthreads();
    const int per = TH_BINS / TH_THREADS;  // 8 bins per thread
    uint32_t mine = 0;
    for (int j = 0; j < per; ++j) mine += hist[t * per + j];
    part[t] = mine;
    __syncthreads();
    if (t == 0) {
      uint32_t rem = s_rem, c = 0;
      int chunk = 0;
      for (; chunk < TH_THREADS; ++chunk) {
        if (c + part[chunk] >= rem) break;
        c += part[chunk];
      }
      if (chunk == TH_THREADS) {
        s_rem = 0xffffffffu;  // fewer than k values in total
      } else {
        rem -= c;
        int b = chunk * per;
        for (;; ++b) {
          if (hist[b] >= rem) break;
          rem -= hist[b];
        }
        s_prefix = prefix | ((uint32_t)b << shift);
        s_rem = rem;
      }
    }
    __syncthreads();
    if (s_rem == 0xffffffffu) enough = false;
  }
  if (t == 0) {
    const float up = 1.000001f;
    const float qn = sqrtf(qsumsq[q]) * up, qr = sqrtf(qresid2[q]) * up;
    const float xn = sqrtf(xmax[0]) * up, xr = sqrtf(xmax[1]) * up;
    // |<q~,x~>_tc - <q,x>| <= |q - q~||x| + |q~||x - x~|  (bf16 rounding of the operands)
    //                         + dp * 2^-21 |q~||x~|        (fp32 accumulation inside the tensor core)
    const float dot_err = qr * xn + (qn + qr) * xr + (float)dp * 4.76837158203125e-7f * (qn + qr) * (xn + xr);
    const float e = fabsf(alpha) * dot_err * 1.0001f + 1e-6f * (fabsf(alpha) * qn * xn + (use_norm ? xmax[0] : 0.f)) + 1e-12f;
    const float slack = 4.76837158203125e-7f * (qn + xn) * (qn + xn) + 1e-9f;  // f64 path vs real arithmetic, |q|^2 in f32
    const float tau = enough ? fkey_inv(s_prefix) : kInf;
    theta[q] = tau < kInf ? tau + 2.f * (e + slack) * 1.0001f : kInf;
    bound[q] = e + slack;
  }
}

constexpr int RF_THREADS = 256;
// One CTA per query.  Pass B nominated every row scoring below the (loose, sample-based) theta and
// kept its approximate score.  tau' = the k-th smallest of those scores is the k-th smallest
// approximate score of the WHOLE corpus (every row below theta is in the buffer, and at least k are),
// so only rows scoring below theta' = tau' + 2 * bound can still enter the exact top k: they go on to
// the float64 re-score; the rest are dropped under the same certificate.
// flags[q]: 0 ok, 1 nomination buffer overflow, 2 survivor buffer overflow, 4 fewer than k nominees
// although rows were withheld.
__global__ void __launch_bounds__(RF_THREADS)
    tc_refine_kernel(const uint2 *__restrict__ sub, const uint32_t *__restrict__ sub_cnt, uint32_t grid,
                     const uint32_t *__restrict__ ovf_cnt, const uint2 *__restrict__ ovf, uint32_t cap, uint32_t nq, int k,
                     const float *__restrict__ theta, const float *__restrict__ bound, uint32_t *__restrict__ fcnt,
                     uint32_t *__restrict__ fid, uint32_t fcap, float *__restrict__ theta_final,
                     uint32_t *__restrict__ flags) {
  __shared__ uint32_t hist[TH_BINS];
  __shared__ uint32_t part[RF_THREADS];
  __shared__ uint32_t s_sub[256];
  __shared__ uint32_t s_prefix, s_rem, s_out, s_total;
  const uint32_t q = blockIdx.x;
  if (q >= nq) return;
  const int t = threadIdx.x;
  const float th = theta[q];
  const uint32_t n_ovf = ovf_cnt[q];
  if (n_ovf > cap) {  // the spill buffer lost nominees: nothing can be certified
    if (t == 0) {
      flags[q] = 1u;
      fcnt[q] = 0u;
      theta_final[q] = th;
    }
    return;
  }
  if (t == 0) s_total = 0u;
  __syncthreads();
  for (uint32_t i = t; i < grid; i += RF_THREADS) {
    uint32_t v = sub_cnt[(size_t)q * grid + i];
    if (v > (uint32_t)SUB) v = SUB;
    s_sub[i] = v;
    atomicAdd(&s_total, v);
  }
  __syncthreads();
  const uint32_t c = s_total + n_ovf;
  const uint2 *qsub = sub + (size_t)q * grid * SUB;
  const uint2 *qovf = ovf + (size_t)q * cap;
  const uint32_t n_slots = grid * SUB + n_ovf;
  // record i of this query's nominee space, or id 0 for an empty slot
  auto rec_at = [&](uint32_t i) -> uint2 {
    if (i < grid * SUB) {
      if ((i % SUB) < s_sub[i / SUB]) return qsub[i];
      return make_uint2(0u, 0u);
    }
    return qovf[i - grid * SUB];
  };
  float thf = th;  // theta = +inf (nothing withheld) or fewer than k nominees: keep them all
  if (c >= (uint32_t)k && th < kInf) {
    if (t == 0) {
      s_prefix = 0u;
      s_rem = (uint32_t)k;
    }
    __syncthreads();
    int shift = 32;
    for (int pass = 0; pass < 3; ++pass) {
      const int bits = pass < 2 ? 11 : 10;
      shift -= bits;
      for (int i = t; i < TH_BINS; i += RF_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const int hi = shift + bits;
      for (uint32_t i = t; i < n_slots; i += RF_THREADS) {
        const uint2 r = rec_at(i);
        if (r.x == 0u) continue;
        const uint32_t key = fkey(__uint_as_float(r.y));
        if (hi >= 32 || (key >> hi) == (prefix >> hi)) atomicAdd(&hist[(key >> shift) & ((1u << bits) - 1u)], 1u);
      }
      __syncthreads();
      const int per = TH_BINS / RF_THREADS;
      uint32_t mine = 0;
      for (int j = 0; j < per; ++j) mine += hist[t * per + j];
      part[t] = mine;
      __syncthreads();
      if (t == 0) {
        uint32_t rem = s_rem, acc = 0;
        int chunk = 0;
        for (; chunk < RF_THREADS - 1; ++chunk) {
          if (acc + part[chunk] >= rem) break;
          acc += part[chunk];
        }
        rem -= acc;
        int b = chunk * per;
        for (; b < TH_BINS - 1; ++b) {
          if (hist[b] >= rem) break;
          rem -= hist[b];
        }
        s_prefix = prefix | ((uint32_t)b << shift);
        s_rem = rem;
      }
      __syncthreads();
    }
    const float tau = fkey_inv(s_prefix);
    const float tf = tau + 2.f * bound[q] * 1.0001f;
    thf = tf < th ? tf : th;
  }
  if (t == 0) s_out = 0u;
  __syncthreads();
  for (uint32_t i = t; i < n_slots; i += RF_THREADS) {
    const uint2 r = rec_at(i);
    if (r.x != 0u && __uint_as_float(r.y) < thf) {
      const uint32_t pos = atomicAdd(&s_out, 1u);
      if (pos < fcap) fid[(size_t)q * fcap + pos] = r.x;
    }
  }
  __syncthreads();
  if (t == 0) {
    const uint32_t n_out = s_out;
    uint32_t flag = 0u;
    if (n_out > fcap) flag = 2u;
    else if (n_out < (uint32_t)k && th < kInf) flag = 4u;
    fcnt[q] = n_out > fcap ? fcap : n_out;
    theta_final[q] = thf;
    flags[q] = flag;
  }
}

// one term of the flat scan's float64 sum, exactly as flat.cu's flat_distances_kernel forms it
template <int MODE, int METRIC>
__device__ __forceinline__ double exact_term(float qf, float xf) {
  if (MODE == 0) {
    const double diff = static_cast<double>(__fsub_rn(qf, xf));
    return __dmul_rn(diff, diff);
  } else if (METRIC == KDBGPU_METRIC_COSINE) {
    return __dmul_rn(static_cast<double>(qf), static_cast<double>(xf));
  } else {
    const double diff = __dsub_rn(static_cast<double>(qf), static_cast<double>(xf));
    return __dmul_rn(diff, diff);
  }
}

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
// floats of every row per pipeline step (template parameter RS_CHUNK): 64 = 256 contiguous bytes per request and
// 139 KB of row tiles per CTA (one CTA per SM); 32 halves the tiles so that two CTAs share an SM and one's waits hide
// behind the other's sums.  Tile rows are padded by 4 floats (conflict-free LDS.128 with lane = row).
constexpr int RS_CHUNK_DEFAULT = 32;

// One CTA per query: exact distances of the surviving rows in the reference's arithmetic (flat.cu's
// flat_distances_kernel order: sequential in the element index, float64), ascending (distance, id),
// first k out, and the certificate
//   theta' - bound  >  (k-th exact distance) - qconst
// i.e. every row that was dropped is provably farther than the k-th result.  flags[q] != 0 sends the
// query to the exhaustive float64 scan.
template <int MODE, int METRIC, int RS_CHUNK>
__global__ void __launch_bounds__(RS_THREADS)
    tc_rescore_kernel(const DevIndex ix, const float *__restrict__ queries, size_t q_stride, uint32_t nq, int k,
                      const uint32_t *__restrict__ fcnt, const uint32_t *__restrict__ fid, uint32_t fcap,
                      const float *__restrict__ theta_final, const float *__restrict__ bound,
                      const float *__restrict__ qsumsq, uint32_t *__restrict__ out_ids, double *__restrict__ out_scores,
                      uint32_t *__restrict__ out_counts, uint32_t *__restrict__ flags,
                      unsigned long long *__restrict__ n_rescored) {
  constexpr int RS_LD = RS_CHUNK + 4;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const uint32_t q = blockIdx.x;
  if (q >= nq) return;
  uint32_t cap2 = 1;
  while (cap2 < fcap) cap2 <<= 1;
  double *s_d = reinterpret_cast<double *>(rs_smem);
  uint32_t *s_id = reinterpret_cast<uint32_t *>(s_d + cap2);
  float *s_q = reinterpret_cast<float *>(s_id + cap2);
  float *s_tile = s_q + ((ix.dim + RS_CHUNK - 1) / RS_CHUNK) * RS_CHUNK;  // [RS_WARPS][2][32][RS_LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t c = flags[q] ? 0u : fcnt[q];
  uint32_t c2 = 1;  // sort only what is there
  while (c2 < c) c2 <<= 1;
  const uint32_t dim_pad = ((ix.dim + RS_CHUNK - 1) / RS_CHUNK) * RS_CHUNK;
  for (uint32_t e = tid; e < dim_pad; e += RS_THREADS) s_q[e] = e < ix.dim ? queries[(size_t)q * q_stride + e] : 0.f;
  for (uint32_t i = tid; i < c2; i += RS_THREADS) {
    s_id[i] = i < c ? fid[(size_t)q * fcap + i] : 0xffffffffu;
    s_d[i] = __longlong_as_double(0x7ff0000000000000LL);
  }
  __syncthreads();
  float *tile = s_tile + (size_t)warp * 2 * 32 * RS_LD;  // two stages per warp
  const uint32_t n_chunks = (ix.dim + RS_CHUNK - 1) / RS_CHUNK;
  for (uint32_t base = warp * 32; base < c; base += RS_WARPS * 32) {
    const uint32_t mine = base + lane < c ? s_id[base + lane] : 0u;
    // rows are stored with stride >= dim rounded up to 128 floats, zero padded (kdb_internal.cuh), so a
    // whole 512-byte chunk of every row can be requested; lane l moves bytes [16l, 16l+16) of each row
    auto issue = [&](uint32_t ci) {
      constexpr int LPR = RS_CHUNK / 4;  // lanes (16 bytes each) per row chunk
      constexpr int RPI = 32 / LPR;      // rows covered by one warp-wide request
      float *dst = tile + (size_t)(ci & 1u) * 32 * RS_LD + (lane / LPR) * RS_LD + 4 * (lane % LPR);
#pragma unroll 8
      for (int r = 0; r < 32; r += RPI) {
        const uint32_t id = __shfl_sync(0xffffffffu, mine, r + lane / LPR);
        if (id != 0u) {
          const float *src = ix.vecs + (size_t)id * ix.stride + (size_t)ci * RS_CHUNK + 4 * (lane % LPR);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + r * RS_LD)), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc = 0.0;
    issue(0);
    for (uint32_t ci = 0; ci < n_chunks; ++ci) {
      if (ci + 1 < n_chunks) {
        issue(ci + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      const uint32_t e0 = ci * RS_CHUNK;
      const int emax = (ix.dim - e0) < (uint32_t)RS_CHUNK ? (int)(ix.dim - e0) : RS_CHUNK;
      const float *trow = tile + (size_t)(ci & 1u) * 32 * RS_LD + lane * RS_LD;
      if (mine != 0u) {
        int e = 0;
        // 16 elements at a time: the products do not depend on the running sum, so they are formed
        // first (independent) and only the 16 additions stay on the loop-carried chain — the order of
        // the additions, and therefore every rounding, is still the reference's
        for (; e + 16 <= emax; e += 16) {
          double term[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 x4 = *reinterpret_cast<const float4 *>(trow + e + 4 * g);
            const float4 q4 = *reinterpret_cast<const float4 *>(s_q + e0 + e + 4 * g);
            term[4 * g + 0] = exact_term<MODE, METRIC>(q4.x, x4.x);
            term[4 * g + 1] = exact_term<MODE, METRIC>(q4.y, x4.y);
            term[4 * g + 2] = exact_term<MODE, METRIC>(q4.z, x4.z);
            term[4 * g + 3] = exact_term<MODE, METRIC>(q4.w, x4.w);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) acc = __dadd_rn(acc, term[j]);
        }
        for (; e < emax; ++e) acc = __dadd_rn(acc, exact_term<MODE, METRIC>(s_q[e0 + e], trow[e]));
      }
      __syncwarp();
    }
    if (MODE == 1 && METRIC == KDBGPU_METRIC_COSINE) acc = __dsub_rn(1.0, acc);
    if (base + lane < c) s_d[base + lane] = acc;
  }
  __syncthreads();
  if (tid == 0) atomicAdd(n_rescored, (unsigned long long)c);
  // ascending (distance, id)
  for (uint32_t size = 2; size <= c2; size <<= 1)
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < c2; i += RS_THREADS) {
        const uint32_t j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const double di = s_d[i], dj = s_d[j];
          const uint32_t ii = s_id[i], ij = s_id[j];
          const bool gt = di > dj || (di == dj && ii > ij);
          if (gt == up) {
            s_d[i] = dj;
            s_d[j] = di;
            s_id[i] = ij;
            s_id[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  const uint32_t kk = c < (uint32_t)k ? c : (uint32_t)k;
  for (int i = tid; i < k; i += RS_THREADS) {
    const bool ok = (uint32_t)i < kk;
    out_ids[(size_t)q * k + i] = ok ? s_id[i] : 0u;
    out_scores[(size_t)q * k + i] = ok ? s_d[i] : 0.0;
  }
  if (tid == 0) {
    out_counts[q] = kk;
    uint32_t flag = flags[q];
    const float th = theta_final[q];
    if (!flag && th < kInf) {
      // exact distance -> score units of the pre-filter: L2  d = score + |q|^2 ; cosine d = 1 + score
      double qconst;
      if (MODE == 0 || METRIC == KDBGPU_METRIC_L2) qconst = (double)qsumsq[q];
      else qconst = 1.0;
      if (kk < (uint32_t)k) {
        flag = 3u;
      } else {
        const double kth_score = s_d[kk - 1] - qconst;
        if (!((double)th - (double)bound[q] > kth_score)) flag = 3u;
      }
    }
    flags[q] = flag;
  }
}

}  // namespace

cudaError_t launch_tc_beta(const DevIndex &ix, const float *sumsq, const uint32_t *allow, int use_norm, uint32_t n_pad,
                           float *beta, cudaStream_t stream) {
  tc_beta_kernel<<<(n_pad + 255) / 256, 256, 0, stream>>>(ix, sumsq, allow, use_norm, n_pad, beta);
  return cudaGetLastError();
}

cudaError_t launch_tc_max(const float *sumsq, const float *resid2, uint32_t n, float *out2, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  if (n) tc_max_kernel<<<296, 256, 0, stream>>>(sumsq, resid2, n, out2);
  return cudaGetLastError();
}

cudaError_t launch_tc_threshold(const float *gmin, uint32_t n_groups, uint32_t nq, int k, const float *qsumsq,
                                const float *qresid2, const float *xmax, float alpha, int use_norm, uint32_t dp,
                                float *theta, float *bound, cudaStream_t stream) {
  tc_threshold_kernel<<<nq, TH_THREADS, 0, stream>>>(gmin, n_groups, nq, k, qsumsq, qresid2, xmax, alpha, use_norm, dp,
                                                    theta, bound);
  return cudaGetLastError();
}

cudaError_t launch_tc_refine(const void *sub, const uint32_t *sub_cnt, uint32_t grid, const uint32_t *ovf_cnt,
                             const void *ovf, uint32_t cap, uint32_t nq, int k, const float *theta, const float *bound,
                             uint32_t *fcnt, uint32_t *fid, uint32_t fcap, float *theta_final, uint32_t *flags,
                             cudaStream_t stream) {
  if (grid > 256) return cudaErrorInvalidValue;
  tc_refine_kernel<<<nq, RF_THREADS, 0, stream>>>(reinterpret_cast<const uint2 *>(sub), sub_cnt, grid, ovf_cnt,
                                                 reinterpret_cast<const uint2 *>(ovf), cap, nq, k, theta, bound, fcnt,
                                                 fid, fcap, theta_final, flags);
  return cudaGetLastError();
}

namespace {
int rescore_chunk() {  // KDBGPU_FLAT_RS_CHUNK = 16 | 32 | 64 (measurement switch, read per launch; same bits)
  const char *v = getenv("KDBGPU_FLAT_RS_CHUNK");
  const int c = v ? atoi(v) : 0;
  return (c == 16 || c == 32 || c == 64) ? c : RS_CHUNK_DEFAULT;
}
size_t rescore_smem(uint32_t dim, uint32_t fcap, int chunk) {
  uint32_t cap2 = 1;
  while (cap2 < fcap) cap2 <<= 1;
  return (size_t)cap2 * (sizeof(double) + sizeof(uint32_t)) + (size_t)((dim + chunk - 1) / chunk) * chunk * sizeof(float) +
         (size_t)RS_WARPS * 2 * 32 * (chunk + 4) * sizeof(float);
}
}  // namespace

size_t tc_rescore_smem(uint32_t dim, uint32_t fcap) { return rescore_smem(dim, fcap, 64); }  // the largest shape

cudaError_t launch_tc_rescore(const DevIndex &ix, int mode, const float *queries, size_t q_stride, uint32_t nq, int k,
                              const uint32_t *fcnt, const uint32_t *fid, uint32_t fcap, const float *theta_final,
                              const float *bound, const float *qsumsq, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, uint32_t *flags, unsigned long long *n_rescored,
                              cudaStream_t stream) {
  const int chunk = rescore_chunk();
  const size_t smem = rescore_smem(ix.dim, fcap, chunk);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e;
#define KDB_RS_LAUNCH_C(MODEv, METRICv, CHUNKv)                                                                    \
  {                                                                                                                \
    auto kern = tc_rescore_kernel<MODEv, METRICv, CHUNKv>;                                                         \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    if (e != cudaSuccess) return e;                                                                                \
    kern<<<nq, RS_THREADS, smem, stream>>>(ix, queries, q_stride, nq, k, fcnt, fid, fcap, theta_final, bound,      \
                                           qsumsq, out_ids, out_scores, out_counts, flags, n_rescored);            \
  }
#define KDB_RS_LAUNCH(MODEv, METRICv)                                                                              \
  {                                                                                                                \
    if (chunk == 16) KDB_RS_LAUNCH_C(MODEv, METRICv, 16)                                                           \
    else if (chunk == 32) KDB_RS_LAUNCH_C(MODEv, METRICv, 32)                                                      \
    else KDB_RS_LAUNCH_C(MODEv, METRICv, 64)                                                                       \
  }
  if (mode == 0) KDB_RS_LAUNCH(0, KDBGPU_METRIC_L2)
  else if (ix.metric == KDBGPU_METRIC_COSINE) KDB_RS_LAUNCH(1, KDBGPU_METRIC_COSINE)
  else KDB_RS_LAUNCH(1, KDBGPU_METRIC_L2)
#undef KDB_RS_LAUNCH_C
#undef KDB_RS_LAUNCH
  return cudaGetLastError();
}

}  // namespace kdb
