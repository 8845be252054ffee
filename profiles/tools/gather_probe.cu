// gather_probe.cu — what HBM delivers for RANDOM row gathers of a given row size (the traversal's access pattern),
// as a ceiling next to the sequential-copy peak of MEASURED_PEAKS.json.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather_probe profiles/tools/gather_probe.cu
//   /tmp/gather_probe            -> one line per (row bytes, method)
// Two methods: 128-bit loads (ROWS rows in flight per warp) and 1-D bulk copies into shared memory
// (cp.async.bulk + mbarrier, 8 rows in flight per warp) — the traversal kernel's own transport.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

template <int ROWS>
__global__ void gather_ldg(const uint4 *__restrict__ base, uint32_t n_rows, uint32_t row_vec, uint32_t per_warp,
                           unsigned long long *sink) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < per_warp; i += ROWS) {
    uint4 v[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const uint32_t row = hash32(warp * 7919u + i + r) % n_rows;
      const uint4 *p = base + (size_t)row * row_vec;
      v[r] = make_uint4(0, 0, 0, 0);
      for (uint32_t c = lane; c < row_vec; c += 32) {
        const uint4 t = p[c];
        v[r].x ^= t.x ^ t.y ^ t.z ^ t.w;
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc ^= v[r].x;
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one warp per CTA (like the traversal), SL rows in flight, every lane reads its 16-byte columns of a landed row
template <int SL>
__global__ void __launch_bounds__(32) gather_bulk(const unsigned char *__restrict__ base, uint32_t n_rows,
                                                  uint32_t row_bytes, uint32_t per_warp, unsigned long long *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  unsigned char *slots = smem + 128;
  const uint32_t lane = threadIdx.x, warp = blockIdx.x;
  if (lane == 0) {
    for (int i = 0; i < SL; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](uint32_t i) {
    const uint32_t row = hash32(warp * 7919u + i) % n_rows, s = i % SL;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(row_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(slots + (size_t)s * row_bytes)),
                 "l"(base + (size_t)row * row_bytes), "r"(row_bytes), "r"(smem_u32(&bars[s]))
                 : "memory");
  };
  if (lane == 0)
    for (uint32_t i = 0; i < SL && i < per_warp; ++i) issue(i);
  uint32_t acc = 0;
  for (uint32_t i = 0; i < per_warp; ++i) {
    const uint32_t s = i % SL, parity = (i / SL) & 1u;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(smem_u32(&bars[s])), "r"(parity)
                   : "memory");
    const uint4 *r4 = reinterpret_cast<const uint4 *>(slots + (size_t)s * row_bytes);
    for (uint32_t c = lane; c < row_bytes / 16; c += 32) acc ^= r4[c].x;
    __syncwarp();
    if (lane == 0 && i + SL < per_warp) issue(i + SL);
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

int main() {
  const size_t total = (size_t)3 << 30;  // 3 GiB corpus, 24x the L2
  unsigned char *d;
  unsigned long long *sink;
  CK(cudaMalloc(&d, total));
  CK(cudaMemset(d, 1, total));
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const uint32_t sizes[] = {256, 512, 768, 1536, 3072, 6144};
  for (uint32_t rb : sizes) {
    const uint32_t n_rows = (uint32_t)(total / rb);
    const size_t want = (size_t)6 << 30;  // bytes gathered per launch
    {
      const uint32_t warps = 148 * 64, per_warp = (uint32_t)(want / rb / warps) / 8 * 8;
      float best = 1e30f;
      for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        gather_ldg<8><<<warps / 8, 256>>>(reinterpret_cast<const uint4 *>(d), n_rows, rb / 16, per_warp, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it && ms < best) best = ms;
      }
      printf("row_bytes=%u method=ldg128x8 warps=%u GB/s=%.1f\n", rb, warps, (double)warps * per_warp * rb / best / 1e6);
    }
    for (int resident = 8; resident <= 24; resident += 8) {
      const uint32_t warps = 148 * resident, per_warp = (uint32_t)(want / rb / warps);
      const size_t smem = 128 + (size_t)8 * rb;
      if (smem * resident > 220 * 1024) continue;
      CK(cudaFuncSetAttribute(gather_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float best = 1e30f;
      for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        gather_bulk<8><<<warps, 32, smem>>>(d, n_rows, rb, per_warp, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it && ms < best) best = ms;
      }
      printf("row_bytes=%u method=bulk8 warps_per_sm=%d GB/s=%.1f\n", rb, resident, (double)warps * per_warp * rb / best / 1e6);
    }
    fflush(stdout);
  }
  return 0;
}
