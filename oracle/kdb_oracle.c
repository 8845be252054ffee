/*
 * kdb_oracle.c — CPU ORACLE (test infrastructure, never shipped; see kdb_oracle.h).
 *
 * A from-scratch C restatement of the reference's HNSW search / insert / flat scan
 * semantics.  Each function names the reference file:line (relative to /root/reference)
 * whose behaviour it follows.  Build: oracle/Makefile (gcc -O3 -ffp-contract=off).
 *
 * -ffp-contract=off is load-bearing: every fused multiply-add below is an explicit
 * fmaf(); every other a*b+c is two roundings, as in Go on amd64 and in Rust.
 */
#include "kdb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define KDBO_HAVE_AVX2 1
#endif

/* ======================================================================================
 * 1. Distances
 * ==================================================================================== */

/* pure-Go loops: squaredEuclideanDistanceGo distance_go.go:57-68, dotProductGo :81-90 */
static float l2_seq(const float *a, const float *b, size_t n) {
  float sum = 0.0f;
  for (size_t i = 0; i < n; i++) {
    float diff = a[i] - b[i];
    sum += diff * diff;
  }
  return sum;
}
static float dot_seq(const float *a, const float *b, size_t n) {
  float sum = 0.0f;
  for (size_t i = 0; i < n; i++) sum += a[i] * b[i];
  return sum;
}

/* Rust AVX2/FMA kernels, native/compute/src/lib.rs:22-99: one 8-lane accumulator updated by
 * vfmadd, horizontal sum in the order of reduce_sum_ps (:22-31), then a scalar tail with
 * separate multiply and add (:53-69, :87-97). */
static inline float hsum8_rust(const float acc[8]) {
  float s0 = acc[0] + acc[4], s1 = acc[1] + acc[5], s2 = acc[2] + acc[6], s3 = acc[3] + acc[7];
  float t0 = s0 + s2, t1 = s1 + s3; /* movehl + add */
  return t0 + t1;                   /* shuffle(1) + add_ss */
}
static float l2_avx2_generic(const float *a, const float *b, size_t n) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; j++) {
      float d = a[i + j] - b[i + j];
      acc[j] = fmaf(d, d, acc[j]);
    }
  float total = hsum8_rust(acc);
  for (; i < n; i++) {
    float d = a[i] - b[i];
    total += d * d;
  }
  return total;
}
static float dot_avx2_generic(const float *a, const float *b, size_t n) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; j++) acc[j] = fmaf(a[i + j], b[i + j], acc[j]);
  float total = hsum8_rust(acc);
  for (; i < n; i++) total += a[i] * b[i];
  return total;
}

/* Kernel order: 128 independent f32 accumulators, element e feeds accumulator
 * e mod 128 by FMA in increasing e (lane l of the warp owns accumulators 4l..4l+3 = one float4
 * column).  Lane partial = (a0+a1)+(a2+a3); then the xor-butterfly 16,8,4,2,1 whose lane-0 value
 * is x[l] += x[l+w] for w = 16,8,4,2,1. */
static inline float kernel_tree(const float acc[128]) {
  float lane[32];
  for (int l = 0; l < 32; l++)
    lane[l] = (acc[4 * l] + acc[4 * l + 1]) + (acc[4 * l + 2] + acc[4 * l + 3]);
  for (int w = 16; w >= 1; w >>= 1)
    for (int l = 0; l < w; l++) lane[l] = lane[l] + lane[l + w];
  return lane[0];
}
static float dot_kernel_generic(const float *a, const float *b, size_t n) {
  float acc[128];
  memset(acc, 0, sizeof acc);
  for (size_t e = 0; e < n; e++) acc[e & 127] = fmaf(a[e], b[e], acc[e & 127]);
  return kernel_tree(acc);
}
static float l2_kernel_generic(const float *a, const float *b, size_t n) {
  float acc[128];
  memset(acc, 0, sizeof acc);
  for (size_t e = 0; e < n; e++) {
    float d = a[e] - b[e];
    acc[e & 127] = fmaf(d, d, acc[e & 127]);
  }
  return kernel_tree(acc);
}

#ifdef KDBO_HAVE_AVX2
static float l2_avx2_fast(const float *a, const float *b, size_t n) {
  __m256 s = _mm256_setzero_ps();
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    __m256 d = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
    s = _mm256_fmadd_ps(d, d, s);
  }
  float acc[8];
  _mm256_storeu_ps(acc, s);
  float total = hsum8_rust(acc);
  for (; i < n; i++) {
    float d = a[i] - b[i];
    total += d * d;
  }
  return total;
}
static float dot_avx2_fast(const float *a, const float *b, size_t n) {
  __m256 s = _mm256_setzero_ps();
  size_t i = 0;
  for (; i + 8 <= n; i += 8) s = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s);
  float acc[8];
  _mm256_storeu_ps(acc, s);
  float total = hsum8_rust(acc);
  for (; i < n; i++) total += a[i] * b[i];
  return total;
}
/* kernel order with register-resident accumulators: two passes of 8 ymm (64 accumulators each) */
static float kernel_fast(const float *a, const float *b, size_t n, int l2) {
  float acc[128];
  size_t full = n & ~(size_t)127;
  for (int half = 0; half < 2; half++) {
    __m256 s[8];
    for (int j = 0; j < 8; j++) s[j] = _mm256_setzero_ps();
    for (size_t t = 0; t < full; t += 128) {
      const float *pa = a + t + 64 * half, *pb = b + t + 64 * half;
      for (int j = 0; j < 8; j++) {
        __m256 x = _mm256_loadu_ps(pa + 8 * j), y = _mm256_loadu_ps(pb + 8 * j);
        if (l2) {
          __m256 d = _mm256_sub_ps(x, y);
          s[j] = _mm256_fmadd_ps(d, d, s[j]);
        } else {
          s[j] = _mm256_fmadd_ps(x, y, s[j]);
        }
      }
    }
    for (int j = 0; j < 8; j++) _mm256_storeu_ps(acc + 64 * half + 8 * j, s[j]);
  }
  for (size_t e = full; e < n; e++) {
    if (l2) {
      float d = a[e] - b[e];
      acc[e & 127] = fmaf(d, d, acc[e & 127]);
    } else {
      acc[e & 127] = fmaf(a[e], b[e], acc[e & 127]);
    }
  }
  return kernel_tree(acc);
}
#endif

static float sq_euclid(int arith, const float *a, const float *b, size_t n, int generic) {
  switch (arith) {
    case KDBO_ARITH_AVX2:
#ifdef KDBO_HAVE_AVX2
      if (!generic) return l2_avx2_fast(a, b, n);
#endif
      return l2_avx2_generic(a, b, n);
    case KDBO_ARITH_KERNEL:
#ifdef KDBO_HAVE_AVX2
      if (!generic) return kernel_fast(a, b, n, 1);
#endif
      return l2_kernel_generic(a, b, n);
    default:
      return l2_seq(a, b, n);
  }
}
static float dot_f32(int arith, const float *a, const float *b, size_t n, int generic) {
  switch (arith) {
    case KDBO_ARITH_AVX2:
#ifdef KDBO_HAVE_AVX2
      if (!generic) return dot_avx2_fast(a, b, n);
#endif
      return dot_avx2_generic(a, b, n);
    case KDBO_ARITH_KERNEL:
#ifdef KDBO_HAVE_AVX2
      if (!generic) return kernel_fast(a, b, n, 0);
#endif
      return dot_kernel_generic(a, b, n);
    default:
      return dot_seq(a, b, n);
  }
}
float kdbo_sq_euclid_f32(int arith, const float *a, const float *b, size_t n) {
  return sq_euclid(arith, a, b, n, 0);
}
float kdbo_dot_f32(int arith, const float *a, const float *b, size_t n) { return dot_f32(arith, a, b, n, 0); }

/* DistanceFuncF32: float64(sum) for L2 (distance_go.go:67), 1.0 - float64(dot) for cosine (:127) */
static inline double dist_fn(int metric, int arith, const float *a, const float *b, size_t n, int generic) {
  if (metric == KDBO_METRIC_COSINE) return 1.0 - (double)dot_f32(arith, a, b, n, generic);
  return (double)sq_euclid(arith, a, b, n, generic);
}
double kdbo_distance(int metric, int arith, const float *a, const float *b, size_t n) {
  return dist_fn(metric, arith, a, b, n, 0);
}
double kdbo_distance_generic(int metric, int arith, const float *a, const float *b, size_t n) {
  return dist_fn(metric, arith, a, b, n, 1);
}

/* ---- float16 / int8 precisions ----------------------------------------
 * float16.Fromfloat32(v).Bits() (hnsw_index.go:427-430, :501-505, :1562-1566): the conversion lives
 * in github.com/x448/float16 v0.8.4 (go.mod:19, not vendored) and is the IEEE 754 binary32 ->
 * binary16 round-to-nearest-even conversion (subnormals kept, overflow to infinity); restated here
 * and pinned against numpy's float16 cast and the F16C instruction in tests/test_oracle_quantized.py. */
uint16_t kdbo_f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t e = (x >> 23) & 0xffu;
  uint32_t m = x & 0x7fffffu;
  if (e == 0xffu) return (uint16_t)(sign | (m ? 0x7e00u : 0x7c00u));
  const int32_t exp = (int32_t)e - 127 + 15;
  if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    m |= 0x800000u;
    const uint32_t shift = (uint32_t)(14 - exp);
    uint32_t half = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half & 1u))) half++;
    return (uint16_t)(sign | half);
  }
  uint32_t half = ((uint32_t)exp << 10) | (m >> 13);
  const uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++; /* a carry may reach 0x7c00 = +inf */
  return (uint16_t)(sign | half);
}
/* float16.Frombits(b).Float32(): exact */
float kdbo_f16_to_f32(uint16_t h) {
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, x;
  if (e == 0) {
    if (m == 0) {
      x = sign;
    } else { /* subnormal: renormalise */
      int sh = 0;
      while (!(m & 0x400u)) {
        m <<= 1;
        sh++;
      }
      m &= 0x3ffu;
      x = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
    }
  } else if (e == 31) {
    x = sign | 0x7f800000u | (m << 13);
  } else {
    x = sign | ((e - 15 + 127) << 23) | (m << 13);
  }
  float f;
  memcpy(&f, &x, 4);
  return f;
}

/* squaredEuclideanGoFloat16 (distance_go.go:93-105): f32 difference of the widened halves,
 * sequential f32 sum, separate multiply and add. */
static float l2_f16_seq(const uint16_t *a, const uint16_t *b, size_t n) {
  float sum = 0.0f;
  for (size_t i = 0; i < n; i++) {
    float diff = kdbo_f16_to_f32(a[i]) - kdbo_f16_to_f32(b[i]);
    sum += diff * diff;
  }
  return sum;
}
/* Rust squared_euclidean_f16_fma (native/compute/src/lib.rs:101-141): 8-lane FMA accumulator,
 * reduce_sum_ps, scalar remainder with separate multiply and add. */
static float l2_f16_avx2(const uint16_t *a, const uint16_t *b, size_t n) {
  size_t i = 0;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#if defined(KDBO_HAVE_AVX2) && defined(__F16C__)
  __m256 s = _mm256_setzero_ps();
  for (; i + 8 <= n; i += 8) {
    __m256 x = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(a + i)));
    __m256 y = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(b + i)));
    __m256 d = _mm256_sub_ps(x, y);
    s = _mm256_fmadd_ps(d, d, s);
  }
  _mm256_storeu_ps(acc, s);
#else
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; j++) {
      float d = kdbo_f16_to_f32(a[i + j]) - kdbo_f16_to_f32(b[i + j]);
      acc[j] = fmaf(d, d, acc[j]);
    }
#endif
  float total = hsum8_rust(acc);
  for (; i < n; i++) {
    float d = kdbo_f16_to_f32(a[i]) - kdbo_f16_to_f32(b[i]);
    total += d * d;
  }
  return total;
}
/* Kernel order for 16-bit rows: a lane's 16-byte column holds 8 elements, so there
 * are 256 independent f32 accumulators; element e feeds accumulator e mod 256 by FMA in increasing
 * e; lane l owns accumulators 8l..8l+7 and sums them ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)); then the
 * same xor-butterfly as the f32 rows. */
static float l2_f16_kernel(const uint16_t *a, const uint16_t *b, size_t n) {
  float acc[256];
  memset(acc, 0, sizeof acc);
  size_t e = 0;
#if defined(KDBO_HAVE_AVX2) && defined(__F16C__)
  const size_t full = n & ~(size_t)255;
  if (full) {
    __m256 s[32];
    for (int j = 0; j < 32; j++) s[j] = _mm256_setzero_ps();
    for (size_t t = 0; t < full; t += 256)
      for (int j = 0; j < 32; j++) {
        __m256 x = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(a + t + 8 * j)));
        __m256 y = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(b + t + 8 * j)));
        __m256 d = _mm256_sub_ps(x, y);
        s[j] = _mm256_fmadd_ps(d, d, s[j]);
      }
    for (int j = 0; j < 32; j++) _mm256_storeu_ps(acc + 8 * j, s[j]);
    e = full;
  }
#endif
  for (; e < n; e++) {
    float d = kdbo_f16_to_f32(a[e]) - kdbo_f16_to_f32(b[e]);
    acc[e & 255] = fmaf(d, d, acc[e & 255]);
  }
  float lane[32];
  for (int l = 0; l < 32; l++) {
    const float *p = acc + 8 * l;
    lane[l] = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
  }
  for (int w = 16; w >= 1; w >>= 1)
    for (int l = 0; l < w; l++) lane[l] = lane[l] + lane[l + w];
  return lane[0];
}
float kdbo_sq_euclid_f16(int arith, const uint16_t *a, const uint16_t *b, size_t n) {
  switch (arith) {
    case KDBO_ARITH_AVX2: return l2_f16_avx2(a, b, n);
    case KDBO_ARITH_KERNEL: return l2_f16_kernel(a, b, n);
    default: return l2_f16_seq(a, b, n);
  }
}
/* dotProductGoInt8 (distance_go.go:108-118): exact int32 sum of int32 products — integer
 * arithmetic, so every summation order gives the same bits.  (The reference's optional Rust AVX2
 * kernel drops two of its four 32-bit partial sums in its horizontal reduction,
 * native/compute/src/lib.rs:165-170 — `(hi64 + lo64) as i32`; the default pure-Go build is what
 * this follows.) */
int32_t kdbo_dot_i8(const int8_t *a, const int8_t *b, size_t n) {
  int32_t sum = 0;
  for (size_t i = 0; i < n; i++) sum += (int32_t)a[i] * (int32_t)b[i];
  return sum;
}
/* computeInt8Norm, hnsw_index.go:3371-3377 */
float kdbo_int8_norm(const int8_t *v, size_t n) {
  int64_t sum = 0;
  for (size_t i = 0; i < n; i++) sum += (int64_t)v[i] * (int64_t)v[i];
  return (float)sqrt((double)sum);
}
/* int8 cosine distance from the integer dot and the two f32 norms
 * (searchLayerUnlocked distFn hnsw_index.go:2421-2449, distanceBetweenNodes :319-336) */
double kdbo_int8_cosine_distance(int32_t dot, float qnorm, float stored_norm) {
  if (stored_norm == 0.0f) return 1.0;
  double sim = (double)dot / ((double)qnorm * (double)stored_norm);
  if (sim > 1.0) sim = 1.0;
  if (sim < -1.0) sim = -1.0;
  return 1.0 - sim;
}
/* Quantizer.Quantize, pkg/core/distance/quantizer.go:135-160 */
void kdbo_quantize(float abs_max, const float *v, int8_t *out, size_t n) {
  if (abs_max == 0.0f) {
    memset(out, 0, n);
    return;
  }
  for (size_t i = 0; i < n; i++) {
    float scaled = (v[i] / abs_max) * 127.0f;
    if (scaled > 127.0f)
      scaled = 127.0f;
    else if (scaled < -127.0f)
      scaled = -127.0f;
    out[i] = (int8_t)round((double)scaled); /* math.Round: half away from zero */
  }
}
/* Quantizer.Train, quantizer.go:49-125: stride sample above 10 000 vectors (10 %, capped at
 * 25 000, floor 10 000), 99.9th percentile of |value|.  Returns AbsMax (0 = nothing to train on). */
static int f32_cmp(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}
float kdbo_train_quantizer(const float *vecs, size_t n, size_t dim) {
  if (n == 0 || dim == 0) return 0.0f;
  size_t step = 1, count = n;
  if (n > 10000) {
    size_t target = n / 10;
    if (target > 25000) target = 25000;
    if (target < 10000) target = 10000;
    step = n / target;
    if (step < 1) step = 1;
    count = 0;
    for (size_t i = 0; i < n; i += step) {
      count++;
      if (count >= target) break;
    }
  }
  float *all = (float *)malloc(count * dim * sizeof(float));
  size_t w = 0, taken = 0;
  for (size_t i = 0; i < n && taken < count; i += step, taken++)
    for (size_t j = 0; j < dim; j++) all[w++] = fabsf(vecs[i * dim + j]);
  qsort(all, w, sizeof(float), f32_cmp);
  long long qi = (long long)((double)w * 0.999);
  if (qi >= (long long)w) qi = (long long)w - 1;
  if (qi < 0) qi = 0;
  float r = all[qi];
  free(all);
  return r;
}

/* normalize / invSqrt, hnsw_index.go:3030-3045: f32 sequential sum of squares, one f64 sqrt,
 * f32 reciprocal, f32 scale; a zero vector is left untouched. */
void kdbo_normalize(float *v, size_t n) {
  float norm_sq = 0.0f;
  for (size_t i = 0; i < n; i++) norm_sq += v[i] * v[i];
  if (norm_sq > 0.0f) {
    float inv = 1.0f / (float)sqrt((double)norm_sq);
    for (size_t i = 0; i < n; i++) v[i] *= inv;
  }
}

/* randomLevel, hnsw_index.go:2616-2625 */
int kdbo_random_level(double u, int m, int current_max) {
  double ml = 1.0 / log((double)m);
  double f = floor(-log(u) * ml);
  int level = (f > 1e9 || f != f) ? 1000000000 : (int)f;
  if (level > current_max + 1) return current_max + 1;
  return level;
}

double kdbo_score_from_distance(double d) { return 1.0 / (1.0 + d); } /* search_utils.go:48-52 */

/* needsRefine boost, hnsw_index.go:387-399 */
int kdbo_effective_ef(int ef_search, int needs_refine) {
  int actual = ef_search;
  if (needs_refine) {
    int boosted = (int)((double)ef_search * 2);
    if (boosted < 80) boosted = 80;
    if (boosted > 200) boosted = 200;
    if (boosted > actual) actual = boosted;
  }
  return actual;
}

/* ======================================================================================
 * 2. Heaps (hnsw_heap.go:18-156) — value semantics, strict comparisons
 * ==================================================================================== */
typedef struct {
  uint32_t id;
  double d;
} cand;
typedef struct {
  cand *a;
  size_t n, cap;
} heap;

static void heap_reserve(heap *h, size_t need) {
  if (need > h->cap) {
    size_t c = h->cap ? h->cap * 2 : 256;
    while (c < need) c *= 2;
    h->a = (cand *)realloc(h->a, c * sizeof(cand));
    h->cap = c;
  }
}
#define SWAP(x, y) \
  do {             \
    cand _t = (x); \
    (x) = (y);     \
    (y) = _t;      \
  } while (0)

static void min_push(heap *h, cand x) { /* Push + up, hnsw_heap.go:33-36, :53-63 */
  heap_reserve(h, h->n + 1);
  h->a[h->n++] = x;
  size_t j = h->n - 1;
  for (;;) {
    size_t i = (j == 0) ? 0 : (j - 1) / 2;
    if (i == j || !(h->a[j].d < h->a[i].d)) break;
    SWAP(h->a[i], h->a[j]);
    j = i;
  }
}
static cand min_pop(heap *h) { /* Pop + down, hnsw_heap.go:39-51, :65-83 */
  cand x = h->a[0];
  h->a[0] = h->a[h->n - 1];
  h->n--;
  size_t n = h->n, i = 0;
  for (;;) {
    size_t j1 = 2 * i + 1;
    if (j1 >= n) break;
    size_t j = j1, j2 = j1 + 1;
    if (j2 < n && h->a[j2].d < h->a[j1].d) j = j2;
    if (!(h->a[j].d < h->a[i].d)) break;
    SWAP(h->a[i], h->a[j]);
    i = j;
  }
  return x;
}
static void max_push(heap *h, cand x) { /* hnsw_heap.go:105-108, :122-132 */
  heap_reserve(h, h->n + 1);
  h->a[h->n++] = x;
  size_t j = h->n - 1;
  for (;;) {
    size_t i = (j == 0) ? 0 : (j - 1) / 2;
    if (i == j || !(h->a[j].d > h->a[i].d)) break;
    SWAP(h->a[i], h->a[j]);
    j = i;
  }
}
static cand max_pop(heap *h) { /* hnsw_heap.go:110-120, :134-151 */
  cand x = h->a[0];
  h->a[0] = h->a[h->n - 1];
  h->n--;
  size_t n = h->n, i = 0;
  for (;;) {
    size_t j1 = 2 * i + 1;
    if (j1 >= n) break;
    size_t j = j1, j2 = j1 + 1;
    if (j2 < n && h->a[j2].d > h->a[j1].d) j = j2;
    if (!(h->a[j].d > h->a[i].d)) break;
    SWAP(h->a[i], h->a[j]);
    i = j;
  }
  return x;
}

void kdbo_heap_roundtrip(int kind, const uint32_t *ids, const double *d, size_t n, uint32_t *out_ids,
                         double *out_d) {
  heap h = {0, 0, 0};
  for (size_t i = 0; i < n; i++) {
    cand c = {ids[i], d[i]};
    if (kind == 0)
      min_push(&h, c);
    else
      max_push(&h, c);
  }
  for (size_t i = 0; i < n; i++) {
    cand c = kind == 0 ? min_pop(&h) : max_pop(&h);
    out_ids[i] = c.id;
    out_d[i] = c.d;
  }
  free(h.a);
}

/* ======================================================================================
 * 3. Index
 * ==================================================================================== */
struct kdbo_index {
  int dim, metric, m, mmax0, efc, arith;
  size_t stride;       /* floats per stored row (dim rounded up to 16 → 64-byte rows) */
  uint32_t cap;        /* highest id that fits                                           */
  uint32_t counter;    /* nodeCounter: last assigned id (ids from 1, hnsw_index.go:590) */
  uint32_t entry;      /* entrypointID                                                   */
  int max_level;       /* -1 = empty                                                     */
  float *vecs;         /* [(cap+1)][stride]                                              */
  int8_t *level;       /* len(Connections)-1; -1 = nil node                              */
  uint8_t *deleted;    /* Node.Deleted                                                   */
  uint32_t *l0;        /* [(cap+1)][mmax0+1]: slot 0 = count                             */
  uint32_t **upper;    /* per node: [level][m+1], slot 0 = count                         */
  volatile char *lock; /* per-node spin lock (stands in for the 128 shard RWMutexes)      */
  volatile char meta_lock;
  /* float16 / int8 precisions: rows in stored form, int8 norms, quantizer */
  int precision;    /* KDBO_PREC_* */
  size_t qstride;   /* elements per quantized row (dim rounded up to 32) */
  uint16_t *rows16; /* [(cap+1)][qstride] float16 bits */
  int8_t *rows8;    /* [(cap+1)][qstride] */
  float *norms;     /* quantizedNorms: computeInt8Norm of every stored row */
  float abs_max;    /* Quantizer.AbsMax */
};

static inline void spin_lock(volatile char *l) {
  while (__atomic_exchange_n(l, 1, __ATOMIC_ACQUIRE)) {
    while (__atomic_load_n(l, __ATOMIC_RELAXED)) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
}
static inline void spin_unlock(volatile char *l) { __atomic_store_n(l, 0, __ATOMIC_RELEASE); }

kdbo_index *kdbo_new(int dim, int metric, int m, int efc, int arith, uint32_t capacity) {
  return kdbo_new_ex(dim, metric, KDBO_PREC_F32, m, efc, arith, capacity);
}
kdbo_index *kdbo_new_ex(int dim, int metric, int precision, int m, int efc, int arith, uint32_t capacity) {
  if (dim <= 0 || capacity == 0) return NULL;
  /* float16Funcs holds Euclidean only, int8Funcs Cosine only (distance_go.go:139-146) */
  if (precision == KDBO_PREC_F16 && metric != KDBO_METRIC_L2) return NULL;
  if (precision == KDBO_PREC_I8 && metric != KDBO_METRIC_COSINE) return NULL;
  if (precision < KDBO_PREC_F32 || precision > KDBO_PREC_I8) return NULL;
  kdbo_index *h = (kdbo_index *)calloc(1, sizeof *h);
  h->precision = precision;
  if (m <= 0) m = 16;       /* hnsw_index.go:140-142 */
  if (efc <= 0) efc = 200;  /* :143-145 */
  h->dim = dim;
  h->metric = metric;
  h->m = m;
  h->mmax0 = m * 2; /* :149 */
  h->efc = efc;
  h->arith = arith;
  h->stride = ((size_t)dim + 15) & ~(size_t)15;
  h->cap = capacity;
  h->max_level = -1;
  size_t n1 = (size_t)capacity + 1;
  h->qstride = ((size_t)dim + 31) & ~(size_t)31;
  if (precision == KDBO_PREC_F32) {
    if (posix_memalign((void **)&h->vecs, 64, n1 * h->stride * sizeof(float))) {
      free(h);
      return NULL;
    }
    memset(h->vecs, 0, n1 * h->stride * sizeof(float));
  } else if (precision == KDBO_PREC_F16) {
    h->rows16 = (uint16_t *)calloc(n1 * h->qstride, sizeof(uint16_t));
  } else {
    h->rows8 = (int8_t *)calloc(n1 * h->qstride, 1);
    h->norms = (float *)calloc(n1, sizeof(float));
  }
  h->level = (int8_t *)malloc(n1);
  memset(h->level, -1, n1);
  h->deleted = (uint8_t *)calloc(n1, 1);
  h->l0 = (uint32_t *)calloc(n1 * (size_t)(h->mmax0 + 1), sizeof(uint32_t));
  h->upper = (uint32_t **)calloc(n1, sizeof(uint32_t *));
  h->lock = (volatile char *)calloc(n1, 1);
  return h;
}
void kdbo_free(kdbo_index *h) {
  if (!h) return;
  for (size_t i = 0; i <= h->cap; i++) free(h->upper[i]);
  free(h->upper);
  free(h->vecs);
  free(h->rows16);
  free(h->rows8);
  free(h->norms);
  free(h->level);
  free(h->deleted);
  free(h->l0);
  free((void *)h->lock);
  free(h);
}
void kdbo_set_arith(kdbo_index *h, int arith) { h->arith = arith; }
uint32_t kdbo_count(const kdbo_index *h) { return h->counter; }
uint32_t kdbo_entry(const kdbo_index *h) { return h->entry; }
int kdbo_max_level(const kdbo_index *h) { return h->max_level; }
int kdbo_dim(const kdbo_index *h) { return h->dim; }
int kdbo_m(const kdbo_index *h) { return h->m; }
size_t kdbo_row_stride(const kdbo_index *h) { return h->stride; }
const float *kdbo_vector(const kdbo_index *h, uint32_t id) { return h->vecs + (size_t)id * h->stride; }

static inline uint32_t *conn_row(const kdbo_index *h, uint32_t id, int level) {
  if (level == 0) return h->l0 + (size_t)id * (size_t)(h->mmax0 + 1);
  return h->upper[id] + (size_t)(level - 1) * (size_t)(h->m + 1);
}
int kdbo_precision(const kdbo_index *h) { return h->precision; }
float kdbo_abs_max(const kdbo_index *h) { return h->abs_max; }
void kdbo_set_quantizer(kdbo_index *h, float abs_max) { h->abs_max = abs_max; }
const float *kdbo_norms(const kdbo_index *h) { return h->norms; }
/* stored row in the index precision (node.GetVectorF32 / F16 / I8, hnsw_node.go:13-39) */
static inline const void *row_ptr(const kdbo_index *h, uint32_t id) {
  switch (h->precision) {
    case KDBO_PREC_F16: return h->rows16 + (size_t)id * h->qstride;
    case KDBO_PREC_I8: return h->rows8 + (size_t)id * h->qstride;
    default: return h->vecs + (size_t)id * h->stride;
  }
}
const void *kdbo_row_raw(const kdbo_index *h, uint32_t id) { return row_ptr(h, id); }
size_t kdbo_row_raw_stride(const kdbo_index *h) { return h->precision == KDBO_PREC_F32 ? h->stride : h->qstride; }
/* the query-side int8 norm of searchLayerUnlocked (hnsw_index.go:2405-2413): 0 -> 1 */
static inline float query_norm_i8(const int8_t *q, size_t n) {
  float qn = kdbo_int8_norm(q, n);
  return qn == 0.0f ? 1.0f : qn;
}
/* distFn(node) of searchLayerUnlocked (hnsw_index.go:2388-2454): q is the prepared query in the
 * index precision, qnorm its int8 norm */
static inline double query_dist(const kdbo_index *h, const void *q, float qnorm, uint32_t id) {
  const size_t dim = (size_t)h->dim;
  switch (h->precision) {
    case KDBO_PREC_F16:
      return (double)kdbo_sq_euclid_f16(h->arith, (const uint16_t *)q, (const uint16_t *)row_ptr(h, id), dim);
    case KDBO_PREC_I8:
      return kdbo_int8_cosine_distance(kdbo_dot_i8((const int8_t *)q, (const int8_t *)row_ptr(h, id), dim), qnorm,
                                       h->norms[id]);
    default:
      return dist_fn(h->metric, h->arith, (const float *)q, kdbo_vector(h, id), dim, 0);
  }
}
/* distanceBetweenNodes, hnsw_index.go:297-341 */
static inline double node_dist(const kdbo_index *h, uint32_t a, uint32_t b) {
  if (h->precision == KDBO_PREC_I8) {
    if (h->norms[a] == 0.0f || h->norms[b] == 0.0f) return 1.0; /* :325-327 */
    return query_dist(h, row_ptr(h, a), h->norms[a], b);
  }
  return query_dist(h, row_ptr(h, a), 0.0f, b);
}
/* converts one incoming f32 vector into the stored form of row `id`
 * (Add :484-520, AddBatch :1553-1577; only cosine + float32 is normalised) */
static void store_row(kdbo_index *h, uint32_t id, const float *vec) {
  const size_t dim = (size_t)h->dim;
  if (h->precision == KDBO_PREC_F16) {
    uint16_t *dst = h->rows16 + (size_t)id * h->qstride;
    memset(dst, 0, h->qstride * sizeof(uint16_t));
    for (size_t i = 0; i < dim; i++) dst[i] = kdbo_f32_to_f16(vec[i]);
  } else if (h->precision == KDBO_PREC_I8) {
    int8_t *dst = h->rows8 + (size_t)id * h->qstride;
    memset(dst, 0, h->qstride);
    kdbo_quantize(h->abs_max, vec, dst, dim);
    h->norms[id] = kdbo_int8_norm(dst, dim); /* :593-597, :1679-1683 */
  } else {
    float *dst = h->vecs + (size_t)id * h->stride;
    memset(dst, 0, h->stride * sizeof(float));
    memcpy(dst, vec, dim * sizeof(float));
    if (h->metric == KDBO_METRIC_COSINE) kdbo_normalize(dst, dim);
  }
}

/* ---- per-thread scratch (the reference's sync.Pool objects, hnsw_index.go:2358-2374) ---- */
typedef struct {
  uint64_t *visited;
  size_t visited_words;
  uint32_t *touched;
  size_t n_touched, cap_touched;
  heap cands, results;
  uint32_t *nbuf;
  cand *out;
  size_t out_cap;
} scratch;

static scratch *scratch_new(const kdbo_index *h) {
  scratch *s = (scratch *)calloc(1, sizeof *s);
  s->visited_words = ((size_t)h->cap >> 6) + 1;
  s->visited = (uint64_t *)calloc(s->visited_words, sizeof(uint64_t));
  s->cap_touched = 4096;
  s->touched = (uint32_t *)malloc(s->cap_touched * sizeof(uint32_t));
  s->nbuf = (uint32_t *)malloc((size_t)(h->mmax0 + 1) * sizeof(uint32_t));
  return s;
}
static void scratch_free(scratch *s) {
  if (!s) return;
  free(s->visited);
  free(s->touched);
  free(s->cands.a);
  free(s->results.a);
  free(s->nbuf);
  free(s->out);
  free(s);
}
static inline int vis_has(const scratch *s, uint32_t n) { /* bitset.go:33-42 */
  size_t b = n >> 6;
  if (b >= s->visited_words) return 0;
  return (s->visited[b] >> (n & 63)) & 1;
}
static inline void vis_add(scratch *s, uint32_t n) { /* bitset.go:22-31 */
  size_t b = n >> 6;
  if (b >= s->visited_words) return; /* ids above capacity cannot exist in this oracle */
  s->visited[b] |= (uint64_t)1 << (n & 63);
  if (s->n_touched == s->cap_touched) {
    s->cap_touched *= 2;
    s->touched = (uint32_t *)realloc(s->touched, s->cap_touched * sizeof(uint32_t));
  }
  s->touched[s->n_touched++] = n;
}
/* visited.Clear() (bitset.go:44-48) — same effect; only the touched words are rewritten */
static inline void vis_clear(scratch *s) {
  for (size_t i = 0; i < s->n_touched; i++) s->visited[s->touched[i] >> 6] = 0;
  s->n_touched = 0;
}
static inline int allow_has(const uint64_t *allow, size_t words, uint32_t id) {
  size_t b = id >> 6;
  if (b >= words) return 0;
  return (allow[b] >> (id & 63)) & 1;
}
static int allow_is_empty(const uint64_t *allow, size_t words) {
  for (size_t i = 0; i < words; i++)
    if (allow[i]) return 0;
  return 1;
}

/* searchLayerUnlocked, hnsw_index.go:2351-2611.  Returns the result count (<= k) written to
 * s->out ascending, or -1 when the entry node is nil (:2466-2468). */
static int search_layer(const kdbo_index *h, scratch *s, const void *q, uint32_t entry, int k, int level,
                        const uint64_t *allow, size_t allow_words, int ef_search, int concurrent,
                        kdbo_stats *st) {
  heap *cands = &s->cands, *results = &s->results;
  cands->n = 0;
  results->n = 0;
  int ef = ef_search; /* :2377-2380 */
  if (ef < k) ef = k;
  const float qnorm = h->precision == KDBO_PREC_I8 ? query_norm_i8((const int8_t *)q, (size_t)h->dim) : 0.0f;
  const int allow_active = allow != NULL && !allow_is_empty(allow, allow_words); /* :2481, :2545 */
  const uint32_t n_nodes = h->counter + 1; /* len(nodes) after growNodes */

  if (entry >= n_nodes || h->level[entry] < 0) return -1; /* :2463-2468 */
  double dist = query_dist(h, q, qnorm, entry);
  if (st) st->dist_evals++;
  cand ep = {entry, dist};
  min_push(cands, ep); /* :2478 */
  vis_add(s, entry);   /* :2479 */
  int ep_valid = 1;
  if (allow_active && !allow_has(allow, allow_words, entry)) ep_valid = 0; /* :2481-2485 */
  if (ep_valid && !h->deleted[entry]) max_push(results, ep);              /* :2487-2489 */

  while (cands->n > 0) { /* :2495 */
    cand cur = min_pop(cands);
    if ((int)results->n >= ef) { /* :2501-2506 */
      if (cur.d > results->a[0].d) break;
    }
    if (cur.id >= n_nodes) continue;                               /* :2510-2512 */
    if (h->level[cur.id] < 0 || level > h->level[cur.id]) continue; /* :2521-2524 */
    /* copy Connections[level] (:2528-2530) */
    uint32_t *row = conn_row(h, cur.id, level);
    if (concurrent) spin_lock(&h->lock[cur.id]);
    uint32_t cnt = row[0];
    memcpy(s->nbuf, row + 1, cnt * sizeof(uint32_t));
    if (concurrent) spin_unlock(&h->lock[cur.id]);
    if (st) {
      st->hops++;
      if (level == 0) st->hops_l0++;
    }
    for (uint32_t i = 0; i < cnt; i++) { /* :2537 */
      uint32_t nb = s->nbuf[i];
      if (vis_has(s, nb)) continue; /* :2539-2541 */
      vis_add(s, nb);               /* :2542 */
      if (allow_active && !allow_has(allow, allow_words, nb)) continue; /* :2545-2549 */
      if (nb >= n_nodes) continue;                                      /* :2553-2556 */
      if (__atomic_load_n(&h->level[nb], __ATOMIC_ACQUIRE) < 0) continue; /* nil node :2559-2561 */
      double d = query_dist(h, q, qnorm, nb); /* :2566 */
      if (st) st->dist_evals++;
      int admit = (int)results->n < ef; /* :2572-2577: worstDist = MaxFloat64 when empty */
      if (!admit) admit = d < results->a[0].d;
      if (admit) {
        cand c = {nb, d};
        min_push(cands, c); /* :2581 */
        if (!h->deleted[nb]) { /* :2584 */
          max_push(results, c);
          if ((int)results->n > ef) (void)max_pop(results); /* :2587-2589 */
        }
      }
    }
  }
  /* :2596-2610 drain from the back, truncate to k */
  size_t count = results->n;
  if (s->out_cap < count + 1) {
    s->out_cap = count + 64;
    s->out = (cand *)realloc(s->out, s->out_cap * sizeof(cand));
  }
  for (size_t i = count; i-- > 0;) s->out[i] = max_pop(results);
  vis_clear(s); /* deferred visited.Clear(), :2367 */
  return (int)(count > (size_t)k ? (size_t)k : count);
}

/* selectNeighbors, hnsw_index.go:2629-2701.  `sel` receives at most m ids; returns the count. */
static int select_neighbors(const kdbo_index *h, const cand *c, size_t n, int m, cand *sel) {
  if (n <= (size_t)m) { /* :2634-2636 */
    memcpy(sel, c, n * sizeof(cand));
    return (int)n;
  }
  cand *disc = (cand *)malloc(n * sizeof(cand));
  size_t nd = 0;
  int nr = 0;
  for (size_t w = 0; w < n && nr < m; w++) { /* :2642 */
    cand e = c[w];
    if (nr == 0) {
      sel[nr++] = e;
      continue;
    }
    int good = 1;
    for (int r = 0; r < nr; r++) { /* :2652-2679 */
      if (h->level[e.id] < 0 || h->level[sel[r].id] < 0) {
        good = 0;
        break;
      }
      double d = node_dist(h, e.id, sel[r].id);
      if (d < e.d) {
        good = 0;
        break;
      }
    }
    if (good)
      sel[nr++] = e;
    else
      disc[nd++] = e;
  }
  for (size_t i = 0; i < nd && nr < m; i++) sel[nr++] = disc[i]; /* :2689-2698 */
  free(disc);
  return nr;
}

int kdbo_select_neighbors(const kdbo_index *h, const uint32_t *ids, const double *d, size_t n, int m,
                          uint32_t *out_ids) {
  cand *c = (cand *)malloc((n + 1) * sizeof(cand)), *sel = (cand *)malloc((n + 1) * sizeof(cand));
  for (size_t i = 0; i < n; i++) {
    c[i].id = ids[i];
    c[i].d = d[i];
  }
  int r = select_neighbors(h, c, n, m, sel);
  for (int i = 0; i < r; i++) out_ids[i] = sel[i].id;
  free(c);
  free(sel);
  return r;
}

/* Index.Add, hnsw_index.go:472-809 */
static uint32_t add_one(kdbo_index *h, scratch *s, const float *vec, double u, int concurrent) {
  /* int8: the first Add trains the quantizer on its own vector (ensureQuantizerTrained, :510-517) */
  if (h->precision == KDBO_PREC_I8 && h->abs_max == 0.0f) h->abs_max = kdbo_train_quantizer(vec, 1, (size_t)h->dim);
  /* PHASE 1 under metaMu (:559-676) */
  spin_lock(&h->meta_lock);
  if (h->counter >= h->cap) {
    spin_unlock(&h->meta_lock);
    return 0;
  }
  uint32_t id = h->counter + 1; /* ids start at 1, :590 */
  store_row(h, id, vec);      /* :485-520 (normalise / convert / quantise) */
  int level = kdbo_random_level(u, h->m, h->max_level); /* :647 */
  if (level > 120) level = 120;
  if (level > 0) h->upper[id] = (uint32_t *)calloc((size_t)level * (size_t)(h->m + 1), sizeof(uint32_t));
  h->l0[(size_t)id * (size_t)(h->mmax0 + 1)] = 0;
  __atomic_store_n(&h->level[id], (int8_t)level, __ATOMIC_RELEASE);
  __atomic_store_n(&h->counter, id, __ATOMIC_RELEASE);
  int cur_max = h->max_level;
  if (cur_max == -1) { /* first node, :658-670 */
    h->entry = id;
    h->max_level = level;
    spin_unlock(&h->meta_lock);
    return id;
  }
  uint32_t ep = h->entry;
  spin_unlock(&h->meta_lock);

  const void *q = row_ptr(h, id); /* currObj := storedVector (:674) */
  const int efc = h->efc;
  /* zoom in, :685-690 */
  for (int l = cur_max; l > level; l--) {
    int n = search_layer(h, s, q, ep, 1, l, NULL, 0, 1, concurrent, NULL);
    if (n > 0) ep = s->out[0].id;
  }
  int top = level < cur_max ? level : cur_max; /* :694-697 */
  cand *cands = (cand *)malloc((size_t)(efc + 1) * sizeof(cand));
  cand *sel = (cand *)malloc((size_t)(efc + h->mmax0 + 2) * sizeof(cand));
  cand *allc = (cand *)malloc((size_t)(h->mmax0 + 2) * sizeof(cand));
  cand *best = (cand *)malloc((size_t)(h->mmax0 + 2) * sizeof(cand));
  uint32_t *cur = (uint32_t *)malloc((size_t)(h->mmax0 + 2) * sizeof(uint32_t));
  for (int l = top; l >= 0; l--) { /* :699 */
    int n = search_layer(h, s, q, ep, efc, l, NULL, 0, efc, concurrent, NULL);
    if (n < 0) continue; /* :702-704 */
    memcpy(cands, s->out, (size_t)n * sizeof(cand));
    int max_m = l == 0 ? h->mmax0 : h->m; /* :707-710 */
    int ns = select_neighbors(h, cands, (size_t)n, max_m, sel);
    /* forward links, :717-722 */
    uint32_t *row = conn_row(h, id, l);
    if (concurrent) spin_lock(&h->lock[id]);
    row[0] = (uint32_t)ns;
    for (int i = 0; i < ns; i++) row[1 + i] = sel[i].id;
    if (concurrent) spin_unlock(&h->lock[id]);
    /* reverse links, :725-783 */
    for (int i = 0; i < ns; i++) {
      uint32_t nb = sel[i].id;
      if (h->level[nb] < 0 || h->deleted[nb]) continue; /* :731-734 */
      if (l > h->level[nb]) continue;                    /* cannot happen outside races (:774-778) */
      uint32_t *nrow = conn_row(h, nb, l);
      if (concurrent) spin_lock(&h->lock[nb]);
      uint32_t ccount = nrow[0];
      memcpy(cur, nrow + 1, ccount * sizeof(uint32_t)); /* :737-742 */
      if (concurrent) spin_unlock(&h->lock[nb]);
      uint32_t nfinal;
      if ((int)ccount < max_m) { /* :748-752 */
        cur[ccount] = id;
        nfinal = ccount + 1;
      } else { /* prune, :753-771 */
        size_t na = 0;
        for (uint32_t j = 0; j < ccount; j++) {
          uint32_t e = cur[j];
          if (e <= h->counter && h->level[e] >= 0 && !h->deleted[e]) {
            allc[na].id = e;
            allc[na].d = node_dist(h, nb, e);
            na++;
          }
        }
        allc[na].id = id;
        allc[na].d = node_dist(h, nb, id);
        na++;
        int nbest = select_neighbors(h, allc, na, max_m, best);
        for (int j = 0; j < nbest; j++) cur[j] = best[j].id;
        nfinal = (uint32_t)nbest;
      }
      if (concurrent) spin_lock(&h->lock[nb]); /* :774-782 */
      memcpy(nrow + 1, cur, nfinal * sizeof(uint32_t));
      nrow[0] = nfinal;
      if (concurrent) spin_unlock(&h->lock[nb]);
    }
    if (n > 0) ep = cands[0].id; /* :786-788 */
  }
  free(cands);
  free(sel);
  free(allc);
  free(best);
  free(cur);
  if (level > cur_max) { /* :793-801 */
    spin_lock(&h->meta_lock);
    if (level > h->max_level) {
      h->max_level = level;
      h->entry = id;
    }
    spin_unlock(&h->meta_lock);
  }
  return id;
}

uint32_t kdbo_add(kdbo_index *h, const float *vec, double u) {
  scratch *s = scratch_new(h);
  uint32_t id = add_one(h, s, vec, u, 0);
  scratch_free(s);
  return id;
}

int kdbo_add_many(kdbo_index *h, const float *vecs, size_t n, const double *u, int threads) {
  const size_t dim = (size_t)h->dim;
  if (threads <= 1) {
    scratch *s = scratch_new(h);
    for (size_t i = 0; i < n; i++)
      if (!add_one(h, s, vecs + i * dim, u[i], 0)) {
        scratch_free(s);
        return -1;
      }
    scratch_free(s);
    return 0;
  }
  int failed = 0;
  /* a short sequential prefix so the graph exists before threads race on it */
  size_t prefix = n < 256 ? n : 256;
  {
    scratch *s = scratch_new(h);
    for (size_t i = 0; i < prefix; i++)
      if (!add_one(h, s, vecs + i * dim, u[i], 0)) failed = 1;
    scratch_free(s);
  }
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    scratch *s = scratch_new(h);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
    for (long long i = (long long)prefix; i < (long long)n; i++)
      if (!add_one(h, s, vecs + (size_t)i * dim, u[i], 1)) failed = 1;
    scratch_free(s);
  }
  return failed ? -1 : 0;
}

static int cand_cmp(const void *a, const void *b);
/* addBatchInternal, hnsw_index.go:1479-2088.  Deterministic restatement: phase 1 (every new node
 * searches the PRE-batch graph, :1789-1853) and phase 3 (per-target commit, :1897-2060) only read
 * state that no other worker writes, so the result does not depend on the thread count.  The two
 * unstable sorts of the reference (:1915, :2028) are made total by (NodeID, Level) / (distance, id). */
typedef struct {
  uint32_t target;
  int32_t level;
  uint32_t src;
} link_req;
static int link_req_cmp(const void *a, const void *b) {
  const link_req *x = (const link_req *)a, *y = (const link_req *)b;
  if (x->target != y->target) return x->target < y->target ? -1 : 1;
  if (x->level != y->level) return x->level < y->level ? -1 : 1;
  if (x->src != y->src) return x->src < y->src ? -1 : 1;
  return 0;
}
static int u32_cmp(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

int kdbo_add_batch(kdbo_index *h, const float *vecs, size_t n, const double *u, int ef_const, int threads) {
  const size_t dim = (size_t)h->dim;
  if (n == 0) return 0;
  if (ef_const <= 0) ef_const = h->efc;
  if (threads < 1) threads = 1;
  if ((uint64_t)h->counter < (uint64_t)ef_const) { /* :1502-1513 small graph -> single Adds */
    scratch *s = scratch_new(h);
    for (size_t i = 0; i < n; i++)
      if (!add_one(h, s, vecs + i * dim, u[i], 0)) {
        scratch_free(s);
        return -1;
      }
    scratch_free(s);
    return 0;
  }
  if ((uint64_t)h->counter + n > (uint64_t)h->cap) return -1;
  /* phase 0/1A/1B: normalise, reserve the id block, create nodes (:1533-1757) */
  const uint32_t start_id = h->counter + 1;
  const int pre_max = h->max_level;
  const uint32_t pre_entry = h->entry;
  for (size_t i = 0; i < n; i++) {
    uint32_t id = start_id + (uint32_t)i;
    store_row(h, id, vecs + i * dim); /* :1553-1577 */
    int level = kdbo_random_level(u[i], h->m, pre_max);              /* :1738 */
    if (level > 120) level = 120;
    if (level > 0) h->upper[id] = (uint32_t *)calloc((size_t)level * (size_t)(h->m + 1), sizeof(uint32_t));
    h->l0[(size_t)id * (size_t)(h->mmax0 + 1)] = 0;
    h->level[id] = (int8_t)level;
  }
  h->counter = start_id + (uint32_t)n - 1;

  /* phase 1: neighbour search on the pre-batch graph (:1789-1853) */
  size_t req_cap = 0;
  for (size_t i = 0; i < n; i++) {
    int L = h->level[start_id + i];
    req_cap += (size_t)((L < pre_max ? L : pre_max) + 1) * (size_t)ef_const * 2;
  }
  link_req *reqs = (link_req *)malloc((req_cap + 1) * sizeof(link_req));
  size_t n_reqs = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    scratch *s = scratch_new(h);
    link_req *local = (link_req *)malloc(((size_t)(pre_max + 1) * (size_t)ef_const * 2 + 1) * sizeof(link_req));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
    for (long long i = 0; i < (long long)n; i++) {
      uint32_t id = start_id + (uint32_t)i;
      const void *q = row_ptr(h, id);
      int node_level = h->level[id];
      uint32_t ep = pre_entry;
      size_t nl = 0;
      for (int l = pre_max; l > node_level; l--) { /* :1819-1824 */
        int c = search_layer(h, s, q, ep, 1, l, NULL, 0, 1, 0, NULL);
        if (c > 0) ep = s->out[0].id;
      }
      for (int l = node_level < pre_max ? node_level : pre_max; l >= 0; l--) { /* :1827-1850 */
        int c = search_layer(h, s, q, ep, ef_const, l, NULL, 0, ef_const, 0, NULL);
        if (c <= 0) continue;
        for (int j = 0; j < c; j++) {
          link_req d = {id, l, s->out[j].id}; /* direct: node <- candidate   (:1880-1881) */
          link_req r = {s->out[j].id, l, id}; /* reverse: candidate <- node (:1885-1892) */
          local[nl++] = d;
          local[nl++] = r;
        }
        ep = s->out[0].id;
      }
#ifdef _OPENMP
#pragma omp critical
#endif
      {
        memcpy(reqs + n_reqs, local, nl * sizeof(link_req));
        n_reqs += nl;
      }
    }
    free(local);
    scratch_free(s);
  }
  qsort(reqs, n_reqs, sizeof(link_req), link_req_cmp);

  /* phase 3: one commit per (target, level) group (:1926-2056) */
  size_t n_groups = 0;
  size_t *gstart = (size_t *)malloc((n_reqs + 1) * sizeof(size_t));
  for (size_t i = 0; i < n_reqs; i++)
    if (i == 0 || reqs[i].target != reqs[i - 1].target || reqs[i].level != reqs[i - 1].level) gstart[n_groups++] = i;
  gstart[n_groups] = n_reqs;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    size_t ucap = 1024;
    uint32_t *uniq = (uint32_t *)malloc(ucap * sizeof(uint32_t));
    cand *cs = (cand *)malloc(ucap * sizeof(cand)), *sel = (cand *)malloc(ucap * sizeof(cand));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
    for (long long g = 0; g < (long long)n_groups; g++) {
      size_t a = gstart[g], b = gstart[g + 1];
      uint32_t t = reqs[a].target;
      int lvl = reqs[a].level;
      if (t > h->counter || h->level[t] < 0 || h->deleted[t]) continue; /* :1933-1936 */
      if (lvl > h->level[t]) continue; /* Connections would be grown (:2045-2049); unreachable here */
      uint32_t *row = conn_row(h, t, lvl);
      size_t need = (size_t)row[0] + (b - a) + 1;
      if (need > ucap) {
        ucap = need * 2;
        uniq = (uint32_t *)realloc(uniq, ucap * sizeof(uint32_t));
        cs = (cand *)realloc(cs, ucap * sizeof(cand));
        sel = (cand *)realloc(sel, ucap * sizeof(cand));
      }
      size_t nu = 0;
      for (uint32_t j = 0; j < row[0]; j++) uniq[nu++] = row[1 + j]; /* :1980-1986 */
      for (size_t j = a; j < b; j++) uniq[nu++] = reqs[j].src;
      qsort(uniq, nu, sizeof(uint32_t), u32_cmp); /* :1990 */
      size_t cnt = 0;                              /* :1992-2008 drop self + duplicates */
      for (size_t j = 0; j < nu; j++) {
        if (uniq[j] == t) continue;
        if (cnt > 0 && uniq[cnt - 1] == uniq[j]) continue;
        uniq[cnt++] = uniq[j];
      }
      int max_m = lvl == 0 ? h->mmax0 : h->m; /* :2010-2013 */
      if (cnt <= (size_t)max_m) {             /* :2016-2018 ascending-id list */
        for (size_t j = 0; j < cnt; j++) row[1 + j] = uniq[j];
        row[0] = (uint32_t)cnt;
      } else { /* prune :2019-2042 */
        size_t nc = 0;
        for (size_t j = 0; j < cnt; j++) {
          uint32_t e = uniq[j];
          if (e <= h->counter && h->level[e] >= 0 && !h->deleted[e]) {
            cs[nc].id = e;
            cs[nc].d = node_dist(h, t, e);
            nc++;
          }
        }
        qsort(cs, nc, sizeof(cand), cand_cmp); /* ascending distance, ties by id */
        int ns = select_neighbors(h, cs, nc, max_m, sel);
        for (int j = 0; j < ns; j++) row[1 + j] = sel[j].id;
        row[0] = (uint32_t)ns;
      }
    }
    free(uniq);
    free(cs);
    free(sel);
  }
  free(gstart);
  free(reqs);
  /* phase 4: entry point (:2066-2080) */
  for (size_t i = 0; i < n; i++) {
    uint32_t id = start_id + (uint32_t)i;
    if (h->level[id] > h->max_level) {
      h->max_level = h->level[id];
      h->entry = id;
    }
  }
  return 0;
}

void kdbo_delete(kdbo_index *h, uint32_t id) { /* hnsw_index.go:2303-2336 */
  if (id == 0 || id > h->counter) return;
  if (h->level[id] >= 0) h->deleted[id] = 1;
}

/* searchInternal, hnsw_index.go:369-468 — returns count, results in s->out */
static int search_internal(const kdbo_index *h, scratch *s, const float *query, int k, int ef_search,
                           int needs_refine, const uint64_t *allow, size_t allow_words, float *qbuf,
                           kdbo_stats *st) {
  uint32_t entry = h->entry;
  int max_level = h->max_level;
  if (max_level == -1) return 0;                           /* :383-385 */
  int actual_ef = kdbo_effective_ef(ef_search, needs_refine); /* :387-399 */
  const void *q = query;
  if (h->metric == KDBO_METRIC_COSINE) { /* :406-414 */
    memcpy(qbuf, query, (size_t)h->dim * sizeof(float));
    kdbo_normalize(qbuf, (size_t)h->dim);
    q = qbuf;
  }
  /* adapt to the precision (:417-434); the converted query lives behind the f32 copy in qbuf */
  if (h->precision == KDBO_PREC_F16) {
    const float *src = (const float *)q;
    uint16_t *q16 = (uint16_t *)(qbuf + h->stride);
    for (int i = 0; i < h->dim; i++) q16[i] = kdbo_f32_to_f16(src[i]);
    q = q16;
  } else if (h->precision == KDBO_PREC_I8) {
    int8_t *q8 = (int8_t *)(qbuf + h->stride);
    kdbo_quantize(h->abs_max, (const float *)q, q8, (size_t)h->dim);
    q = q8;
  }
  if (allow != NULL) { /* :436-447 smart entry point */
    if (!allow_has(allow, allow_words, entry)) {
      uint32_t first = 0;
      int found = 0;
      for (size_t w = 0; w < allow_words && !found; w++)
        if (allow[w]) {
          first = (uint32_t)(w * 64 + (size_t)__builtin_ctzll(allow[w]));
          found = 1;
        }
      if (!found) return 0;
      entry = first;
    }
  }
  for (int l = max_level; l > 0; l--) { /* :450-459 */
    int n = search_layer(h, s, q, entry, 1, l, allow, allow_words, 0, 0, st);
    if (n < 0) return 0;  /* err -> SearchWithScores logs and returns [] (:355-359) */
    if (n == 0) return 0; /* "search failed at level" -> [] */
    entry = s->out[0].id;
  }
  int n = search_layer(h, s, q, entry, k, 0, allow, allow_words, actual_ef, 0, st); /* :462 */
  return n < 0 ? 0 : n;
}

int kdbo_search(const kdbo_index *h, const float *query, int k, int ef_search, int needs_refine,
                const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                kdbo_stats *stats) {
  scratch *s = scratch_new(h);
  float *qbuf = (float *)malloc(2 * h->stride * sizeof(float));
  int n = search_internal(h, s, query, k, ef_search, needs_refine, allow, allow_words, qbuf, stats);
  for (int i = 0; i < n; i++) { /* SearchWithScores :361-364: Score = raw distance */
    out_ids[i] = s->out[i].id;
    out_scores[i] = s->out[i].d;
  }
  free(qbuf);
  scratch_free(s);
  return n;
}

int kdbo_search_batch(const kdbo_index *h, const float *queries, size_t nq, int k, int ef_search,
                      int needs_refine, const uint64_t *allow, size_t allow_words, uint32_t *out_ids,
                      double *out_scores, int32_t *out_counts, kdbo_stats *stats, int threads) {
  if (threads < 1) threads = 1;
  kdbo_stats total = {0, 0, 0};
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    scratch *s = scratch_new(h);
    float *qbuf = (float *)malloc(2 * h->stride * sizeof(float));
    kdbo_stats local = {0, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
    for (long long qi = 0; qi < (long long)nq; qi++) {
      int n = search_internal(h, s, queries + (size_t)qi * (size_t)h->dim, k, ef_search, needs_refine, allow,
                              allow_words, qbuf, &local);
      for (int i = 0; i < k; i++) {
        out_ids[(size_t)qi * (size_t)k + (size_t)i] = i < n ? s->out[i].id : 0;
        out_scores[(size_t)qi * (size_t)k + (size_t)i] = i < n ? s->out[i].d : 0.0;
      }
      out_counts[qi] = n;
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    {
      total.dist_evals += local.dist_evals;
      total.hops += local.hops;
      total.hops_l0 += local.hops_l0;
    }
    free(qbuf);
    scratch_free(s);
  }
  if (stats) *stats = total;
  return 0;
}

int kdbo_search_layer(const kdbo_index *h, const void *q, uint32_t entry, int k, int level,
                      const uint64_t *allow, size_t allow_words, int ef_search, uint32_t *out_ids,
                      double *out_scores, kdbo_stats *stats) {
  scratch *s = scratch_new(h);
  int n = search_layer(h, s, q, entry, k, level, allow, allow_words, ef_search, 0, stats);
  for (int i = 0; i < n; i++) {
    out_ids[i] = s->out[i].id;
    out_scores[i] = s->out[i].d;
  }
  scratch_free(s);
  return n;
}

/* ======================================================================================
 * 4. Flat scan
 * ==================================================================================== */
static inline int flat_less(double da, uint32_t ia, double db, uint32_t ib) {
  return da < db || (da == db && ia < ib);
}
/* bounded max-heap on (d, id) keeping the k smallest */
static void topk_offer(cand *hp, int *n, int k, cand c) {
  if (*n < k) {
    int j = (*n)++;
    hp[j] = c;
    while (j > 0) {
      int i = (j - 1) / 2;
      if (!flat_less(hp[i].d, hp[i].id, hp[j].d, hp[j].id)) break;
      SWAP(hp[i], hp[j]);
      j = i;
    }
    return;
  }
  if (!flat_less(c.d, c.id, hp[0].d, hp[0].id)) return;
  hp[0] = c;
  int i = 0;
  for (;;) {
    int j1 = 2 * i + 1;
    if (j1 >= k) break;
    int j = j1;
    if (j1 + 1 < k && flat_less(hp[j1].d, hp[j1].id, hp[j1 + 1].d, hp[j1 + 1].id)) j = j1 + 1;
    if (!flat_less(hp[i].d, hp[i].id, hp[j].d, hp[j].id)) break;
    SWAP(hp[i], hp[j]);
    i = j;
  }
}
static int cand_cmp(const void *a, const void *b) {
  const cand *x = (const cand *)a, *y = (const cand *)b;
  if (flat_less(x->d, x->id, y->d, y->id)) return -1;
  if (flat_less(y->d, y->id, x->d, x->id)) return 1;
  return 0;
}

int kdbo_flat_search_batch(const kdbo_index *h, const float *queries, size_t nq, int k, int mode,
                           const uint64_t *allow, size_t allow_words, uint32_t *out_ids,
                           double *out_scores, int32_t *out_counts, int threads) {
  if (threads < 1) threads = 1;
  if (h->precision != KDBO_PREC_F32) return -1; /* BruteForceIndex holds []float32 only */
  const size_t dim = (size_t)h->dim;
  /* BruteForceIndex treats an empty allow-list as unfiltered (vector_index.go:132) */
  const int allow_active = allow != NULL && !allow_is_empty(allow, allow_words);
  const int cosine = mode == 1 && h->metric == KDBO_METRIC_COSINE;
  const size_t QB = 16; /* queries per block so a corpus row is reused from L1 */
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    cand *hp = (cand *)malloc(QB * (size_t)k * sizeof(cand));
    int *hn = (int *)malloc(QB * sizeof(int));
    double *qd = (double *)malloc(QB * dim * sizeof(double));
    float *qf = (float *)malloc(dim * sizeof(float));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (long long b = 0; b < (long long)((nq + QB - 1) / QB); b++) {
      size_t q0 = (size_t)b * QB, q1 = q0 + QB > nq ? nq : q0 + QB;
      for (size_t qi = q0; qi < q1; qi++) {
        memcpy(qf, queries + qi * dim, dim * sizeof(float));
        if (cosine) kdbo_normalize(qf, dim);
        for (size_t e = 0; e < dim; e++) qd[(qi - q0) * dim + e] = (double)qf[e];
        hn[qi - q0] = 0;
      }
      for (uint32_t id = 1; id <= h->counter; id++) {
        if (h->level[id] < 0 || h->deleted[id]) continue;
        if (allow_active && !allow_has(allow, allow_words, id)) continue;
        const float *x = kdbo_vector(h, id);
        for (size_t qi = q0; qi < q1; qi++) {
          const double *qq = qd + (qi - q0) * dim;
          const float *qfl = queries + qi * dim;
          double sum = 0.0;
          if (cosine) {
            for (size_t e = 0; e < dim; e++) sum += qq[e] * (double)x[e];
            sum = 1.0 - sum;
          } else if (mode == 0) { /* diff in f32, then widened: vector_index.go:158 */
            for (size_t e = 0; e < dim; e++) {
              double diff = (double)(qfl[e] - x[e]);
              sum += diff * diff;
            }
          } else {
            for (size_t e = 0; e < dim; e++) {
              double diff = qq[e] - (double)x[e];
              sum += diff * diff;
            }
          }
          cand c = {id, sum};
          topk_offer(hp + (qi - q0) * (size_t)k, &hn[qi - q0], k, c);
        }
      }
      for (size_t qi = q0; qi < q1; qi++) {
        cand *r = hp + (qi - q0) * (size_t)k;
        int n = hn[qi - q0];
        qsort(r, (size_t)n, sizeof(cand), cand_cmp);
        for (int i = 0; i < k; i++) {
          out_ids[qi * (size_t)k + (size_t)i] = i < n ? r[i].id : 0;
          out_scores[qi * (size_t)k + (size_t)i] = i < n ? r[i].d : 0.0;
        }
        out_counts[qi] = n;
      }
    }
    free(hp);
    free(hn);
    free(qd);
    free(qf);
  }
  return 0;
}

/* ======================================================================================
 * 5. Graph exchange
 * ==================================================================================== */
void kdbo_export_sizes(const kdbo_index *h, uint64_t *n_rows, uint64_t *n_edges) {
  uint64_t rows = 0, edges = 0;
  for (uint32_t id = 0; id <= h->counter; id++) {
    int L = h->level[id];
    for (int l = 0; l <= L; l++) {
      rows++;
      edges += conn_row(h, id, l)[0];
    }
  }
  *n_rows = rows;
  *n_edges = edges;
}
void kdbo_export_graph(const kdbo_index *h, int32_t *levels, uint64_t *node_row, uint64_t *row_off,
                       uint32_t *nbrs, uint8_t *deleted) {
  uint64_t r = 0, e = 0;
  for (uint32_t id = 0; id <= h->counter; id++) {
    int L = h->level[id];
    levels[id] = L;
    deleted[id] = h->deleted[id];
    node_row[id] = r;
    for (int l = 0; l <= L; l++) {
      const uint32_t *row = conn_row(h, id, l);
      row_off[r++] = e;
      memcpy(nbrs + e, row + 1, row[0] * sizeof(uint32_t));
      e += row[0];
    }
  }
  node_row[h->counter + 1] = r;
  row_off[r] = e;
}
int kdbo_import_graph(kdbo_index *h, uint32_t n, const void *rows_any, size_t row_stride, const int32_t *levels,
                      const uint64_t *node_row, const uint64_t *row_off, const uint32_t *nbrs,
                      const uint8_t *deleted, uint32_t entry, int max_level) {
  if (n > h->cap) return -1;
  for (uint32_t id = 0; id <= h->counter; id++) {
    free(h->upper[id]);
    h->upper[id] = NULL;
    h->level[id] = -1;
    h->deleted[id] = 0;
  }
  for (uint32_t id = 1; id <= n; id++) {
    if (h->precision == KDBO_PREC_F16) { /* rows in stored form: float16 bits */
      uint16_t *dst = h->rows16 + (size_t)id * h->qstride;
      memset(dst, 0, h->qstride * sizeof(uint16_t));
      memcpy(dst, (const uint16_t *)rows_any + (size_t)id * row_stride, (size_t)h->dim * sizeof(uint16_t));
    } else if (h->precision == KDBO_PREC_I8) { /* int8 rows; norms follow from them (computeInt8Norm) */
      int8_t *dst = h->rows8 + (size_t)id * h->qstride;
      memset(dst, 0, h->qstride);
      memcpy(dst, (const int8_t *)rows_any + (size_t)id * row_stride, (size_t)h->dim);
      h->norms[id] = kdbo_int8_norm(dst, (size_t)h->dim);
    } else {
      const float *rows = (const float *)rows_any;
      memset(h->vecs + (size_t)id * h->stride, 0, h->stride * sizeof(float));
      memcpy(h->vecs + (size_t)id * h->stride, rows + (size_t)id * row_stride, (size_t)h->dim * sizeof(float));
    }
    int L = levels[id];
    if (L > 120) return -1;
    h->level[id] = (int8_t)L;
    h->deleted[id] = deleted ? deleted[id] : 0;
    if (L > 0) h->upper[id] = (uint32_t *)calloc((size_t)L * (size_t)(h->m + 1), sizeof(uint32_t));
    for (int l = 0; l <= L; l++) {
      uint64_t r = node_row[id] + (uint64_t)l;
      uint64_t cnt = row_off[r + 1] - row_off[r];
      uint64_t maxc = (uint64_t)(l == 0 ? h->mmax0 : h->m);
      if (cnt > maxc) return -2;
      uint32_t *row = conn_row(h, id, l);
      row[0] = (uint32_t)cnt;
      memcpy(row + 1, nbrs + row_off[r], cnt * sizeof(uint32_t));
    }
  }
  h->counter = n;
  h->entry = entry;
  h->max_level = max_level;
  return 0;
}
