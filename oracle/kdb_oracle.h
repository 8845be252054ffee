/*
 * kdb_oracle.h — CPU ORACLE for the KektorDB vector-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (sanonone/kektordb @ 4b02a6e); it is NOT part of the product.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / CPU baseline.  The product path
 * (kektordb_b200/, libkektordb_gpu.so) never links or calls anything in here.
 *
 * Parity pinning: the reference is Go 1.26 + Rust; neither toolchain exists in
 * the build container, so the reference itself cannot be executed.  The oracle
 * is pinned against every known-answer vector the reference's own tests hold
 * for this path (pkg/core/distance/distance_test.go:35-88,
 * native/compute/src/lib.rs:423-458, pkg/core/hnsw/hnsw_heap_test.go:9-54) and
 * the behavioural tests (exact-match-first, filter id sets, recall >= 0.95 on
 * 10k x 64 L2) — see tests/test_oracle_golden.py.  The cosine-f32 summation
 * order lives in gonum v0.16.0 (not vendored): "parity unpinned" for that one
 * order; absorbed by the 1e-5 score tolerance.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef KDB_ORACLE_H
#define KDB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDBO_METRIC_L2 0     /* squared Euclidean, no sqrt (distance_go.go:57-68)   */
#define KDBO_METRIC_COSINE 1 /* 1 - dot on unit vectors   (distance_go.go:122-128) */

/* Summation order used for the f32 distance.  Search logic is identical in all modes. */
#define KDBO_ARITH_SEQ 0    /* sequential f32, no FMA  (pure-Go loops, distance_go.go:57-89)        */
#define KDBO_ARITH_AVX2 1   /* 8-lane FMA + hadd + scalar tail (native/compute/src/lib.rs:22-99)   */
#define KDBO_ARITH_KERNEL 2 /* the exact lane/tree order of the sm_100a kernel      */

/* PrecisionType (distance_go.go:41-46).  float16 exists for Euclidean only, int8 for Cosine only
 * (float16Funcs / int8Funcs, distance_go.go:139-146). */
#define KDBO_PREC_F32 0
#define KDBO_PREC_F16 1
#define KDBO_PREC_I8 2

typedef struct kdbo_index kdbo_index;

typedef struct {
  uint64_t dist_evals; /* E: distance evaluations (entry points included)                */
  uint64_t hops;       /* H: candidate expansions that read an adjacency row             */
  uint64_t hops_l0;    /* expansions on level 0 (row of 2M ids); the rest are upper rows  */
} kdbo_stats;

/* ---- lifecycle ------------------------------------------------------------------ */
/* hnsw.New defaults: m<=0 -> 16, efc<=0 -> 200, mMax0 = 2m, ml = 1/ln m (hnsw_index.go:138-151) */
kdbo_index *kdbo_new(int dim, int metric, int m, int ef_construction, int arith, uint32_t capacity);
/* hnsw.New with a precision; NULL for a (metric, precision) pair the reference does not offer */
kdbo_index *kdbo_new_ex(int dim, int metric, int precision, int m, int ef_construction, int arith,
                        uint32_t capacity);
void kdbo_free(kdbo_index *);
int kdbo_precision(const kdbo_index *);
/* Quantizer.AbsMax (pkg/core/distance/quantizer.go:19-22); set before adding int8 rows, as
 * DB.Compress does through TrainQuantizer (pkg/core/core.go:1224-1232) */
void kdbo_set_quantizer(kdbo_index *, float abs_max);
float kdbo_abs_max(const kdbo_index *);
void kdbo_set_arith(kdbo_index *, int arith);

/* ---- build (graph construction stays CPU in the product; restated to obtain graphs) -- */
/* Index.Add (hnsw_index.go:472-809).  u in [0,1) is the rand.Float64() draw of randomLevel
 * (:2616-2625).  Returns the internal id (ids start at 1, :590), 0 on error. */
uint32_t kdbo_add(kdbo_index *, const float *vec, double u);
/* n Adds in slice order.  threads<=1: sequential, deterministic (Appendix A rule 15).
 * threads>1: concurrent Add calls (legal reference behaviour, not reproducible). */
int kdbo_add_many(kdbo_index *, const float *vecs, size_t n, const double *u, int threads);
/* addBatchInternal (hnsw_index.go:1479-2088): below ef_const nodes the batch goes through single
 * Adds (:1502-1513); otherwise every member searches the pre-batch graph, then per-node commits.
 * ef_const <= 0 -> efConstruction (AddBatch :1466); AddBatchFast passes max(2M, 40) (:1470-1476).
 * Deterministic for any thread count. */
int kdbo_add_batch(kdbo_index *, const float *vecs, size_t n, const double *u, int ef_const, int threads);
/* Index.Delete: soft delete (hnsw_index.go:2303-2336). */
void kdbo_delete(kdbo_index *, uint32_t internal_id);

/* ---- search --------------------------------------------------------------------- */
/* SearchWithScores (hnsw_index.go:343-468).  `allow`: NULL = nil bitmap; otherwise a dense
 * bitset over internal ids (bit i of word i/64), allow_words long — same membership as the
 * roaring bitmap from FindIDsByFilter.  ef_search is the value the caller passes
 * (0 => ef = k).  needs_refine applies the boost of :387-399.  Returns number of results
 * (<= k), ascending distance; out_scores are the raw float64 distances. */
int kdbo_search(const kdbo_index *, const float *query, int k, int ef_search, int needs_refine,
                const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                kdbo_stats *stats);
/* nq independent searches, one query per thread.  out_* are [nq][k], out_counts [nq]. */
int kdbo_search_batch(const kdbo_index *, const float *queries, size_t nq, int k, int ef_search,
                      int needs_refine, const uint64_t *allow, size_t allow_words, uint32_t *out_ids,
                      double *out_scores, int32_t *out_counts, kdbo_stats *stats, int threads);
/* searchLayerUnlocked (hnsw_index.go:2351-2611) on one level, exposed for tests. The query
 * must already be prepared (normalised for cosine). */
int kdbo_search_layer(const kdbo_index *, const void *prepared_query, uint32_t entry, int k,
                      int level, const uint64_t *allow, size_t allow_words, int ef_search,
                      uint32_t *out_ids, double *out_scores, kdbo_stats *stats);

/* ---- flat ----------------------------------------------------------------------- */
/* mode 0: BruteForceIndex (pkg/core/vector_index.go:104-162): sum of float64(q_i - x_i)^2,
 *         ascending, allow-list post-filter, first k.  Runs on the RAW query (no normalise).
 * mode 1: exact f64 distance under the index metric on the stored vectors (ground truth
 *         for recall: cosine = 1 - sum f64(q^_i * x_i), L2 = as mode 0); query prepared as
 *         in searchInternal.  Ties broken by ascending id (reference order is unspecified).
 * Deleted nodes are skipped. */
int kdbo_flat_search_batch(const kdbo_index *, const float *queries, size_t nq, int k, int mode,
                           const uint64_t *allow, size_t allow_words, uint32_t *out_ids,
                           double *out_scores, int32_t *out_counts, int threads);

/* ---- primitives, exposed for golden-vector tests ----------------------------------- */
double kdbo_distance(int metric, int arith, const float *a, const float *b, size_t n);
/* portable restatement of the same order without intrinsics; must be bit-identical */
double kdbo_distance_generic(int metric, int arith, const float *a, const float *b, size_t n);
float kdbo_sq_euclid_f32(int arith, const float *a, const float *b, size_t n);
float kdbo_dot_f32(int arith, const float *a, const float *b, size_t n);
void kdbo_normalize(float *v, size_t n);      /* hnsw_index.go:3030-3045 */
/* float16 / int8 precisions */
uint16_t kdbo_f32_to_f16(float f);             /* float16.Fromfloat32(f).Bits(), IEEE RNE        */
float kdbo_f16_to_f32(uint16_t h);             /* float16.Frombits(h).Float32(), exact           */
float kdbo_sq_euclid_f16(int arith, const uint16_t *a, const uint16_t *b, size_t n); /* distance_go.go:93-105, lib.rs:101-141 */
int32_t kdbo_dot_i8(const int8_t *a, const int8_t *b, size_t n);                     /* distance_go.go:108-118 */
float kdbo_int8_norm(const int8_t *v, size_t n);                                     /* hnsw_index.go:3371-3377 */
double kdbo_int8_cosine_distance(int32_t dot, float qnorm, float stored_norm);       /* hnsw_index.go:2421-2449 */
void kdbo_quantize(float abs_max, const float *v, int8_t *out, size_t n);            /* quantizer.go:135-160 */
float kdbo_train_quantizer(const float *vecs, size_t n, size_t dim);                 /* quantizer.go:49-125 */
int kdbo_random_level(double u, int m, int current_max); /* hnsw_index.go:2616-2625 */
double kdbo_score_from_distance(double d);    /* 1/(1+d), pkg/engine/search_utils.go:48-52 */
int kdbo_effective_ef(int ef_search, int needs_refine); /* hnsw_index.go:387-399 */
/* heaps (hnsw_heap.go:18-156): push all, then pop all; kind 0 = minHeap, 1 = maxHeap */
void kdbo_heap_roundtrip(int kind, const uint32_t *ids, const double *d, size_t n, uint32_t *out_ids,
                         double *out_d);
/* selectNeighbors (hnsw_index.go:2629-2701) over candidates (id, dist) in the given order */
int kdbo_select_neighbors(const kdbo_index *, const uint32_t *ids, const double *d, size_t n, int m,
                          uint32_t *out_ids);

/* ---- introspection / graph exchange ------------------------------------------------ */
uint32_t kdbo_count(const kdbo_index *);     /* nodeCounter: highest internal id */
uint32_t kdbo_entry(const kdbo_index *);
int kdbo_max_level(const kdbo_index *);
int kdbo_dim(const kdbo_index *);
int kdbo_m(const kdbo_index *);
const float *kdbo_vector(const kdbo_index *, uint32_t id); /* stored (normalised) row; float32 indexes only */
const void *kdbo_row_raw(const kdbo_index *, uint32_t id);  /* stored row in the index precision  */
size_t kdbo_row_raw_stride(const kdbo_index *);             /* elements between stored rows       */
const float *kdbo_norms(const kdbo_index *);                /* int8: quantizedNorms, else NULL    */
size_t kdbo_row_stride(const kdbo_index *);                 /* floats between rows      */
/* Flattened adjacency: node i owns rows node_row[i]..node_row[i+1]-1 (one per level 0..L_i),
 * row r holds nbrs[row_off[r]..row_off[r+1]-1] in reference order. */
void kdbo_export_sizes(const kdbo_index *, uint64_t *n_rows, uint64_t *n_edges);
void kdbo_export_graph(const kdbo_index *, int32_t *levels /*[n+1]*/, uint64_t *node_row /*[n+2]*/,
                       uint64_t *row_off /*[rows+1]*/, uint32_t *nbrs /*[edges]*/,
                       uint8_t *deleted /*[n+1]*/);
/* Load vectors (already in stored form: f32 / float16 bits / int8 by the index precision; row_stride
 * in elements) + a graph produced elsewhere; replaces the content. */
int kdbo_import_graph(kdbo_index *, uint32_t n, const void *rows, size_t row_stride,
                      const int32_t *levels, const uint64_t *node_row, const uint64_t *row_off,
                      const uint32_t *nbrs, const uint8_t *deleted, uint32_t entry, int max_level);

#ifdef __cplusplus
}
#endif
#endif
