/*
 * kektordb_gpu.h — C ABI of the B200 (sm_100a) vector-search hot path for KektorDB.
 *
 * This header is the drop-in boundary: what the Go host binds through cgo
 * (`//go:build cuda`), next to the existing native header
 * native/compute/include/kektordb_compute.h:8-11 whose conventions it keeps:
 * plain C, borrowed pointers valid for the call only, no exceptions across the
 * boundary, int return (0 = ok, <0 = error), caller owns every buffer it passes,
 * the library owns device memory behind the opaque handle.
 *
 * Every entry point cites the reference interface it stands in for (paths relative
 * to the reference tree).  ids are the reference's internal ids: uint32, 0 reserved
 * (pkg/core/hnsw/hnsw_index.go:590).  Scores are the raw float64 distances the
 * reference returns (types.SearchResult{DocID uint32; Score float64},
 * pkg/core/types/types.go:11-14): squared L2, or 1 - dot on unit vectors.
 *
 * There is no CPU fallback: without a CUDA device every call fails with
 * KDBGPU_ERR_CUDA and kdbgpu_last_error() says why.
 *
 * Threading: any thread may call; searches on one handle run concurrently (up to 4 batches in
 * flight), calls that change the mirror are exclusive.  A call leaves the handle's device as the
 * calling thread's current CUDA device.
 */
#ifndef KEKTORDB_GPU_H
#define KEKTORDB_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDBGPU_OK 0
#define KDBGPU_ERR_INVALID (-1)  /* bad argument                                   */
#define KDBGPU_ERR_CUDA (-2)     /* CUDA runtime / no device                       */
#define KDBGPU_ERR_NOMEM (-3)    /* device or host allocation failed               */
#define KDBGPU_ERR_STATE (-4)    /* call order (e.g. search before graph upload)   */
#define KDBGPU_ERR_OVERFLOW (-5) /* a per-query candidate heap exceeded its bound  */

/* distance.DistanceMetric (pkg/core/distance/distance_go.go:34-39) */
#define KDBGPU_METRIC_L2 0     /* "euclidean": squared, no sqrt (:57-68) */
#define KDBGPU_METRIC_COSINE 1 /* "cosine": 1 - dot on unit vectors (:122-128) */

/* distance.PrecisionType (distance_go.go:41-46).  The reference offers float16 for "euclidean"
 * only and int8 for "cosine" only (float16Funcs / int8Funcs, :139-146). */
#define KDBGPU_PRECISION_F32 0
#define KDBGPU_PRECISION_F16 1  /* rows = float16 bits (uint16), squaredEuclideanGoFloat16 (:93-105)      */
#define KDBGPU_PRECISION_INT8 2 /* rows = int8 + per-row f32 norm, dotProductGoInt8 (:108-118) scaled as  */
                                /* in searchLayerUnlocked (hnsw_index.go:2398-2449)                       */

typedef struct kdbgpu_index kdbgpu_index;

/* Exact per-batch counters, for the roofline accounting of SURVEY.md §8(d). */
typedef struct {
  uint64_t dist_evals; /* E: query x stored-vector distance evaluations            */
  uint64_t hops;       /* H: candidate expansions (adjacency rows read)            */
  uint64_t hops_l0;    /* of which on level 0 (rows of 2M ids; the rest M ids)      */
  float kernel_ms;     /* device time of the traversal kernel(s), CUDA events      */
  float total_ms;      /* device time of the whole call incl. H2D / D2H copies     */
  uint32_t heap_pass_queries; /* queries re-answered by the heap pass after a distance tie in the
                                 fast pass (kdbgpu_set_fast_path); 0 when the fast pass is off */
} kdbgpu_stats;

/* ---- library ---------------------------------------------------------------------- */
int kdbgpu_device_count(void);
/* Message of the last failing call on this thread ("" if none).  Never NULL. */
const char *kdbgpu_last_error(void);
/* "kektordb_gpu <version> sm_100a" */
const char *kdbgpu_version(void);

/* ---- lifecycle: stands in for hnsw.New / Index.Close (hnsw_index.go:138, :3533) ------ */
/* One handle mirrors one hnsw.Index on one GPU.  m is the index's M (mMax0 = 2*m, :149),
 * capacity the highest internal id the mirror can hold. */
int kdbgpu_index_create(int device, int dim, int metric, int m, uint32_t capacity, kdbgpu_index **out);
/* hnsw.New with its precision argument (hnsw_index.go:138).  A float16 / int8 handle keeps the rows
 * in their stored form (2 / 1 bytes per element: the traversal kernel is HBM-bound, so bytes per row
 * are what it pays for) and answers kdbgpu_search_batch / kdbgpu_distance_batch with the reference's
 * arithmetic for that precision, kdbgpu_add_batch builds with that precision's distances (as DB.Compress
 * does through AddBatch, pkg/core/core.go:1210-1270); the flat scan is float32-only (BruteForceIndex). */
int kdbgpu_index_create_ex(int device, int dim, int metric, int precision, int m, uint32_t capacity,
                           kdbgpu_index **out);
int kdbgpu_index_destroy(kdbgpu_index *);
int kdbgpu_index_precision(const kdbgpu_index *);
/* Quantizer.AbsMax of an int8 index (pkg/core/distance/quantizer.go:19-22), as trained by
 * Quantizer.Train (:49-125) on the Go side.  Needed before float32 rows or queries are converted. */
int kdbgpu_set_quantizer(kdbgpu_index *, float abs_max);
/* Quantizer.Train (quantizer.go:49-125) itself, for hosts that would rather not sort on the CPU: n
 * float32 rows [n][dim] (host memory, or device memory with row_stride in floats), the reference's
 * stride sample, AbsMax = the 99.9th-percentile |value| (exact rank select on the device).  Sets the
 * handle's quantizer and returns it in *abs_max (may be NULL).  This is what TrainQuantizer does in
 * DB.Compress before the int8 rows are inserted (pkg/core/core.go:1224-1232). */
int kdbgpu_train_quantizer(kdbgpu_index *, const float *rows, uint32_t n, float *abs_max);
int kdbgpu_train_quantizer_device(kdbgpu_index *, const float *d_rows, size_t row_stride, uint32_t n, float *abs_max);

/* ---- staging: the GPU mirror of Index.nodes / Node.vec (hnsw_node.go:13-39) ----------- */
/* Rows exactly as the reference stores them (already unit-normalised for cosine,
 * hnsw_index.go:485-493); row i is internal id first_id + i.  Source of the bytes in the
 * reference: VectorArena.GetBytes (pkg/storage/mmap/arena.go:378). */
int kdbgpu_upload_vectors(kdbgpu_index *, uint32_t first_id, uint32_t count, const float *rows);
/* On a float16 / int8 handle the two calls above take float32 rows and convert them on the device as
 * Add / AddBatch do (float16.Fromfloat32, Quantizer.Quantize + computeInt8Norm; hnsw_index.go:497-520,
 * :1553-1577, :3371-3377).  kdbgpu_upload_rows_raw takes the rows already in stored form — the bytes of
 * VectorArena.GetBytes for that precision ([count][dim] float32 / uint16 / int8, no padding). */
int kdbgpu_upload_rows_raw(kdbgpu_index *, uint32_t first_id, uint32_t count, const void *rows);
int kdbgpu_download_rows_raw(kdbgpu_index *, uint32_t first_id, uint32_t count, void *rows);
/* int8: quantizedNorms[first_id .. first_id+count) (hnsw_index.go:87, :3371-3377) */
int kdbgpu_download_norms(kdbgpu_index *, uint32_t first_id, uint32_t count, float *norms);
/* Arena -> HBM staging: the reference keeps the stored rows in memory-mapped chunk files
 * (pkg/storage/mmap/arena.go: <dir>/arena_%04d.bin, 64 MiB each, 64-byte header {magic 0x4B414F4E,
 * version 1, dim, precision}, then (64 MiB - 64) / vectorSize rows; logical id -> physical slot through
 * ArenaState.SlotTable, 0xFFFFFFFF = unallocated; :14-19, :121-152, :307-376, :378-444).  These calls
 * move whole chunks to the device at DMA rate and place every row at its logical id there, instead of
 * VectorArena.GetBytes + a copy per row.  slot_table is ArenaState.SlotTable (indexed by internal id,
 * table_len entries; NULL = sequential slots, id i in slot i-1, with table_len = highest id + 1).
 * Header magic / version / dim / precision are validated as addChunk does (:346-364).
 *   kdbgpu_arena_load_dir     reads the chunk files of `dir` itself (cold start);
 *   kdbgpu_arena_stage_chunk  takes one chunk the host already has mapped (header included).
 * *rows_staged (may be NULL) receives the number of rows placed. */
/* Host-only look at an arena directory: dim / precision (0 f32, 1 f16, 2 int8, = KDBGPU_PRECISION_*) from the
 * chunk headers, number of chunk files, rows per chunk.  Any out pointer may be NULL.  Needs no device. */
int kdbgpu_arena_probe(const char *dir, uint32_t *dim, int *precision, uint32_t *n_chunks, uint32_t *vecs_per_chunk);
int kdbgpu_arena_load_dir(kdbgpu_index *, const char *dir, const uint32_t *slot_table, uint32_t table_len,
                          uint64_t *rows_staged);
int kdbgpu_arena_stage_chunk(kdbgpu_index *, uint32_t chunk_id, const void *chunk, size_t chunk_bytes,
                             const uint32_t *slot_table, uint32_t table_len, uint32_t *rows_staged);
/* kdbgpu_arena_stage_chunk page-locks the caller's mapping in place for the copy (cudaHostRegister) whenever the
 * chunk pointer is page aligned — an mmap always is — so that the DMA engine reads the mapping directly; this
 * counts the calls that did (0 for malloc'ed buffers or when registration is unavailable). */
uint64_t kdbgpu_arena_chunks_registered(const kdbgpu_index *);
/* Same, from device memory on the handle's device (row_stride in floats). */
int kdbgpu_upload_vectors_device(kdbgpu_index *, uint32_t first_id, uint32_t count, const float *d_rows,
                                 size_t row_stride);
/* Topology snapshot (what SnapshotData() exposes, hnsw_index.go:3064): n = nodeCounter;
 * levels[i] = len(Connections)-1 of node i or -1 for a nil slot; node i owns rows
 * node_row[i] .. node_row[i+1]-1 (level 0 first); row r lists nbrs[row_off[r] .. row_off[r+1]-1]
 * in the reference's order.  entry / max_level = entrypointID / maxLevel.
 * The rows are padded to their fixed degree (2M / M) and nil neighbours dropped on the device, slice by slice.  Per-node
 * errors (level, row count, entry point) are reported before the mirror is touched; a row with more than 2M / M live
 * neighbours or offsets out of order is found while the rows are written and leaves the mirror WITHOUT a graph
 * (searches return KDBGPU_ERR_STATE until a kdbgpu_set_graph succeeds). */
int kdbgpu_set_graph(kdbgpu_index *, uint32_t n, const int32_t *levels, const uint64_t *node_row,
                     const uint64_t *row_off, const uint32_t *nbrs, uint32_t entry, int max_level);
/* The same topology through a flat binary sidecar file, so that a 1 M - 10 M-node graph does not have to be rebuilt
 * in Go slices at every start (the reference keeps it inside its gob-encoded .kdb snapshot, pkg/core/core.go:177-306
 * / SnapshotData hnsw_index.go:3064-3150, which only Go can decode).  Layout (little endian, sections 8-byte
 * aligned): 64-byte header {u32 magic 0x4742444B, u32 version 1, u32 n, u32 entry, i32 max_level, u32 m, u64 n_rows,
 * u64 n_edges}, then levels i32[n+1], node_row u64[n+2], row_off u64[n_rows+1], nbrs u32[n_edges] — the arguments of
 * kdbgpu_set_graph.
 *   kdbgpu_graph_file_write  writes one from host arrays (atomically: temp file + rename); host only;
 *   kdbgpu_save_graph_file   writes the mirror's current topology (after kdbgpu_add_batch / a refresh);
 *   kdbgpu_graph_file_probe  header fields of a file; host only, any out pointer may be NULL;
 *   kdbgpu_set_graph_file    maps the file and stages it exactly as kdbgpu_set_graph would. */
int kdbgpu_graph_file_write(const char *path, uint32_t n, int m, const int32_t *levels, const uint64_t *node_row,
                            const uint64_t *row_off, const uint32_t *nbrs, uint32_t entry, int max_level);
int kdbgpu_save_graph_file(kdbgpu_index *, const char *path);
int kdbgpu_graph_file_probe(const char *path, uint32_t *n, int *m, uint64_t *n_rows, uint64_t *n_edges, uint32_t *entry,
                            int *max_level);
int kdbgpu_set_graph_file(kdbgpu_index *, const char *path);
/* Incremental refresh — follow the CPU index without re-staging everything:
 *   kdbgpu_register_nodes  ids first_id .. first_id+count-1 (= nodeCounter+1 onwards) come to exist with
 *                          levels[i] = len(Connections)-1 and empty rows (Add phase 1, hnsw_index.go:559-655;
 *                          -1 = an id that stays nil); their vectors go through kdbgpu_upload_vectors / _rows_raw;
 *   kdbgpu_patch_rows      node ids[i]'s Connections[row_levels[i]] := nbrs[row_off[i] .. row_off[i+1]) — the rows
 *                          Add's forward / reverse links (:717-783), Vacuum's reconnectNode and Refine's commits
 *                          (optimizer.go:195-222, :288-468) rewrite;
 *   kdbgpu_remove_nodes    nodes[id] = nil (Vacuum's physical cleanup, optimizer.go:252-274; it has already
 *                          rewired every live row that pointed at them);
 *   kdbgpu_set_entry       entrypointID / maxLevel (:793-801, optimizer.go:231-249; max_level -1 = empty graph).
 * Each call takes the handle exclusively and lets the searches in flight finish on the old state first. */
int kdbgpu_register_nodes(kdbgpu_index *, uint32_t first_id, uint32_t count, const int32_t *levels);
int kdbgpu_patch_rows(kdbgpu_index *, uint32_t count, const uint32_t *ids, const int32_t *row_levels,
                      const uint64_t *row_off, const uint32_t *nbrs);
int kdbgpu_remove_nodes(kdbgpu_index *, uint32_t count, const uint32_t *ids);
int kdbgpu_set_entry(kdbgpu_index *, uint32_t entry, int max_level);
/* The staleness policy on top of those calls (csrc/refresher.cpp).  Each of them takes the handle exclusively and
 * drains the searches in flight, so a host that mirrored every single Add would starve its own queries.  A
 * refresher QUEUES the changes — adjacency rows per (node, level) with last-write-wins, new nodes with their stored
 * vector, removals, Node.Deleted flips, the entry point — and applies them in ONE exclusive section when the queue
 * holds max_pending_rows rows, when the oldest queued change is max_lag_ms old (a background thread watches the
 * clock; 0 = no clock, flush by size / by call only), or when kdbgpu_refresher_flush is called.  Between flushes
 * searches see the mirror as of the last flush (bounded staleness: at most max_lag_ms / max_pending_rows behind).
 *   add_node      Add phase 1 (hnsw_index.go:559-655): id (> the mirror's count; skipped ids stay nil), level =
 *                 len(Connections)-1, row_raw = the vector in stored form (dim x 4 / 2 / 1 bytes, VectorArena bytes)
 *   set_row       Connections[level] of node id := nbrs (forward / reverse links :717-783, reconnectNode
 *                 optimizer.go:195-222, Refine :288-468)
 *   remove_node   nodes[id] = nil (optimizer.go:252-274);  set_deleted  Node.Deleted (:2303-2336)
 *   set_entry     entrypointID / maxLevel (:793-801, optimizer.go:231-249)
 * destroy flushes what is queued (no other call may race it).  Errors of a background flush are kept in
 * stats.last_error.  A flush that fails part-way (device out of memory, an id the mirror already holds) has consumed
 * its batch: the mirror is then behind the CPU index by those changes and must be re-staged (kdbgpu_set_graph /
 * kdbgpu_set_graph_file) before it is trusted again — the Go shim does exactly that when last_error is set. */
typedef struct kdbgpu_refresher kdbgpu_refresher;
typedef struct {
  uint64_t pending_rows, pending_nodes;
  uint64_t flushes, flushes_by_rows, flushes_by_lag, flushes_by_call;
  uint64_t rows_queued, rows_applied; /* queued minus applied = rewrites coalesced by last-write-wins */
  uint64_t nodes_applied;
  float oldest_pending_ms, last_flush_ms;
  int last_error;
} kdbgpu_refresher_stats_t;
int kdbgpu_refresher_create(kdbgpu_index *, uint32_t max_pending_rows, uint32_t max_lag_ms, kdbgpu_refresher **out);
int kdbgpu_refresher_destroy(kdbgpu_refresher *);
int kdbgpu_refresher_add_node(kdbgpu_refresher *, uint32_t id, int level, const void *row_raw);
int kdbgpu_refresher_set_row(kdbgpu_refresher *, uint32_t id, int level, const uint32_t *nbrs, uint32_t count);
int kdbgpu_refresher_remove_node(kdbgpu_refresher *, uint32_t id);
int kdbgpu_refresher_set_deleted(kdbgpu_refresher *, uint32_t id, int is_deleted);
int kdbgpu_refresher_set_entry(kdbgpu_refresher *, uint32_t entry, int max_level);
int kdbgpu_refresher_flush(kdbgpu_refresher *);
int kdbgpu_refresher_stats(kdbgpu_refresher *, kdbgpu_refresher_stats_t *out);
/* Node.Deleted flags as a dense bitset over ids (bit i of word i/64); NULL clears all. */
int kdbgpu_set_deleted(kdbgpu_index *, const uint64_t *bitset, size_t words);

/* ---- query: stands in for (*Index).SearchWithScores (hnsw_index.go:343-468) ----------- */
/* nq independent searches in one call (the Go shim's micro-batcher forms the batch).
 *   queries   [nq][dim] raw float32; for cosine the library normalises a copy exactly as
 *             normalize() does (hnsw_index.go:3030-3045).
 *   k         results wanted per query; ef_search as passed to SearchWithScores
 *             (ef = max(ef_search, k), :2377-2380; the needsRefine boost :387-399 is applied
 *             by the caller).
 *   allow     NULL = nil allow-list.  Otherwise a dense bitset over internal ids with the same
 *             membership as the *roaring.Bitmap from DB.FindIDsByFilter (pkg/core/core.go:1766),
 *             shared by the whole batch; semantics of :436-447, :2480-2485, :2545-2549.
 *   out_ids / out_scores  [nq][k], ascending distance; rows are zero-filled past out_counts[q].
 * A query whose search "fails" in the reference (nil entry, empty level result) yields count 0,
 * as SearchWithScores does (:355-359). */
int kdbgpu_search_batch(kdbgpu_index *, const float *queries, uint32_t nq, int k, int ef_search,
                        const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                        uint32_t *out_counts, kdbgpu_stats *stats);
/* Same with every buffer resident on the handle's device; enqueued on `stream` (a cudaStream_t,
 * NULL = the handle's own stream) without host synchronisation.  d_allow may be NULL;
 * allow_first_id is the smallest member of the allow-list (ignored when d_allow is NULL). */
int kdbgpu_search_batch_device(kdbgpu_index *, const float *d_queries, uint32_t nq, int k, int ef_search,
                               const uint64_t *d_allow, size_t allow_words, uint32_t allow_first_id,
                               uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts,
                               void *stream);

/* ---- the literal inner-loop hook: DistanceFuncF32 over a candidate list ---------------- */
/* out[i] = distFn(query, node ids[i]) — the closure at hnsw_index.go:2393-2396 evaluated for a
 * whole neighbour list in one launch.  `query` must already be prepared (normalised for
 * cosine), as inside searchLayerUnlocked. */
int kdbgpu_distance_batch(kdbgpu_index *, const float *query, const uint32_t *ids, uint32_t n, double *out);

/* ---- flat: BruteForceIndex.SearchWithScores (pkg/core/vector_index.go:104-162) ---------- */
/* Exhaustive scan of the staged rows (deleted ids skipped, allow-list applied as a filter).
 *   mode 0  reference arithmetic: sum of float64(q_i - x_i)^2 on the RAW query, any metric.
 *   mode 1  exact float64 distance under the index metric (cosine: query normalised as in
 *           searchInternal) — the ground truth used for recall.
 * Ties are returned in ascending id order (the reference's order among ties is unspecified).
 *
 * mode | KDBGPU_FLAT_PREFILTER: same results, bit for bit, found through a bf16 tensor-core
 * (tcgen05) Q x K^T pass that only NOMINATES rows; the nominated rows are re-scored in the float64
 * arithmetic above and a per-query certificate proves no other row can enter the top k (a query
 * whose certificate does not close is answered by the exhaustive scan).  stats->dist_evals = exact
 * float64 evaluations, stats->hops = queries answered by the exhaustive scan, stats->kernel_ms =
 * device time of all kernels of the call (copies excluded), stats->hops_l0 = device time of the
 * tensor-core passes alone, in nanoseconds. */
#define KDBGPU_FLAT_PREFILTER 0x10
int kdbgpu_flat_search_batch(kdbgpu_index *, const float *queries, uint32_t nq, int k, int mode,
                             const uint64_t *allow, size_t allow_words, uint32_t *out_ids,
                             double *out_scores, uint32_t *out_counts, kdbgpu_stats *stats);

/* Validation hook for the pre-filter (tests only): the approximate scores the tensor-core pass
 * assigns, out_scores [nq][n] (row r = id r+1; exact distance = score + |q|^2 for L2 / mode 0,
 * = score + 1 for cosine mode 1; +inf for rows that may not be nominated), and out_bound [nq], the
 * certified bound on |approximate - exact| the thresholds are built from.  nq * n <= 2^30. */
int kdbgpu_flat_prefilter_scores(kdbgpu_index *, const float *queries, uint32_t nq, int mode, float *out_scores,
                                 float *out_bound);

/* ---- multi-GPU: merge of per-shard top-k (id-range shards) -------------- */
/* d_ids / d_scores: [n_shards][nq][k] gathered candidates (global ids), d_counts [n_shards][nq].
 * Writes the k best per query, ascending (distance, id).  Device buffers, async on `stream`. */
int kdbgpu_merge_topk_device(kdbgpu_index *, int n_shards, uint32_t nq, int k, const uint32_t *d_ids,
                             const double *d_scores, const uint32_t *d_counts, uint32_t *d_out_ids,
                             double *d_out_scores, uint32_t *d_out_counts, void *stream);

/* ---- multi-GPU: id-range shard groups -------------------------------------------
 * The corpus is split by contiguous internal-id range over G GPUs; shard g is a self-contained hnsw.Index
 * over its ids (an HNSW graph cannot be cut by id range), mirrored by one kdbgpu_index on its own GPU whose
 * LOCAL ids 1..n stand for the GLOBAL ids id_base+1 .. id_base+n.  A group answers
 * idx.SearchWithScores(query, k, allowList, efSearch) (pkg/engine/ops.go:1006) over the whole corpus: every
 * shard runs the unchanged traversal (its epilogue writes global ids), ONE exchange moves the per-shard
 * top-k — scores, ids, counts, counters and error flag packed in one buffer per shard — and a merge kernel
 * keeps the k best by (distance, id).  The reference has no sharded mode; the parity oracle is "G reference
 * indexes + exact merge" (tests/test_gpu_shard.py).
 *
 * Two ways to form a group:
 *   kdbgpu_shard_group_create_local  every shard lives in THIS process (the Go host: one process, G GPUs).
 *                                    The exchange is a peer copy of each shard's packed result into the merge
 *                                    GPU's gather buffer over NVLink (cudaMemcpyPeerAsync); no NCCL.  Shards
 *                                    may share a device (tests).
 *   kdbgpu_shard_group_create_rank   one shard per process / rank (torchrun-style, one process per GPU).  The
 *                                    exchange is one ncclAllGather of the packed results per batch; NCCL
 *                                    (libnccl.so.2) is loaded on first use.  Rank 0 obtains the 128-byte id with
 *                                    kdbgpu_shard_unique_id and hands it to the other ranks by any channel.
 *                                    Calls on a rank group are COLLECTIVE: every rank issues the same searches
 *                                    in the same order, and every rank receives the merged result.
 * Up to 4 batches are in flight per group: the exchange + merge of batch i (on a high-priority stream) overlap the
 * traversal of batch i+1. */
typedef struct kdbgpu_shard_group kdbgpu_shard_group;
#define KDBGPU_SHARD_ID_BYTES 128
typedef struct {
  uint64_t dist_evals, hops, hops_l0; /* summed over all shards                                       */
  float traversal_ms;                 /* slowest local shard: H2D + prep + traversal (device time)    */
  float exchange_ms;                  /* all-gather / peer copies                                     */
  float merge_ms;                     /* merge kernel                                                 */
  float total_ms;                     /* first H2D .. merged result on the host                       */
  uint32_t n_shards;
} kdbgpu_shard_stats;
int kdbgpu_shard_unique_id(unsigned char id[KDBGPU_SHARD_ID_BYTES]);
int kdbgpu_shard_group_create_rank(kdbgpu_index *local, int rank, int world, const unsigned char id[KDBGPU_SHARD_ID_BYTES],
                                   uint32_t id_base, kdbgpu_shard_group **out);
int kdbgpu_shard_group_create_local(kdbgpu_index *const *shards, int n_shards, const uint32_t *id_bases,
                                    kdbgpu_shard_group **out);
/* Waits for the batches in flight, then frees the group (not the indexes). */
int kdbgpu_shard_group_destroy(kdbgpu_shard_group *);
int kdbgpu_shard_group_size(const kdbgpu_shard_group *); /* G */
/* SearchWithScores over the sharded corpus; arguments as kdbgpu_search_batch.  `allow` is a dense bitset over
 * GLOBAL ids; each shard applies its own slice with the semantics of hnsw_index.go:436-447 (a shard whose slice
 * is empty contributes nothing, as an index searched with an empty allow-list returns []).  out_ids are global. */
int kdbgpu_shard_search_batch(kdbgpu_shard_group *, const float *queries, uint32_t nq, int k, int ef_search,
                              const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, kdbgpu_shard_stats *stats);
/* The same call split in two so that ONE thread can keep several batches in flight (collectives stay in issue
 * order): submit queues the H2D copy, traversals, exchange, merge and D2H copy and returns a ticket; wait blocks
 * until that batch's results are in the caller's buffers.  `queries` must stay valid until wait returns. */
typedef struct kdbgpu_shard_ticket kdbgpu_shard_ticket;
int kdbgpu_shard_search_submit(kdbgpu_shard_group *, const float *queries, uint32_t nq, int k, int ef_search,
                               const uint64_t *allow, size_t allow_words, kdbgpu_shard_ticket **ticket);
int kdbgpu_shard_search_wait(kdbgpu_shard_ticket *, uint32_t *out_ids, double *out_scores, uint32_t *out_counts,
                             kdbgpu_shard_stats *stats);
/* Device-resident form: queries and outputs on the device of the group's first local shard; everything is queued
 * without host synchronisation and `stream` (a cudaStream_t, NULL = an internal one) is made to wait for the
 * merged result.  Errors of the launch (candidate-heap overflow) surface in kdbgpu_shard_sync. */
int kdbgpu_shard_search_batch_device(kdbgpu_shard_group *, const float *d_queries, uint32_t nq, int k, int ef_search,
                                     uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream);
/* Waits for every batch queued by the device-resident form; stats (may be NULL) = the last batch's counters and
 * device times.  Returns KDBGPU_ERR_OVERFLOW if any shard's traversal overflowed since the last sync. */
int kdbgpu_shard_sync(kdbgpu_shard_group *, kdbgpu_shard_stats *stats);
/* BruteForceIndex.SearchWithScores (pkg/core/vector_index.go:104-162) over the sharded corpus: every shard scans
 * its rows (mode as kdbgpu_flat_search_batch, KDBGPU_FLAT_PREFILTER included), same exchange and merge; the
 * merged answer equals the unsharded scan bit for bit (ties by id).  No host round trip between scan and merge. */
int kdbgpu_shard_flat_search_batch(kdbgpu_shard_group *, const float *queries, uint32_t nq, int k, int mode,
                                   const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                                   uint32_t *out_counts, kdbgpu_shard_stats *stats);

/* ---- graph construction on the device (extension; the reference builds on the CPU) --------- */
/* (*Index).AddBatch (hnsw_index.go:1466, addBatchInternal :1479-2088) for `count` raw vectors that
 * receive ids n+1 .. n+count.  level_draws[i] is the rand.Float64() of randomLevel (:2616-2625).
 * ef_const <= 0 -> 200 (efConstruction default); AddBatchFast passes max(2M, 40) (:1470-1476).
 * While the index holds fewer than ef_const nodes the whole batch goes through sequential single
 * Adds (:1502-1513, Add :472-809), as in the reference.  Deterministic. */
int kdbgpu_add_batch(kdbgpu_index *, uint32_t count, const float *rows, const double *level_draws, int ef_const);
/* Same with the rows resident on the handle's device (row_stride in floats). */
int kdbgpu_add_batch_device(kdbgpu_index *, uint32_t count, const float *d_rows, size_t row_stride,
                            const double *level_draws, int ef_const);
/* Read the topology back (same layout as kdbgpu_set_graph; levels [n+1], node_row [n+2],
 * row_off [n_rows+1], nbrs [n_edges]) and the stored rows ([count][dim]). */
int kdbgpu_get_graph_sizes(kdbgpu_index *, uint32_t *n, uint64_t *n_rows, uint64_t *n_edges, uint32_t *entry,
                           int *max_level);
int kdbgpu_get_graph(kdbgpu_index *, int32_t *levels, uint64_t *node_row, uint64_t *row_off, uint32_t *nbrs);
int kdbgpu_download_vectors(kdbgpu_index *, uint32_t first_id, uint32_t count, float *rows);

/* ---- introspection -------------------------------------------------------------------- */
/* Counters (dist_evals, hops, hops_l0) of the most recent traversal launch on this handle;
 * synchronises the device.  For callers of the *_device entry points: returns KDBGPU_ERR_OVERFLOW (counters
 * still filled) when a query of that launch exceeded the candidate-heap bound and was answered with count 0. */
int kdbgpu_last_search_stats(kdbgpu_index *, kdbgpu_stats *stats);
int kdbgpu_index_device(const kdbgpu_index *);
int kdbgpu_index_dim(const kdbgpu_index *);
int kdbgpu_index_m(const kdbgpu_index *);
uint32_t kdbgpu_index_count(const kdbgpu_index *);   /* n of the last kdbgpu_set_graph        */
uint64_t kdbgpu_index_device_bytes(const kdbgpu_index *);
/* Resident CTAs the traversal kernel runs with for (k, ef) — queries in flight per launch. */
int kdbgpu_search_concurrency(kdbgpu_index *, int k, int ef_search);
/* Reserves every launch workspace (visited bitsets, staging, pinned result buffers) for batches of up
 * to nq queries of this (k, ef_search) shape, so the first searches do not allocate.  Optional. */
int kdbgpu_prepare_search(kdbgpu_index *, uint32_t nq, int k, int ef_search);

/* ---- tuning hook (tests / benchmarks only; not bound by the Go shim) --------------------- */
/* Shape of the traversal kernel (one warp per query): bulk-copy row slots per query, candidate-heap
 * entries kept in shared memory, cap on resident query-warps per SM (0 = no cap).  A value <= 0
 * (< 0 for the cap) keeps the current setting.  Results never depend on the shape. */
int kdbgpu_set_tuning(kdbgpu_index *, int slots, int cand_smem, int max_ctas_per_sm);
/* Slot count of a launch that finds no other batch of the handle in flight (0 = same as `slots`): with more rows in
 * flight per query a lone batch finishes sooner (batch latency), with fewer a stream of overlapping batches moves
 * more queries per second.  The defaults are per precision. */
int kdbgpu_set_idle_slots(kdbgpu_index *, int slots_idle);
/* The reference's candidate heap is unbounded; ours holds cand_smem entries in shared memory and spills up to
 * spill_entries (default 32768) to HBM per query.  A query that would need more returns count 0 and the call
 * fails with KDBGPU_ERR_OVERFLOW (kdbgpu_search_batch: at once; the *_device forms: from
 * kdbgpu_last_search_stats / kdbgpu_shard_sync).  This hook sets the spill bound (tests force the overflow with it;
 * a host may raise it). */
int kdbgpu_set_candidate_bound(kdbgpu_index *, uint32_t spill_entries);
/* The traversal answers a batch in two passes: a fast pass that keeps the candidate / result queues as one
 * sorted list in registers (valid while all distances a query meets are distinct — any priority queue then
 * pops what the reference's binary heaps pop), and the heap pass (hnsw_heap.go restated exactly) over the
 * queries where a distance tie made a pop / eviction / the final order ambiguous, and over everything when
 * soft-deleted nodes exist or ef > 128.
 * Results are identical either way.  on = 0: heap pass only; 1 (default): the fast pass where ties are
 * practically absent (int8: distances are full-precision float64 ratios); 2: the fast pass for every
 * precision (float32 / float16 distances are float32 sums: 2-3 % of the queries at 1 M x 768 meet a relevant
 * tie and are re-run, and the pass measures slower than the heaps there — kept for testing). */
int kdbgpu_set_fast_path(kdbgpu_index *, int on);

/* ---- micro-batcher: the reference's call shape on top of the batched entry point ----------------
 * Every search in the reference is one blocking call per query from its own goroutine
 * (idx.SearchWithScores(query, k, allowList, efSearch), pkg/engine/ops.go:1006, :1296).  A batcher keeps
 * exactly that shape for the cgo shim — kdbgpu_batcher_search takes ONE query, blocks, returns its
 * result — and forms device batches underneath: callers with the same (k, ef_search, allow-list) are
 * grouped; a group is dispatched at once while the device is idle, and otherwise when it reaches
 * max_batch queries or max_wait_us has passed since its first query.  Any number of threads may call.
 * A failed search returns the error code and an empty result (hnsw_index.go:355-359).  Results are the
 * ones kdbgpu_search_batch gives for that query alone: queries of a batch are independent. */
typedef struct kdbgpu_batcher kdbgpu_batcher;
/* The batch executor a batcher fronts; kdbgpu_batcher_create uses kdbgpu_search_batch on one handle,
 * kdbgpu_batcher_create_fn takes any function of that shape (a sharded multi-GPU search, a test double). */
typedef int (*kdbgpu_batch_fn)(void *ctx, const float *queries, uint32_t nq, int k, int ef_search,
                               const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                               uint32_t *out_counts);
typedef struct {
  uint64_t queries, batches, max_batch_seen;
  uint64_t dispatched_idle;     /* batches sent at once because nothing was in flight */
  uint64_t dispatched_full;     /* batches sent because they reached max_batch         */
  uint64_t dispatched_deadline; /* batches sent because max_wait_us passed             */
} kdbgpu_batcher_stats_t;
int kdbgpu_batcher_create(kdbgpu_index *, uint32_t max_batch, uint32_t max_wait_us, kdbgpu_batcher **out);
int kdbgpu_batcher_create_fn(kdbgpu_batch_fn fn, void *ctx, int dim, uint32_t max_batch, uint32_t max_wait_us,
                             kdbgpu_batcher **out);
/* Waits for the callers inside to finish, then frees the batcher (not the index). */
int kdbgpu_batcher_destroy(kdbgpu_batcher *);
/* (*Index).SearchWithScores for one query: out_ids / out_scores hold k entries, *out_count <= k. */
int kdbgpu_batcher_search(kdbgpu_batcher *, const float *query, int k, int ef_search, const uint64_t *allow,
                          size_t allow_words, uint32_t *out_ids, double *out_scores, uint32_t *out_count);
int kdbgpu_batcher_stats(kdbgpu_batcher *, kdbgpu_batcher_stats_t *out);
/* A batcher over a shard group (kdbgpu_shard_search_batch as the executor): one blocking / asynchronous call per
 * query in front of G GPUs.  Local groups only (a rank group's calls are collective). */
int kdbgpu_batcher_create_group(kdbgpu_shard_group *, int dim, uint32_t max_batch, uint32_t max_wait_us,
                                kdbgpu_batcher **out);
/* Asynchronous form — for hosts whose callers must not block inside the C call (a Go shim: a goroutine blocked in
 * cgo pins an OS thread, so thousands of in-flight searches would pin thousands of threads):
 *   kdbgpu_batcher_submit  copies the query into the batcher's staging (the caller's buffers are not retained: cgo
 *                          pointer rules) and returns a ticket at once.  filter_id != 0 names an allow-list
 *                          registered with kdbgpu_batcher_register_filter (no per-query hashing / copying of MBs of
 *                          bitset); otherwise allow / allow_words as in kdbgpu_batcher_search.
 *   kdbgpu_batcher_poll    blocks up to timeout_us for finished queries and returns up to max_tickets of their
 *                          tickets — ONE dispatcher thread (goroutine) calls it in a loop and wakes the waiters.
 *   kdbgpu_batcher_take    copies the result of a finished ticket into the caller's buffers (k entries each, the k of
 *                          the submit) and releases the ticket; returns that query's error code.  Blocks until the
 *                          query is finished if it is not yet.  Every ticket must be taken exactly once.
 * The batches are run by the batcher's own worker threads (KDBGPU_BATCHER_WORKERS, default 4). */
int kdbgpu_batcher_submit(kdbgpu_batcher *, const float *query, int k, int ef_search, const uint64_t *allow,
                          size_t allow_words, uint64_t filter_id, uint64_t *ticket);
int kdbgpu_batcher_poll(kdbgpu_batcher *, uint64_t *tickets, uint32_t max_tickets, uint32_t timeout_us, uint32_t *n_done);
int kdbgpu_batcher_take(kdbgpu_batcher *, uint64_t ticket, uint32_t *out_ids, double *out_scores, uint32_t *out_count);
int kdbgpu_batcher_register_filter(kdbgpu_batcher *, const uint64_t *allow, size_t allow_words, uint64_t *filter_id);
int kdbgpu_batcher_release_filter(kdbgpu_batcher *, uint64_t filter_id);

/* ---- host staging memory ------------------------------------------------------------------------- */
/* Page-locked host memory (cudaHostAlloc) for buffers that cross the boundary often — query batches, result
 * arrays: copies from / to it run at DMA rate and asynchronously.  Without a CUDA device the memory is ordinary
 * (aligned) host memory.  Release with kdbgpu_host_free (NULL is fine). */
int kdbgpu_host_alloc(void **out, size_t bytes);
void kdbgpu_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* KEKTORDB_GPU_H */
