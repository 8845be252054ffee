// shard.cu — id-range shard groups: the sharded form of SearchWithScores / the flat scan
// behind the C ABI.  Every shard is an ordinary kdbgpu_index (its own HNSW over its id range); a group
// runs the unchanged traversal on every shard, moves the per-shard results with ONE exchange per batch
// and merges them on the device:
//   rank groups   one process per GPU: ncclAllGather of the packed per-shard results (NCCL loaded with
//                 dlopen on first use, so single-GPU hosts never need it), merge on every rank;
//   local groups  all shards in this process: each shard's packed result is copied straight into the
//                 merge GPU's gather buffer over NVLink (cudaMemcpyPeerAsync), merge on that GPU.
// A batch occupies one of kSlots slots; the exchange + merge + D2H of a slot run on the slot's own
// high-priority stream, so they overlap the traversals of the following batches.
// The reference has no sharded mode (call site: pkg/engine/ops.go:1006); the parity oracle is
// "G reference-semantic indexes + exact merge by (distance, id)".
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <future>
#include <new>

#include "handle.h"

using namespace kdb;

namespace {

// ---- NCCL, loaded at run time ---------------------------------------------------------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string err;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("KDBGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
      api.err = dlerror();
    }
    if (!api.lib) return;
    bool ok = true;
    auto sym = [&](const char *name) -> void * {
      void *p = dlsym(api.lib, name);
      if (!p) {
        ok = false;
        api.err = std::string("missing symbol ") + name;
      }
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!ok) {
      dlclose(api.lib);
      api.lib = nullptr;
    }
  });
  return &api;  // callers test ->lib
}

#define NCCL_TRY(api, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) return set_error(KDBGPU_ERR_CUDA, "%s: %s", #expr, (api)->GetErrorString(_r)); \
  } while (0)

constexpr int kSlots = 4;
constexpr int kMaxShards = 64;

struct Member {
  kdbgpu_index *h = nullptr;
  uint32_t id_base = 0;
  int device = 0;
};

struct Slot {
  bool busy = false;
  bool used = false;          // ev_done has been recorded at least once
  bool device_form = false;   // last use was the device-resident form (its error flag has not been read yet)
  cudaStream_t xs = nullptr;  // exchange + merge + D2H, high priority, on the merge device
  cudaEvent_t ev_start = nullptr, ev_x0 = nullptr, ev_x1 = nullptr, ev_m = nullptr, ev_done = nullptr;
  std::vector<cudaEvent_t> ev_m0, ev_m1, ev_m2;  // per member: start, traversal done, result sent
  std::vector<DevBuf<unsigned char>> send;       // per member, on its device: the packed per-shard result
  std::vector<DevBuf<uint32_t>> allow_g;         // per member: the global allow-list as staged
  DevBuf<unsigned char> gather, outb;            // merge device
  unsigned char *h_out = nullptr;                // pinned
  size_t h_out_bytes = 0;
  uint32_t nq = 0;
  int k = 0;
  BlobLayout L{};
};

}  // namespace

struct kdbgpu_shard_group {
  bool rank_mode = false;
  std::vector<Member> members;  // the shards living in this process
  int n_shards = 0;             // G
  int rank = 0;                 // rank mode: index of the local shard
  int merge_device = 0;
  ncclComm_t comm = nullptr;
  Slot slots[kSlots];
  std::mutex mu;  // one submit at a time: collectives and slot use stay in issue order
  std::condition_variable cv;
  unsigned next_slot = 0;
  int last_device_slot = -1;
  int sticky_err = 0;
  cudaStream_t user_default = nullptr;  // device-resident form with stream == NULL
};

struct kdbgpu_shard_ticket {
  kdbgpu_shard_group *g;
  int slot;
};

namespace {

int first_set_in_range(const uint64_t *bits, size_t words, uint64_t lo, uint64_t hi, uint64_t *out) {
  // smallest set bit position p with lo <= p <= hi; returns 0 if none
  for (uint64_t p = lo; p <= hi;) {
    const size_t w = (size_t)(p >> 6);
    if (w >= words) return 0;
    uint64_t v = bits[w] >> (p & 63);
    if (v) {
      const uint64_t pos = p + (uint64_t)__builtin_ctzll(v);
      if (pos > hi) return 0;
      *out = pos;
      return 1;
    }
    p = ((uint64_t)w + 1) << 6;
  }
  return 0;
}

int slot_init(kdbgpu_shard_group *g, Slot &s) {
  DeviceGuard dg(g->merge_device);
  int lo = 0, hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = greatest priority (numerically lowest)
  CUDA_TRY(cudaStreamCreateWithPriority(&s.xs, cudaStreamNonBlocking, hi));
  for (cudaEvent_t *e : {&s.ev_start, &s.ev_x0, &s.ev_x1, &s.ev_m, &s.ev_done}) CUDA_TRY(cudaEventCreate(e));
  const size_t nm = g->members.size();
  s.ev_m0.assign(nm, nullptr);
  s.ev_m1.assign(nm, nullptr);
  s.ev_m2.assign(nm, nullptr);
  s.send.resize(nm);
  s.allow_g.resize(nm);
  for (size_t m = 0; m < nm; ++m) {
    DeviceGuard dm(g->members[m].device);
    CUDA_TRY(cudaEventCreate(&s.ev_m0[m]));
    CUDA_TRY(cudaEventCreate(&s.ev_m1[m]));
    CUDA_TRY(cudaEventCreate(&s.ev_m2[m]));
  }
  return KDBGPU_OK;
}

void slot_destroy(kdbgpu_shard_group *g, Slot &s) {
  for (size_t m = 0; m < g->members.size(); ++m) {
    DeviceGuard dm(g->members[m].device);
    if (m < s.send.size()) s.send[m].release();
    if (m < s.allow_g.size()) s.allow_g[m].release();
    for (auto *v : {&s.ev_m0, &s.ev_m1, &s.ev_m2})
      if (m < v->size() && (*v)[m]) cudaEventDestroy((*v)[m]);
  }
  DeviceGuard dg(g->merge_device);
  s.gather.release();
  s.outb.release();
  if (s.h_out) cudaFreeHost(s.h_out);
  for (cudaEvent_t e : {s.ev_start, s.ev_x0, s.ev_x1, s.ev_m, s.ev_done})
    if (e) cudaEventDestroy(e);
  if (s.xs) cudaStreamDestroy(s.xs);
}

// takes a free slot (blocks while all are busy); g->mu held by the caller through `lk`
int slot_acquire(kdbgpu_shard_group *g, std::unique_lock<std::mutex> &lk) {
  for (;;) {
    for (int i = 0; i < kSlots; ++i) {
      const int j = (int)((g->next_slot + (unsigned)i) % kSlots);
      if (!g->slots[j].busy) {
        g->slots[j].busy = true;
        g->next_slot = (unsigned)j + 1;
        return j;
      }
    }
    g->cv.wait(lk);
  }
}

void slot_release(kdbgpu_shard_group *g, int j) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->slots[j].busy = false;
  }
  g->cv.notify_all();
}

// a slot last used by the device-resident form still holds that batch's error flag: fold it in before reuse
int slot_fold_error(kdbgpu_shard_group *g, Slot &s) {
  if (!s.used || !s.device_form) return KDBGPU_OK;
  DeviceGuard dg(g->merge_device);
  CUDA_TRY(cudaEventSynchronize(s.ev_done));
  int err = 0;
  memcpy(&err, s.h_out + s.L.o_err, sizeof err);
  if (err && !g->sticky_err) g->sticky_err = err;
  s.device_form = false;
  return KDBGPU_OK;
}

int slot_reserve(kdbgpu_shard_group *g, Slot &s, uint32_t nq, int k) {
  s.nq = nq;
  s.k = k;
  s.L = blob_layout(nq, k);
  DeviceGuard dg(g->merge_device);
  CUDA_TRY(s.gather.reserve(s.L.bytes * (size_t)g->n_shards));
  CUDA_TRY(s.outb.reserve(s.L.bytes));
  if (s.h_out_bytes < s.L.bytes) {
    if (s.h_out) cudaFreeHost(s.h_out);
    s.h_out = nullptr;
    s.h_out_bytes = 0;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&s.h_out), s.L.bytes + s.L.bytes / 4, cudaHostAllocDefault));
    s.h_out_bytes = s.L.bytes + s.L.bytes / 4;
  }
  return KDBGPU_OK;
}

// where shard `index_in_group`'s packed result lands in the gather buffer
inline int shard_index(const kdbgpu_shard_group *g, size_t m) { return g->rank_mode ? g->rank : (int)m; }

// Queue one shard's part of a batch: queries -> workspace, preparation, traversal with its results packed into
// s.send[m] (global ids), and in a local group the peer copy of that buffer into the merge GPU's gather buffer.
// q_src is a host pointer, or a device pointer on q_src_device (>= 0) that `q_ready` (may be NULL) guards.
int enqueue_member(kdbgpu_shard_group *g, Slot &s, size_t m, const float *q_src, int q_src_device, cudaEvent_t q_ready,
                   uint32_t nq, int k, int ef, const uint64_t *allow, size_t allow_words) {
  Member &mb = g->members[m];
  kdbgpu_index *h = mb.h;
  DeviceGuard dg(mb.device);
  std::shared_lock<std::shared_mutex> lk(h->mu);
  if (!h->has_graph) return set_error(KDBGPU_ERR_STATE, "shard %d: kdbgpu_set_graph / kdbgpu_add_batch has not been called", shard_index(g, m));
  CUDA_TRY(s.send[m].reserve(s.L.bytes));
  const int wi = acquire_ws(h);
  kdbgpu_index::SearchWs &w = h->sws[wi];
  struct Release {
    kdbgpu_index *h;
    int wi;
    ~Release() { release_ws(h, wi); }
  } releaser{h, wi};
  cudaStream_t st = w.stream;
  CUDA_TRY(cudaStreamWaitEvent(st, w.done, 0));
  if (s.used) CUDA_TRY(cudaStreamWaitEvent(st, s.ev_done, 0));  // the slot's previous batch has left s.send / s.gather
  if (q_ready) CUDA_TRY(cudaStreamWaitEvent(st, q_ready, 0));
  CUDA_TRY(cudaEventRecord(s.ev_m0[m], st));
  unsigned char *blob = s.send[m].p;
  // the local slice of the allow-list; its smallest member is the smart entry point (hnsw_index.go:436-447)
  bool empty = h->max_level < 0 || h->n == 0;
  uint32_t allow_first = 0;
  const uint32_t *d_allow = nullptr;
  if (allow && !empty) {
    uint64_t pos = 0;
    if (!first_set_in_range(allow, allow_words, (uint64_t)mb.id_base + 1, (uint64_t)mb.id_base + h->n, &pos)) {
      empty = true;  // an index searched with an empty allow-list returns [] (:443-445)
    } else {
      allow_first = (uint32_t)(pos - mb.id_base);
      const size_t g32 = allow_words * 2;
      const size_t need32 = ((size_t)h->capacity + 1 + 31) / 32 + 2;
      CUDA_TRY(s.allow_g[m].reserve(g32));
      CUDA_TRY(w.allow.reserve(need32));
      CUDA_TRY(cudaMemcpyAsync(s.allow_g[m].p, allow, g32 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      CUDA_TRY(launch_slice_bits(s.allow_g[m].p, g32, mb.id_base, w.allow.p, need32, st));
      d_allow = w.allow.p;
    }
  }
  int rc = KDBGPU_OK;
  if (empty) {
    CUDA_TRY(cudaMemsetAsync(blob, 0, s.L.bytes, st));
  } else {
    CUDA_TRY(w.q_raw.reserve((size_t)nq * h->dim));
    if (q_src_device >= 0 && q_src_device != mb.device)
      CUDA_TRY(cudaMemcpyPeerAsync(w.q_raw.p, mb.device, q_src, q_src_device, (size_t)nq * h->dim * sizeof(float), st));
    else
      CUDA_TRY(cudaMemcpyAsync(w.q_raw.p, q_src, (size_t)nq * h->dim * sizeof(float), cudaMemcpyDefault, st));
    rc = prepare_queries(h, w, w.q_raw.p, nq, st);
    if (rc == KDBGPU_OK)
      rc = enqueue_search(h, w, w.q_prep.p, nq, k, ef, d_allow, allow_first, reinterpret_cast<uint32_t *>(blob + s.L.o_ids),
                          reinterpret_cast<double *>(blob + s.L.o_scores), reinterpret_cast<uint32_t *>(blob + s.L.o_counts),
                          st, reinterpret_cast<unsigned long long *>(blob + s.L.o_stats),
                          reinterpret_cast<int *>(blob + s.L.o_err), mb.id_base);
  }
  cudaError_t e = cudaEventRecord(s.ev_m1[m], st);
  if (rc == KDBGPU_OK && e == cudaSuccess && !g->rank_mode) {
    unsigned char *dst = s.gather.p + (size_t)m * s.L.bytes;
    if (mb.device == g->merge_device)
      e = cudaMemcpyAsync(dst, blob, s.L.bytes, cudaMemcpyDeviceToDevice, st);
    else
      e = cudaMemcpyPeerAsync(dst, g->merge_device, blob, mb.device, s.L.bytes, st);
  }
  if (e == cudaSuccess) e = cudaEventRecord(s.ev_m2[m], st);
  cudaError_t e2 = cudaEventRecord(w.done, st);
  if (rc) return rc;
  CUDA_TRY(e);
  CUDA_TRY(e2);
  return KDBGPU_OK;
}

// the exchange (rank groups: one all-gather; local groups: wait for the peer copies) and the merge, on the slot's
// own stream; the merged result is left in s.outb
int enqueue_exchange_merge(kdbgpu_shard_group *g, Slot &s) {
  DeviceGuard dg(g->merge_device);
  for (size_t m = 0; m < g->members.size(); ++m) CUDA_TRY(cudaStreamWaitEvent(s.xs, s.ev_m2[m], 0));
  CUDA_TRY(cudaEventRecord(s.ev_x0, s.xs));
  if (g->rank_mode) {
    NcclApi *api = nccl_api();
    NCCL_TRY(api, api->AllGather(s.send[0].p, s.gather.p, s.L.bytes, ncclChar, g->comm, s.xs));
  }
  CUDA_TRY(cudaEventRecord(s.ev_x1, s.xs));
  CUDA_TRY(launch_merge_packed(g->n_shards, s.nq, s.k, s.gather.p, s.L.bytes, s.L, s.outb.p, s.xs));
  CUDA_TRY(cudaEventRecord(s.ev_m, s.xs));
  return KDBGPU_OK;
}

void fill_stats(kdbgpu_shard_group *g, Slot &s, kdbgpu_shard_stats *stats) {
  if (!stats) return;
  memset(stats, 0, sizeof *stats);
  unsigned long long st[4];
  memcpy(st, s.h_out + s.L.o_stats, sizeof st);
  stats->dist_evals = st[0];
  stats->hops = st[1];
  stats->hops_l0 = st[2];
  stats->n_shards = (uint32_t)g->n_shards;
  float trav = 0.f, xch = 0.f;
  for (size_t m = 0; m < g->members.size(); ++m) {
    DeviceGuard dm(g->members[m].device);
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, s.ev_m0[m], s.ev_m1[m]) == cudaSuccess && a > trav) trav = a;
    if (cudaEventElapsedTime(&b, s.ev_m1[m], s.ev_m2[m]) == cudaSuccess && b > xch) xch = b;
  }
  (void)cudaGetLastError();
  DeviceGuard dg(g->merge_device);
  stats->traversal_ms = trav;
  float x = 0.f;
  cudaEventElapsedTime(&x, s.ev_x0, s.ev_x1);
  stats->exchange_ms = g->rank_mode ? x : xch;
  cudaEventElapsedTime(&stats->merge_ms, s.ev_x1, s.ev_m);
  cudaEventElapsedTime(&stats->total_ms, s.ev_start, s.ev_done);
  (void)cudaGetLastError();
}

int check_shape(kdbgpu_shard_group *g, uint32_t nq, int k) {
  if (k <= 0 || k > 10000) return set_error(KDBGPU_ERR_INVALID, "k %d outside 1..10000", k);
  if ((size_t)g->n_shards * (size_t)k * 12 > 200 * 1024)
    return set_error(KDBGPU_ERR_INVALID, "k %d x %d shards exceeds the merge kernel's shared memory", k, g->n_shards);
  (void)nq;
  return KDBGPU_OK;
}

int submit_impl(kdbgpu_shard_group *g, const float *queries, int q_device, cudaEvent_t q_ready, uint32_t nq, int k,
                int ef_search, const uint64_t *allow, size_t allow_words, int *slot_out,
                std::unique_lock<std::mutex> &lk) {
  const int ef = ef_search < k ? k : ef_search;  // hnsw_index.go:2377-2380
  const int si = slot_acquire(g, lk);
  Slot &s = g->slots[si];
  int rc = slot_fold_error(g, s);
  if (rc == KDBGPU_OK) rc = slot_reserve(g, s, nq, k);
  if (rc == KDBGPU_OK) {
    DeviceGuard dg(g->merge_device);
    cudaError_t e = cudaEventRecord(s.ev_start, s.xs);
    if (e != cudaSuccess) rc = set_error(KDBGPU_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
  }
  for (size_t m = 0; rc == KDBGPU_OK && m < g->members.size(); ++m)
    rc = enqueue_member(g, s, m, queries, q_device, q_ready, nq, k, ef, allow, allow_words);
  if (rc == KDBGPU_OK) rc = enqueue_exchange_merge(g, s);
  if (rc) {
    s.busy = false;
    g->cv.notify_all();
    return rc;
  }
  *slot_out = si;
  return KDBGPU_OK;
}

}  // namespace

extern "C" {

int kdbgpu_shard_unique_id(unsigned char id[KDBGPU_SHARD_ID_BYTES]) {
  if (!id) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  static_assert(sizeof(ncclUniqueId) == KDBGPU_SHARD_ID_BYTES, "ncclUniqueId is 128 bytes");
  NcclApi *api = nccl_api();
  if (!api->lib) return set_error(KDBGPU_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded: %s", api->err.c_str());
  ncclUniqueId u;
  NCCL_TRY(api, api->GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return KDBGPU_OK;
}

int kdbgpu_shard_group_create_rank(kdbgpu_index *local, int rank, int world, const unsigned char id[KDBGPU_SHARD_ID_BYTES],
                                   uint32_t id_base, kdbgpu_shard_group **out) {
  if (!local || !id || !out) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  if (world < 1 || world > kMaxShards || rank < 0 || rank >= world)
    return set_error(KDBGPU_ERR_INVALID, "rank %d / world %d (1..%d shards)", rank, world, kMaxShards);
  NcclApi *api = nccl_api();
  if (!api->lib)
    return set_error(KDBGPU_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded (%s); set KDBGPU_NCCL_LIB", api->err.c_str());
  kdbgpu_shard_group *g = new (std::nothrow) kdbgpu_shard_group();
  if (!g) return set_error(KDBGPU_ERR_NOMEM, "out of host memory");
  g->rank_mode = true;
  g->n_shards = world;
  g->rank = rank;
  Member mb;
  mb.h = local;
  mb.id_base = id_base;
  mb.device = local->device;
  g->members.push_back(mb);
  g->merge_device = local->device;
  DeviceGuard dg(local->device);
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclResult_t r = api->CommInitRank(&g->comm, world, u, rank);
  if (r != ncclSuccess) {
    delete g;
    return set_error(KDBGPU_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString(r));
  }
  for (auto &s : g->slots) {
    int rc = slot_init(g, s);
    if (rc) {
      kdbgpu_shard_group_destroy(g);
      return rc;
    }
  }
  *out = g;
  return KDBGPU_OK;
}

int kdbgpu_shard_group_create_local(kdbgpu_index *const *shards, int n_shards, const uint32_t *id_bases,
                                    kdbgpu_shard_group **out) {
  if (!shards || !id_bases || !out) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  if (n_shards < 1 || n_shards > kMaxShards) return set_error(KDBGPU_ERR_INVALID, "n_shards %d outside 1..%d", n_shards, kMaxShards);
  for (int i = 0; i < n_shards; ++i) {
    if (!shards[i]) return set_error(KDBGPU_ERR_INVALID, "shard %d is NULL", i);
    if (shards[i]->dim != shards[0]->dim || shards[i]->metric != shards[0]->metric ||
        shards[i]->precision != shards[0]->precision)
      return set_error(KDBGPU_ERR_INVALID, "shard %d differs from shard 0 in dim / metric / precision", i);
  }
  kdbgpu_shard_group *g = new (std::nothrow) kdbgpu_shard_group();
  if (!g) return set_error(KDBGPU_ERR_NOMEM, "out of host memory");
  g->rank_mode = false;
  g->n_shards = n_shards;
  g->merge_device = shards[0]->device;
  for (int i = 0; i < n_shards; ++i) {
    Member mb;
    mb.h = shards[i];
    mb.id_base = id_bases[i];
    mb.device = shards[i]->device;
    g->members.push_back(mb);
    if (mb.device != g->merge_device) {  // direct NVLink path for the peer copies (staged through the host otherwise)
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, mb.device, g->merge_device) == cudaSuccess && can) {
        DeviceGuard dm(mb.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(g->merge_device, 0);
        if (e != cudaSuccess) (void)cudaGetLastError();  // already enabled is fine
      }
      (void)cudaGetLastError();
    }
  }
  for (auto &s : g->slots) {
    int rc = slot_init(g, s);
    if (rc) {
      kdbgpu_shard_group_destroy(g);
      return rc;
    }
  }
  *out = g;
  return KDBGPU_OK;
}

int kdbgpu_shard_group_destroy(kdbgpu_shard_group *g) {
  if (!g) return KDBGPU_OK;
  {
    std::unique_lock<std::mutex> lk(g->mu);
    g->cv.wait(lk, [&] {
      for (auto &s : g->slots)
        if (s.busy) return false;
      return true;
    });
  }
  for (auto &mb : g->members) {
    DeviceGuard dm(mb.device);
    cudaDeviceSynchronize();
  }
  for (auto &s : g->slots) slot_destroy(g, s);
  if (g->user_default) {
    DeviceGuard dg(g->merge_device);
    cudaStreamDestroy(g->user_default);
  }
  if (g->comm) {
    NcclApi *api = nccl_api();
    if (api->lib) api->CommDestroy(g->comm);
  }
  (void)cudaGetLastError();
  delete g;
  return KDBGPU_OK;
}

int kdbgpu_shard_group_size(const kdbgpu_shard_group *g) { return g ? g->n_shards : 0; }

int kdbgpu_shard_search_submit(kdbgpu_shard_group *g, const float *queries, uint32_t nq, int k, int ef_search,
                               const uint64_t *allow, size_t allow_words, kdbgpu_shard_ticket **ticket) {
  if (!g || !ticket) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  *ticket = nullptr;
  if (nq == 0 || !queries) return set_error(KDBGPU_ERR_INVALID, "empty batch");
  int rc = check_shape(g, nq, k);
  if (rc) return rc;
  kdbgpu_shard_ticket *t = new (std::nothrow) kdbgpu_shard_ticket();
  if (!t) return set_error(KDBGPU_ERR_NOMEM, "out of host memory");
  std::unique_lock<std::mutex> lk(g->mu);
  int si = -1;
  rc = submit_impl(g, queries, -1, nullptr, nq, k, ef_search, allow, allow_words, &si, lk);
  if (rc) {
    delete t;
    return rc;
  }
  Slot &s = g->slots[si];
  {
    DeviceGuard dg(g->merge_device);
    cudaError_t e = cudaMemcpyAsync(s.h_out, s.outb.p, s.L.bytes, cudaMemcpyDeviceToHost, s.xs);
    if (e == cudaSuccess) e = cudaEventRecord(s.ev_done, s.xs);
    s.used = true;
    s.device_form = false;
    if (e != cudaSuccess) {
      // the slot's stream state is unknown: drain it before handing the slot back
      cudaStreamSynchronize(s.xs);
      (void)cudaGetLastError();
      s.busy = false;
      g->cv.notify_all();
      delete t;
      return set_error(KDBGPU_ERR_CUDA, "queueing the result copy: %s", cudaGetErrorString(e));
    }
  }
  t->g = g;
  t->slot = si;
  *ticket = t;
  return KDBGPU_OK;
}

int kdbgpu_shard_search_wait(kdbgpu_shard_ticket *t, uint32_t *out_ids, double *out_scores, uint32_t *out_counts,
                             kdbgpu_shard_stats *stats) {
  if (!t) return set_error(KDBGPU_ERR_INVALID, "NULL ticket");
  kdbgpu_shard_group *g = t->g;
  Slot &s = g->slots[t->slot];
  int rc = KDBGPU_OK;
  {
    DeviceGuard dg(g->merge_device);
    cudaError_t e = cudaEventSynchronize(s.ev_done);
    if (e != cudaSuccess) rc = set_error(KDBGPU_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
  }
  if (rc == KDBGPU_OK) {
    const size_t nk = (size_t)s.nq * s.k;
    if (out_scores) memcpy(out_scores, s.h_out + s.L.o_scores, nk * sizeof(double));
    if (out_ids) memcpy(out_ids, s.h_out + s.L.o_ids, nk * sizeof(uint32_t));
    if (out_counts) memcpy(out_counts, s.h_out + s.L.o_counts, (size_t)s.nq * sizeof(uint32_t));
    fill_stats(g, s, stats);
    int err = 0;
    memcpy(&err, s.h_out + s.L.o_err, sizeof err);
    if (err == KDBGPU_ERR_OVERFLOW)
      rc = set_error(KDBGPU_ERR_OVERFLOW, "candidate heap bound exceeded for at least one query on at least one shard");
    else if (err)
      rc = set_error(err, "a shard reported error %d", err);
  }
  slot_release(g, t->slot);
  delete t;
  return rc;
}

int kdbgpu_shard_search_batch(kdbgpu_shard_group *g, const float *queries, uint32_t nq, int k, int ef_search,
                              const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                              uint32_t *out_counts, kdbgpu_shard_stats *stats) {
  if (!g) return set_error(KDBGPU_ERR_INVALID, "NULL group");
  if (stats) memset(stats, 0, sizeof *stats);
  if (nq == 0) return KDBGPU_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  kdbgpu_shard_ticket *t = nullptr;
  int rc = kdbgpu_shard_search_submit(g, queries, nq, k, ef_search, allow, allow_words, &t);
  if (rc) return rc;
  return kdbgpu_shard_search_wait(t, out_ids, out_scores, out_counts, stats);
}

int kdbgpu_shard_search_batch_device(kdbgpu_shard_group *g, const float *d_queries, uint32_t nq, int k, int ef_search,
                                     uint32_t *d_out_ids, double *d_out_scores, uint32_t *d_out_counts, void *stream) {
  if (!g) return set_error(KDBGPU_ERR_INVALID, "NULL group");
  if (nq == 0) return KDBGPU_OK;
  if (!d_queries || !d_out_ids || !d_out_scores || !d_out_counts) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  int rc = check_shape(g, nq, k);
  if (rc) return rc;
  std::unique_lock<std::mutex> lk(g->mu);
  DeviceGuard dg(g->merge_device);
  cudaStream_t us = reinterpret_cast<cudaStream_t>(stream);
  if (!us) {
    if (!g->user_default) CUDA_TRY(cudaStreamCreateWithFlags(&g->user_default, cudaStreamNonBlocking));
    us = g->user_default;
  }
  // the queries are ready when `us` reaches this point
  cudaEvent_t q_ready = nullptr;
  CUDA_TRY(cudaEventCreateWithFlags(&q_ready, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(q_ready, us);
  int si = -1;
  if (e == cudaSuccess)
    rc = submit_impl(g, d_queries, g->members[0].device, q_ready, nq, k, ef_search, nullptr, 0, &si, lk);
  cudaEventDestroy(q_ready);  // released once the waits queued on it have been satisfied
  if (e != cudaSuccess) return set_error(KDBGPU_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
  if (rc) return rc;
  Slot &s = g->slots[si];
  DeviceGuard dg2(g->merge_device);
  const size_t nk = (size_t)nq * k;
  e = cudaMemcpyAsync(d_out_scores, s.outb.p + s.L.o_scores, nk * sizeof(double), cudaMemcpyDeviceToDevice, s.xs);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_out_ids, s.outb.p + s.L.o_ids, nk * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s.xs);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_out_counts, s.outb.p + s.L.o_counts, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s.xs);
  if (e == cudaSuccess)  // counters + error flag only: the results stay on the device
    e = cudaMemcpyAsync(s.h_out + s.L.o_stats, s.outb.p + s.L.o_stats, s.L.bytes - s.L.o_stats, cudaMemcpyDeviceToHost, s.xs);
  if (e == cudaSuccess) e = cudaEventRecord(s.ev_done, s.xs);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(us, s.ev_done, 0);
  s.used = true;
  s.device_form = true;
  g->last_device_slot = si;
  s.busy = false;  // asynchronous form: the next user orders itself behind ev_done on the device
  g->cv.notify_all();
  if (e != cudaSuccess) return set_error(KDBGPU_ERR_CUDA, "queueing the result copies: %s", cudaGetErrorString(e));
  return KDBGPU_OK;
}

int kdbgpu_shard_sync(kdbgpu_shard_group *g, kdbgpu_shard_stats *stats) {
  if (!g) return set_error(KDBGPU_ERR_INVALID, "NULL group");
  if (stats) memset(stats, 0, sizeof *stats);
  std::unique_lock<std::mutex> lk(g->mu);
  for (auto &s : g->slots) {
    int rc = slot_fold_error(g, s);
    if (rc) return rc;
  }
  if (g->last_device_slot >= 0) fill_stats(g, g->slots[g->last_device_slot], stats);
  const int err = g->sticky_err;
  g->sticky_err = 0;
  if (err == KDBGPU_ERR_OVERFLOW)
    return set_error(KDBGPU_ERR_OVERFLOW, "candidate heap bound exceeded for at least one query on at least one shard");
  if (err) return set_error(err, "a shard reported error %d", err);
  return KDBGPU_OK;
}

int kdbgpu_shard_flat_search_batch(kdbgpu_shard_group *g, const float *queries, uint32_t nq, int k, int mode,
                                   const uint64_t *allow, size_t allow_words, uint32_t *out_ids, double *out_scores,
                                   uint32_t *out_counts, kdbgpu_shard_stats *stats) {
  if (!g) return set_error(KDBGPU_ERR_INVALID, "NULL group");
  if (stats) memset(stats, 0, sizeof *stats);
  if (nq == 0) return KDBGPU_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  if (k <= 0 || k > 1024) return set_error(KDBGPU_ERR_INVALID, "flat k %d outside 1..1024", k);
  const bool prefilter = (mode & KDBGPU_FLAT_PREFILTER) != 0;
  const int fmode = mode & ~KDBGPU_FLAT_PREFILTER;
  if (fmode != 0 && fmode != 1) return set_error(KDBGPU_ERR_INVALID, "mode %d", fmode);
  int rc = check_shape(g, nq, k);
  if (rc) return rc;
  // BruteForceIndex treats an empty allow-list as unfiltered (vector_index.go:132) — decided on the WHOLE list
  bool filtered = false;
  if (allow) (void)first_set_bit(allow, allow_words, &filtered);
  std::unique_lock<std::mutex> lk(g->mu);
  const int si = slot_acquire(g, lk);
  Slot &s = g->slots[si];
  rc = slot_fold_error(g, s);
  if (rc == KDBGPU_OK) rc = slot_reserve(g, s, nq, k);
  if (rc == KDBGPU_OK) {
    DeviceGuard dg(g->merge_device);
    if (s.used) cudaEventSynchronize(s.ev_done);
    cudaError_t e = cudaEventRecord(s.ev_start, s.xs);
    if (e != cudaSuccess) rc = set_error(KDBGPU_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
  }
  // every local shard scans its rows (its own thread: the scan synchronises its stream chunk by chunk), results
  // packed into s.send[m] on the device, global ids
  auto scan_one = [&](size_t m) -> int {
    Member &mb = g->members[m];
    kdbgpu_index *h = mb.h;
    DeviceGuard dm(mb.device);
    if (h->precision != KDBGPU_PRECISION_F32) return set_error(KDBGPU_ERR_INVALID, "the flat scan exists for float32 indexes only");
    std::shared_lock<std::shared_mutex> hl(h->mu);
    if (!h->has_graph) return set_error(KDBGPU_ERR_STATE, "shard %d: no rows staged", shard_index(g, m));
    const int wi = acquire_fws(h);
    kdbgpu_index::FlatWs &w = h->fws[wi];
    struct Release {
      kdbgpu_index *h;
      int wi;
      ~Release() { release_fws(h, wi); }
    } releaser{h, wi};
    CUDA_TRY(s.send[m].reserve(s.L.bytes));
    unsigned char *blob = s.send[m].p;
    cudaStream_t st = w.stream;
    CUDA_TRY(cudaEventRecord(s.ev_m0[m], st));
    CUDA_TRY(cudaMemsetAsync(blob, 0, s.L.bytes, st));
    kdbgpu_stats fs;
    memset(&fs, 0, sizeof fs);
    if (h->n > 0) {
      const uint32_t *d_allow = nullptr;
      if (filtered) {
        const size_t g32 = allow_words * 2;
        const size_t need32 = ((size_t)h->capacity + 1 + 31) / 32 + 2;
        CUDA_TRY(s.allow_g[m].reserve(g32));
        CUDA_TRY(w.allow.reserve(need32));
        CUDA_TRY(cudaMemcpyAsync(s.allow_g[m].p, allow, g32 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(launch_slice_bits(s.allow_g[m].p, g32, mb.id_base, w.allow.p, need32, st));
        d_allow = w.allow.p;
      }
      int frc = flat_search_ws(h, w, queries, nq, k, fmode, prefilter, d_allow, reinterpret_cast<uint32_t *>(blob + s.L.o_ids),
                               reinterpret_cast<double *>(blob + s.L.o_scores),
                               reinterpret_cast<uint32_t *>(blob + s.L.o_counts), &fs);
      if (frc) return frc;
      CUDA_TRY(launch_add_id_base(reinterpret_cast<uint32_t *>(blob + s.L.o_ids), (size_t)nq * k, mb.id_base, st));
      const unsigned long long evals = fs.dist_evals;
      CUDA_TRY(cudaMemcpyAsync(blob + s.L.o_stats, &evals, sizeof evals, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaEventRecord(s.ev_m1[m], st));
    if (!g->rank_mode) {
      unsigned char *dst = s.gather.p + (size_t)m * s.L.bytes;
      if (mb.device == g->merge_device)
        CUDA_TRY(cudaMemcpyAsync(dst, blob, s.L.bytes, cudaMemcpyDeviceToDevice, st));
      else
        CUDA_TRY(cudaMemcpyPeerAsync(dst, g->merge_device, blob, mb.device, s.L.bytes, st));
    }
    CUDA_TRY(cudaEventRecord(s.ev_m2[m], st));
    // the workspace goes back to the pool when this returns: its stream must be drained first
    CUDA_TRY(cudaStreamSynchronize(st));
    return KDBGPU_OK;
  };
  if (rc == KDBGPU_OK) {
    if (g->members.size() == 1) {
      rc = scan_one(0);
    } else {
      std::vector<std::future<int>> fut;
      std::vector<std::string> errs;  // outlives every scan thread (they are joined in the handler below)
      try {
        errs.resize(g->members.size());
        fut.reserve(g->members.size());
        for (size_t m = 0; m < g->members.size(); ++m)
          fut.push_back(std::async(std::launch::async, [&, m]() -> int {
            const int r = scan_one(m);
            if (r) errs[m] = kdbgpu_last_error();  // the message lives in the worker thread
            return r;
          }));
        for (size_t m = 0; m < fut.size(); ++m) {
          const int r = fut[m].get();
          if (r && rc == KDBGPU_OK) rc = set_error(r, "%s", errs[m].c_str());
        }
      } catch (...) {  // no thread / no memory: the scans already started are joined (the futures' destructors below),
                       // the slot is handed back by the common exit, nothing unwinds across the C boundary
        for (auto &f : fut)
          if (f.valid()) f.wait();
        rc = set_error(KDBGPU_ERR_NOMEM, "could not start one scan thread per local shard");
      }
    }
  }
  if (rc == KDBGPU_OK) rc = enqueue_exchange_merge(g, s);
  if (rc == KDBGPU_OK) {
    DeviceGuard dg(g->merge_device);
    cudaError_t e = cudaMemcpyAsync(s.h_out, s.outb.p, s.L.bytes, cudaMemcpyDeviceToHost, s.xs);
    if (e == cudaSuccess) e = cudaEventRecord(s.ev_done, s.xs);
    if (e == cudaSuccess) e = cudaEventSynchronize(s.ev_done);
    s.used = true;
    s.device_form = false;
    if (e != cudaSuccess) rc = set_error(KDBGPU_ERR_CUDA, "result copy: %s", cudaGetErrorString(e));
  }
  if (rc == KDBGPU_OK) {
    const size_t nk = (size_t)nq * k;
    memcpy(out_scores, s.h_out + s.L.o_scores, nk * sizeof(double));
    memcpy(out_ids, s.h_out + s.L.o_ids, nk * sizeof(uint32_t));
    memcpy(out_counts, s.h_out + s.L.o_counts, (size_t)nq * sizeof(uint32_t));
    fill_stats(g, s, stats);
  }
  s.busy = false;
  g->cv.notify_all();
  return rc;
}

}  // extern "C"
