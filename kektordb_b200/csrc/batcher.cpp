// batcher.cpp — the micro-batcher in front of kdbgpu_search_batch (host side).
//
// The reference answers ONE query per call: every HTTP / MCP / RAG request runs on its own goroutine,
// calls idx.SearchWithScores(query, k, allowList, efSearch) (pkg/engine/ops.go:1006, :1296) and blocks
// until its own result is ready.  The device wants hundreds of queries per launch.  This layer keeps
// the reference's call shape — one blocking call per query, any number of caller threads — and forms
// the batches underneath:
//
//   * leader / follower, no dispatcher thread: the first caller of a group becomes its leader, later
//     callers with the same (k, ef, allow-list) join it and sleep; the leader gathers the queries, runs
//     ONE batch call, scatters the results and wakes everybody.
//   * adaptive: a leader dispatches at once while the device is idle (a lone caller pays no batching
//     latency); while other batches are in flight it keeps collecting until the group is full or
//     max_wait_us has passed — so batch size follows load.
//   * several groups may be in flight at a time (the library overlaps up to 4 batches per handle).
//
// Errors follow the reference: a failed search yields an empty result for that caller
// (hnsw_index.go:355-359) and the error code is returned.
#include <atomic>
#include <chrono>
#include <thread>
#include <condition_variable>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/kektordb_gpu.h"

namespace {

struct Request {
  const float *query;
  uint32_t *out_ids;
  double *out_scores;
  uint32_t *out_count;
  int rc = KDBGPU_OK;
};

struct Group {
  int k, ef;
  const uint64_t *allow;  // the leader's bitset (borrowed for the duration of its call)
  size_t allow_words;
  std::vector<Request *> reqs;
  std::chrono::steady_clock::time_point deadline;
  bool open = true;  // still accepting members (guarded by the batcher mutex)
  std::condition_variable cv;  // leader: "full / device idle / deadline" (waits with the batcher mutex)
  // completion is signalled under the group's own mutex, so that a thousand followers waking up contend
  // with each other only — not with the callers that are forming the next groups
  std::mutex done_mu;
  std::condition_variable done_cv;
  bool done = false;
};

}  // namespace

struct kdbgpu_batcher {
  kdbgpu_batch_fn fn = nullptr;
  void *ctx = nullptr;
  int dim = 0;
  uint32_t max_batch = 1024;
  uint32_t max_wait_us = 200;
  std::mutex mu;
  std::list<std::shared_ptr<Group>> open_groups;
  uint32_t inflight = 0;  // batch calls currently executing
  std::atomic<uint32_t> callers{0};  // threads inside kdbgpu_batcher_search
  std::atomic<bool> closing{false};
  // counters
  uint64_t n_queries = 0, n_batches = 0, max_seen = 0, n_immediate = 0, n_full = 0, n_timeout = 0;
};

namespace {

int index_exec(void *ctx, const float *queries, uint32_t nq, int k, int ef_search, const uint64_t *allow,
               size_t allow_words, uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
  return kdbgpu_search_batch(static_cast<kdbgpu_index *>(ctx), queries, nq, k, ef_search, allow, allow_words, out_ids,
                             out_scores, out_counts, nullptr);
}

bool same_filter(const Group &g, const uint64_t *allow, size_t words) {
  if ((g.allow == nullptr) != (allow == nullptr)) return false;
  if (allow == nullptr) return true;
  if (g.allow_words != words) return false;
  return g.allow == allow || memcmp(g.allow, allow, words * sizeof(uint64_t)) == 0;
}

}  // namespace

extern "C" {

int kdbgpu_batcher_create_fn(kdbgpu_batch_fn fn, void *ctx, int dim, uint32_t max_batch, uint32_t max_wait_us,
                             kdbgpu_batcher **out) {
  if (!out) return KDBGPU_ERR_INVALID;
  *out = nullptr;
  if (!fn || dim <= 0 || max_batch == 0 || max_batch > 65536) return KDBGPU_ERR_INVALID;
  kdbgpu_batcher *b = new (std::nothrow) kdbgpu_batcher();
  if (!b) return KDBGPU_ERR_NOMEM;
  b->fn = fn;
  b->ctx = ctx;
  b->dim = dim;
  b->max_batch = max_batch;
  b->max_wait_us = max_wait_us;
  *out = b;
  return KDBGPU_OK;
}

int kdbgpu_batcher_create(kdbgpu_index *index, uint32_t max_batch, uint32_t max_wait_us, kdbgpu_batcher **out) {
  if (!index) return KDBGPU_ERR_INVALID;
  return kdbgpu_batcher_create_fn(index_exec, index, kdbgpu_index_dim(index), max_batch, max_wait_us, out);
}

int kdbgpu_batcher_destroy(kdbgpu_batcher *b) {
  if (!b) return KDBGPU_OK;
  b->closing.store(true);  // new callers are refused; the ones inside finish
  while (b->callers.load() != 0) std::this_thread::yield();
  {
    std::unique_lock<std::mutex> lk(b->mu);  // the last leader has left its critical section
  }
  delete b;
  return KDBGPU_OK;
}

int kdbgpu_batcher_search(kdbgpu_batcher *b, const float *query, int k, int ef_search, const uint64_t *allow,
                          size_t allow_words, uint32_t *out_ids, double *out_scores, uint32_t *out_count) {
  if (!b || !query || !out_ids || !out_scores || !out_count || k <= 0) return KDBGPU_ERR_INVALID;
  *out_count = 0;
  Request req;
  req.query = query;
  req.out_ids = out_ids;
  req.out_scores = out_scores;
  req.out_count = out_count;

  struct Inside {  // counts the caller for kdbgpu_batcher_destroy
    kdbgpu_batcher *b;
    explicit Inside(kdbgpu_batcher *bb) : b(bb) { b->callers.fetch_add(1); }
    ~Inside() { b->callers.fetch_sub(1); }
  } inside(b);
  if (b->closing.load()) return KDBGPU_ERR_STATE;
  std::unique_lock<std::mutex> lk(b->mu);
  // ---- join an open group with the same (k, ef, filter), or open one and lead it
  std::shared_ptr<Group> g;
  for (auto &og : b->open_groups)
    if (og->open && og->k == k && og->ef == ef_search && og->reqs.size() < b->max_batch &&
        same_filter(*og, allow, allow_words)) {
      g = og;
      break;
    }
  const bool leader = !g;
  if (leader) {
    g = std::make_shared<Group>();
    g->k = k;
    g->ef = ef_search;
    g->allow = allow;
    g->allow_words = allow_words;
    g->deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(b->max_wait_us);
    b->open_groups.push_back(g);
  }
  g->reqs.push_back(&req);
  if (!leader) {
    if (g->reqs.size() >= b->max_batch) g->cv.notify_all();  // full: wake the leader
    lk.unlock();
    std::unique_lock<std::mutex> dl(g->done_mu);
    g->done_cv.wait(dl, [&] { return g->done; });
    return req.rc;
  }
  // ---- leader: collect while the device is busy, then run the batch
  int why = 0;  // 0 immediate, 1 full, 2 deadline
  for (;;) {
    if (g->reqs.size() >= b->max_batch) {
      why = 1;
      break;
    }
    if (b->inflight == 0) {
      why = 0;
      break;
    }
    if (g->cv.wait_until(lk, g->deadline) == std::cv_status::timeout) {
      why = g->reqs.size() >= b->max_batch ? 1 : 2;
      break;
    }
  }
  g->open = false;
  b->open_groups.remove(g);
  b->inflight++;
  const uint32_t nq = (uint32_t)g->reqs.size();
  b->n_queries += nq;
  b->n_batches++;
  if (nq > b->max_seen) b->max_seen = nq;
  (why == 0 ? b->n_immediate : why == 1 ? b->n_full : b->n_timeout)++;
  lk.unlock();

  int rc = KDBGPU_OK;
  if (nq == 1) {  // no gather / scatter for a lone query
    rc = b->fn(b->ctx, query, 1, k, ef_search, allow, allow_words, out_ids, out_scores, out_count);
    if (rc != KDBGPU_OK) *out_count = 0;
    req.rc = rc;
  } else {
    const size_t dim = (size_t)b->dim;
    std::vector<float> q((size_t)nq * dim);
    std::vector<uint32_t> ids((size_t)nq * k), cnt(nq);
    std::vector<double> sc((size_t)nq * k);
    for (uint32_t i = 0; i < nq; ++i) memcpy(&q[(size_t)i * dim], g->reqs[i]->query, dim * sizeof(float));
    rc = b->fn(b->ctx, q.data(), nq, k, ef_search, allow, allow_words, ids.data(), sc.data(), cnt.data());
    for (uint32_t i = 0; i < nq; ++i) {
      Request *r = g->reqs[i];
      r->rc = rc;
      if (rc == KDBGPU_OK) {
        memcpy(r->out_ids, &ids[(size_t)i * k], (size_t)k * sizeof(uint32_t));
        memcpy(r->out_scores, &sc[(size_t)i * k], (size_t)k * sizeof(double));
        *r->out_count = cnt[i];
      } else {
        *r->out_count = 0;  // a failed search yields an empty result (hnsw_index.go:355-359)
      }
    }
  }
  {
    std::lock_guard<std::mutex> dl(g->done_mu);
    g->done = true;
  }
  g->done_cv.notify_all();
  lk.lock();
  b->inflight--;
  // a batch finished: leaders that were collecting may go now
  for (auto &og : b->open_groups) og->cv.notify_all();
  return rc;
}

int kdbgpu_batcher_stats(kdbgpu_batcher *b, kdbgpu_batcher_stats_t *out) {
  if (!b || !out) return KDBGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(b->mu);
  out->queries = b->n_queries;
  out->batches = b->n_batches;
  out->max_batch_seen = b->max_seen;
  out->dispatched_idle = b->n_immediate;
  out->dispatched_full = b->n_full;
  out->dispatched_deadline = b->n_timeout;
  return KDBGPU_OK;
}

}  // extern "C"
