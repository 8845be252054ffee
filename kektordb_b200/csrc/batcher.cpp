// batcher.cpp — the micro-batcher in front of kdbgpu_search_batch (host side).
//
// The reference answers ONE query per call: every HTTP / MCP / RAG request runs on its own goroutine,
// calls idx.SearchWithScores(query, k, allowList, efSearch) (pkg/engine/ops.go:1006, :1296) and blocks
// until its own result is ready.  The device wants hundreds of queries per launch.  This layer keeps
// the reference's call shape and forms the batches underneath:
//
//   * queries with the same (k, ef, allow-list) are grouped; a group is dispatched at once while the
//     device is idle (a lone caller pays no batching latency), and otherwise when it reaches max_batch
//     queries or max_wait_us has passed since its first query — so batch size follows load;
//   * the batches are run by the batcher's own worker threads (as many as batches may be in flight), so
//     NO caller thread is needed to carry a batch.  That is what makes the asynchronous form possible:
//     kdbgpu_batcher_submit copies the query into the group's (pinned) staging buffer and returns a
//     ticket; kdbgpu_batcher_poll hands out the tickets of finished queries; kdbgpu_batcher_take copies one
//     result out.  A Go host serves thousands of in-flight searches from a handful of OS threads this way
//     (goroutines block on channels, not inside cgo) — SURVEY.md §7 "no batch search API upstream";
//   * kdbgpu_batcher_search (one blocking call per query) is submit + wait on the same machinery.
//
// Errors follow the reference: a failed search yields an empty result for that caller
// (hnsw_index.go:355-359) and the error code is returned.  Nothing throws across the C boundary.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <list>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/kektordb_gpu.h"

namespace {

using Clock = std::chrono::steady_clock;

// staging of one group: queries in, results out; page-locked when a CUDA device is present (kdbgpu_host_alloc)
struct GroupBuf {
  float *q = nullptr;
  double *sc = nullptr;
  uint32_t *ids = nullptr, *cnt = nullptr;
  size_t k_cap = 0;
  void release() {
    kdbgpu_host_free(q);
    kdbgpu_host_free(sc);
    kdbgpu_host_free(ids);
    kdbgpu_host_free(cnt);
    q = nullptr;
    sc = nullptr;
    ids = cnt = nullptr;
    k_cap = 0;
  }
};

struct Filter {  // an allow-list as the batcher keeps it: immutable copy + hash
  std::vector<uint64_t> words;
  uint64_t hash = 0;
};

uint64_t hash_words(const uint64_t *w, size_t n) {
  uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)n;
  for (size_t i = 0; i < n; ++i) {
    h ^= w[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
  }
  return h;
}

struct Group {
  uint64_t seq = 0;
  int k = 0, ef = 0;
  std::shared_ptr<const Filter> filter;  // nullptr = nil allow-list
  GroupBuf buf;
  uint32_t n = 0;                    // slots handed out (guarded by the batcher mutex)
  std::atomic<uint32_t> filled{0};   // queries copied into buf.q
  // per slot: 0 = blocking caller, 1 = ticket outstanding (announced through the poll queue), 2 = ticket taken
  std::unique_ptr<std::atomic<uint8_t>[]> notify;
  Clock::time_point deadline;
  bool open = true;  // still accepting members (guarded by the batcher mutex)
  // completion is signalled under the group's own mutex, so that a thousand blocked callers waking up contend
  // with each other only — not with the callers that are forming the next groups
  std::mutex done_mu;
  std::condition_variable done_cv;
  bool done = false;
  int rc = KDBGPU_OK;
  std::atomic<uint32_t> taken{0};  // results copied out
};

}  // namespace

struct kdbgpu_batcher {
  kdbgpu_batch_fn fn = nullptr;
  void *ctx = nullptr;
  int dim = 0;
  uint32_t max_batch = 1024;
  uint32_t max_wait_us = 200;
  std::mutex mu;
  std::condition_variable work_cv;  // workers: a group may have become ready
  std::list<std::shared_ptr<Group>> open_groups;
  std::unordered_map<uint64_t, std::shared_ptr<Group>> live;  // seq -> group, while results are outstanding (async form)
  std::vector<GroupBuf> free_bufs;
  std::unordered_map<uint64_t, std::shared_ptr<const Filter>> filters;  // registered allow-lists
  uint64_t next_filter = 1;
  uint64_t next_seq = 1;
  uint32_t inflight = 0;  // batch calls currently executing
  bool closing = false;
  std::atomic<uint32_t> callers{0};  // threads inside a blocking entry point
  std::vector<std::thread> workers;
  // completion queue of the asynchronous form
  std::mutex cq_mu;
  std::condition_variable cq_cv;
  std::deque<uint64_t> cq;
  // counters
  uint64_t n_queries = 0, n_batches = 0, max_seen = 0, n_immediate = 0, n_full = 0, n_timeout = 0;
};

namespace {

int index_exec(void *ctx, const float *queries, uint32_t nq, int k, int ef_search, const uint64_t *allow,
               size_t allow_words, uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
  return kdbgpu_search_batch(static_cast<kdbgpu_index *>(ctx), queries, nq, k, ef_search, allow, allow_words, out_ids,
                             out_scores, out_counts, nullptr);
}

int group_exec(void *ctx, const float *queries, uint32_t nq, int k, int ef_search, const uint64_t *allow,
               size_t allow_words, uint32_t *out_ids, double *out_scores, uint32_t *out_counts) {
  return kdbgpu_shard_search_batch(static_cast<kdbgpu_shard_group *>(ctx), queries, nq, k, ef_search, allow, allow_words,
                                   out_ids, out_scores, out_counts, nullptr);
}

// staging for a new group (caller holds b->mu); returns false when memory is exhausted
bool take_buf(kdbgpu_batcher *b, int k, GroupBuf *out) {
  for (size_t i = 0; i < b->free_bufs.size(); ++i)
    if (b->free_bufs[i].k_cap >= (size_t)k) {
      *out = b->free_bufs[i];
      b->free_bufs[i] = b->free_bufs.back();
      b->free_bufs.pop_back();
      return true;
    }
  GroupBuf nb;
  const size_t mb = b->max_batch, kc = (size_t)k < 16 ? 16 : (size_t)k;
  auto alloc = [](size_t bytes) -> void * {
    void *p = nullptr;
    return kdbgpu_host_alloc(&p, bytes) == KDBGPU_OK ? p : nullptr;
  };
  nb.q = static_cast<float *>(alloc(mb * (size_t)b->dim * sizeof(float)));
  nb.sc = static_cast<double *>(alloc(mb * kc * sizeof(double)));
  nb.ids = static_cast<uint32_t *>(alloc(mb * kc * sizeof(uint32_t)));
  nb.cnt = static_cast<uint32_t *>(alloc(mb * sizeof(uint32_t)));
  const bool ok = nb.q && nb.sc && nb.ids && nb.cnt;
  if (!ok) {
    nb.release();
    return false;
  }
  nb.k_cap = kc;
  *out = nb;
  return true;
}

void give_buf(kdbgpu_batcher *b, GroupBuf &buf) {  // caller holds b->mu
  if (!buf.q) return;
  if (b->free_bufs.size() < 16)
    b->free_bufs.push_back(buf);
  else
    buf.release();
  buf = GroupBuf();
}

bool ready(const kdbgpu_batcher *b, const Group &g, Clock::time_point now, int *why) {
  if (g.n >= b->max_batch) {
    *why = 1;
    return true;
  }
  if (b->inflight == 0) {
    *why = 0;
    return true;
  }
  if (now >= g.deadline) {
    *why = 2;
    return true;
  }
  return false;
}

void worker_main(kdbgpu_batcher *b) {
  std::unique_lock<std::mutex> lk(b->mu);
  for (;;) {
    // ---- pick the oldest ready group, or sleep until one can become ready
    std::shared_ptr<Group> g;
    int why = 0;
    const auto now = Clock::now();
    auto wake = Clock::time_point::max();
    for (auto &og : b->open_groups) {
      if (ready(b, *og, now, &why)) {
        g = og;
        break;
      }
      if (og->deadline < wake) wake = og->deadline;
    }
    if (!g) {
      if (b->closing && b->open_groups.empty()) return;
      if (wake == Clock::time_point::max())
        b->work_cv.wait(lk);
      else
        b->work_cv.wait_until(lk, wake);
      continue;
    }
    g->open = false;
    b->open_groups.remove(g);
    b->inflight++;
    const uint32_t nq = g->n;
    b->n_queries += nq;
    b->n_batches++;
    if (nq > b->max_seen) b->max_seen = nq;
    (why == 0 ? b->n_immediate : why == 1 ? b->n_full : b->n_timeout)++;
    lk.unlock();
    // ---- run the batch.  Members reserve their slot under the mutex and copy their query outside it.
    while (g->filled.load(std::memory_order_acquire) < nq) std::this_thread::yield();
    const Filter *f = g->filter.get();
    int rc = b->fn(b->ctx, g->buf.q, nq, g->k, g->ef, f ? f->words.data() : nullptr, f ? f->words.size() : 0, g->buf.ids,
                   g->buf.sc, g->buf.cnt);
    if (rc != KDBGPU_OK)  // a failed search yields an empty result (hnsw_index.go:355-359)
      memset(g->buf.cnt, 0, (size_t)nq * sizeof(uint32_t));
    {
      std::lock_guard<std::mutex> dl(g->done_mu);
      g->rc = rc;
      g->done = true;
    }
    g->done_cv.notify_all();
    bool any = false;
    {
      std::lock_guard<std::mutex> ql(b->cq_mu);
      for (uint32_t i = 0; i < nq; ++i)
        if (g->notify[i].load(std::memory_order_relaxed)) {
          b->cq.push_back((g->seq << 16) | i);
          any = true;
        }
    }
    if (any) b->cq_cv.notify_all();
    lk.lock();
    b->inflight--;
    b->work_cv.notify_all();  // a batch finished: collecting groups may go now
  }
}

// Reserves a slot in a group for (k, ef, filter), copies the query in.  Returns the group and the slot.
int enqueue(kdbgpu_batcher *b, const float *query, int k, int ef, std::shared_ptr<const Filter> filter,
            const uint64_t *raw_allow, size_t raw_words, bool notify, std::shared_ptr<Group> *out_g, uint32_t *out_idx) {
  uint64_t raw_hash = 0;
  if (!filter && raw_allow) raw_hash = hash_words(raw_allow, raw_words);  // outside the mutex
  std::shared_ptr<Group> g;
  uint32_t idx = 0;
  for (int attempt = 0; !g; ++attempt) {
    std::shared_ptr<Group> cand;
    {
      std::unique_lock<std::mutex> lk(b->mu);
      if (b->closing) return KDBGPU_ERR_STATE;
      for (auto &og : b->open_groups) {
        if (!og->open || og->k != k || og->ef != ef || og->n >= b->max_batch) continue;
        const Filter *gf = og->filter.get();
        if (filter) {
          if (gf != filter.get()) continue;
        } else if (raw_allow) {
          if (!gf || gf->hash != raw_hash || gf->words.size() != raw_words) continue;
          cand = og;  // same hash: the words are compared outside the mutex, below
          break;
        } else if (gf) {
          continue;
        }
        idx = og->n++;
        og->notify[idx].store(notify ? 1 : 0, std::memory_order_relaxed);
        g = og;
        if (og->n >= b->max_batch) b->work_cv.notify_all();
        break;
      }
      if (!g && !cand) {  // open a new group
        auto ng = std::make_shared<Group>();
        if (!take_buf(b, k, &ng->buf)) return KDBGPU_ERR_NOMEM;
        ng->seq = b->next_seq++;
        ng->k = k;
        ng->ef = ef;
        if (filter) {
          ng->filter = filter;
        } else if (raw_allow) {
          auto nf = std::make_shared<Filter>();
          nf->words.assign(raw_allow, raw_allow + raw_words);
          nf->hash = raw_hash;
          ng->filter = nf;
        }
        ng->notify.reset(new std::atomic<uint8_t>[b->max_batch]());
        ng->deadline = Clock::now() + std::chrono::microseconds(b->max_wait_us);
        idx = ng->n++;
        ng->notify[idx].store(notify ? 1 : 0, std::memory_order_relaxed);
        b->open_groups.push_back(ng);
        if (notify) b->live[ng->seq] = ng;
        g = ng;
        b->work_cv.notify_all();
      } else if (g && notify) {
        b->live[g->seq] = g;
      }
    }
    if (!g && cand) {
      // equal hashes: compare the words without holding the batcher mutex (MBs at 10 M ids), then join
      const bool same = memcmp(cand->filter->words.data(), raw_allow, raw_words * sizeof(uint64_t)) == 0;
      if (same) {
        std::unique_lock<std::mutex> lk(b->mu);
        if (cand->open && cand->n < b->max_batch) {
          idx = cand->n++;
          cand->notify[idx].store(notify ? 1 : 0, std::memory_order_relaxed);
          if (notify) b->live[cand->seq] = cand;
          g = cand;
          if (cand->n >= b->max_batch) b->work_cv.notify_all();
        }
      } else {
        raw_hash ^= 0x5bd1e9955bd1e995ull * (uint64_t)(attempt + 1);  // a true collision: never joins that group
      }
    }
  }
  memcpy(g->buf.q + (size_t)idx * b->dim, query, (size_t)b->dim * sizeof(float));
  g->filled.fetch_add(1, std::memory_order_release);
  *out_g = g;
  *out_idx = idx;
  return KDBGPU_OK;
}

int copy_out(kdbgpu_batcher *b, const std::shared_ptr<Group> &g, uint32_t idx, uint32_t *out_ids, double *out_scores,
             uint32_t *out_count) {
  {
    std::unique_lock<std::mutex> dl(g->done_mu);
    g->done_cv.wait(dl, [&] { return g->done; });
  }
  const int rc = g->rc;
  const size_t k = (size_t)g->k;
  if (rc == KDBGPU_OK) {
    memcpy(out_ids, g->buf.ids + idx * k, k * sizeof(uint32_t));
    memcpy(out_scores, g->buf.sc + idx * k, k * sizeof(double));
    *out_count = g->buf.cnt[idx];
  } else {
    *out_count = 0;
  }
  if (g->taken.fetch_add(1) + 1 == g->n) {  // last result out (the group is closed: n is final): recycle the staging
    std::lock_guard<std::mutex> lk(b->mu);
    give_buf(b, g->buf);
    b->live.erase(g->seq);
  }
  return rc;
}

struct Inside {  // counts the caller for kdbgpu_batcher_destroy
  kdbgpu_batcher *b;
  explicit Inside(kdbgpu_batcher *bb) : b(bb) { b->callers.fetch_add(1); }
  ~Inside() { b->callers.fetch_sub(1); }
};

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
  int nw = 4;  // = the batches one handle overlaps
  const char *env = getenv("KDBGPU_BATCHER_WORKERS");
  if (env && atoi(env) >= 1 && atoi(env) <= 64) nw = atoi(env);
  try {
    for (int i = 0; i < nw; ++i) b->workers.emplace_back(worker_main, b);
  } catch (...) {
    {
      std::lock_guard<std::mutex> lk(b->mu);
      b->closing = true;
    }
    b->work_cv.notify_all();
    for (auto &t : b->workers) t.join();
    delete b;
    return KDBGPU_ERR_NOMEM;
  }
  *out = b;
  return KDBGPU_OK;
}

int kdbgpu_batcher_create(kdbgpu_index *index, uint32_t max_batch, uint32_t max_wait_us, kdbgpu_batcher **out) {
  if (!index) return KDBGPU_ERR_INVALID;
  return kdbgpu_batcher_create_fn(index_exec, index, kdbgpu_index_dim(index), max_batch, max_wait_us, out);
}

int kdbgpu_batcher_create_group(kdbgpu_shard_group *group, int dim, uint32_t max_batch, uint32_t max_wait_us,
                                kdbgpu_batcher **out) {
  if (!group) return KDBGPU_ERR_INVALID;
  return kdbgpu_batcher_create_fn(group_exec, group, dim, max_batch, max_wait_us, out);
}

// Must not race new calls on the same batcher: the caller stops issuing searches first (queries already inside
// are answered).  Outstanding tickets become invalid.
int kdbgpu_batcher_destroy(kdbgpu_batcher *b) {
  if (!b) return KDBGPU_OK;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->closing = true;  // new queries are refused; the groups already formed are still run
  }
  b->work_cv.notify_all();
  for (auto &t : b->workers) t.join();
  b->cq_cv.notify_all();
  while (b->callers.load() != 0) std::this_thread::yield();
  {
    std::lock_guard<std::mutex> lk(b->mu);
    for (auto &kv : b->live) kv.second->buf.release();
    b->live.clear();
    for (auto &fb : b->free_bufs) fb.release();
    b->free_bufs.clear();
  }
  delete b;
  return KDBGPU_OK;
}

int kdbgpu_batcher_search(kdbgpu_batcher *b, const float *query, int k, int ef_search, const uint64_t *allow,
                          size_t allow_words, uint32_t *out_ids, double *out_scores, uint32_t *out_count) {
  if (!b || !query || !out_ids || !out_scores || !out_count || k <= 0) return KDBGPU_ERR_INVALID;
  *out_count = 0;
  Inside inside(b);
  try {
    std::shared_ptr<Group> g;
    uint32_t idx = 0;
    int rc = enqueue(b, query, k, ef_search, nullptr, allow, allow ? allow_words : 0, false, &g, &idx);
    if (rc) return rc;
    return copy_out(b, g, idx, out_ids, out_scores, out_count);
  } catch (...) {  // std::bad_alloc in the containers: nothing unwinds across the C boundary
    *out_count = 0;
    return KDBGPU_ERR_NOMEM;
  }
}

int kdbgpu_batcher_register_filter(kdbgpu_batcher *b, const uint64_t *allow, size_t allow_words, uint64_t *filter_id) {
  if (!b || !allow || !filter_id) return KDBGPU_ERR_INVALID;
  try {
    auto f = std::make_shared<Filter>();
    f->words.assign(allow, allow + allow_words);
    f->hash = hash_words(allow, allow_words);
    std::lock_guard<std::mutex> lk(b->mu);
    *filter_id = b->next_filter++;
    b->filters[*filter_id] = f;
    return KDBGPU_OK;
  } catch (...) {
    return KDBGPU_ERR_NOMEM;
  }
}

int kdbgpu_batcher_release_filter(kdbgpu_batcher *b, uint64_t filter_id) {
  if (!b) return KDBGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(b->mu);
  return b->filters.erase(filter_id) ? KDBGPU_OK : KDBGPU_ERR_INVALID;
}

int kdbgpu_batcher_submit(kdbgpu_batcher *b, const float *query, int k, int ef_search, const uint64_t *allow,
                          size_t allow_words, uint64_t filter_id, uint64_t *ticket) {
  if (!b || !query || !ticket || k <= 0) return KDBGPU_ERR_INVALID;
  try {
    std::shared_ptr<const Filter> f;
    if (filter_id) {
      std::lock_guard<std::mutex> lk(b->mu);
      auto it = b->filters.find(filter_id);
      if (it == b->filters.end()) return KDBGPU_ERR_INVALID;
      f = it->second;
    }
    std::shared_ptr<Group> g;
    uint32_t idx = 0;
    int rc = enqueue(b, query, k, ef_search, f, f ? nullptr : allow, (f || !allow) ? 0 : allow_words, true, &g, &idx);
    if (rc) return rc;
    *ticket = (g->seq << 16) | idx;
    return KDBGPU_OK;
  } catch (...) {
    return KDBGPU_ERR_NOMEM;
  }
}

int kdbgpu_batcher_poll(kdbgpu_batcher *b, uint64_t *tickets, uint32_t max_tickets, uint32_t timeout_us, uint32_t *n_done) {
  if (!b || !tickets || !n_done || max_tickets == 0) return KDBGPU_ERR_INVALID;
  *n_done = 0;
  Inside inside(b);
  std::unique_lock<std::mutex> ql(b->cq_mu);
  if (b->cq.empty() && timeout_us)
    b->cq_cv.wait_for(ql, std::chrono::microseconds(timeout_us), [&] { return !b->cq.empty(); });
  uint32_t n = 0;
  while (n < max_tickets && !b->cq.empty()) {
    tickets[n++] = b->cq.front();
    b->cq.pop_front();
  }
  *n_done = n;
  return KDBGPU_OK;
}

int kdbgpu_batcher_take(kdbgpu_batcher *b, uint64_t ticket, uint32_t *out_ids, double *out_scores, uint32_t *out_count) {
  if (!b || !out_ids || !out_scores || !out_count) return KDBGPU_ERR_INVALID;
  *out_count = 0;
  Inside inside(b);
  std::shared_ptr<Group> g;
  const uint32_t idx = (uint32_t)(ticket & 0xffffu);
  {
    std::lock_guard<std::mutex> lk(b->mu);
    auto it = b->live.find(ticket >> 16);
    if (it == b->live.end()) return KDBGPU_ERR_INVALID;
    g = it->second;
    if (idx >= g->n) return KDBGPU_ERR_INVALID;  // a slot that was never handed out
  }
  // a ticket is good for ONE take: a second one would count the group's results out early and recycle its staging
  // while other callers are still copying
  uint8_t outstanding = 1;
  if (!g->notify[idx].compare_exchange_strong(outstanding, 2)) return KDBGPU_ERR_INVALID;
  return copy_out(b, g, idx, out_ids, out_scores, out_count);
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
