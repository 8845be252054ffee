// refresher.cpp — the staleness policy of the GPU mirror (third part).
//
// The CPU index keeps changing while searches run: Add / AddBatch rewrite forward and reverse links
// (reference pkg/core/hnsw/hnsw_index.go:472-809, :717-783), Delete flips Node.Deleted (:2303-2336), Vacuum
// rewires parents, nils node slots and may re-elect the entry point (optimizer.go:195-274), Refine rewrites
// neighbour lists (:288-468).  Every mirror-changing call of the C ABI takes the handle exclusively and
// drains the searches in flight — one such section per changed row would starve the query path.  So the
// changes are QUEUED here, per (node, level) with last-write-wins (an insert typically rewrites the same
// hub rows many times), and applied in ONE exclusive section — upload the new rows, register the new nodes,
// patch every pending adjacency row, nil the removed nodes, the deleted bitset, the entry point — when
//   * the queue holds max_pending_rows pending adjacency rows, or
//   * the oldest pending change is max_lag_ms old (a background thread watches the clock), or
//   * the host calls kdbgpu_refresher_flush (e.g. before a read that must see its own write).
// Between flushes searches see the mirror as of the last flush: a bounded-staleness snapshot, which is what
// the reference's own concurrent readers see of a graph under insertion (lock-free reads of Connections).
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/kektordb_gpu.h"

namespace {
using Clock = std::chrono::steady_clock;

struct NewNode {
  int level;
  std::vector<unsigned char> row;  // the vector in stored form (dim x 4 / 2 / 1 bytes)
};
}  // namespace

struct kdbgpu_refresher {
  kdbgpu_index *h = nullptr;
  uint32_t max_pending_rows = 4096;
  uint32_t max_lag_ms = 50;
  size_t row_bytes = 0;
  std::mutex mu;               // the queue
  std::mutex flush_mu;         // one flush at a time
  std::condition_variable cv;  // wakes the clock thread
  std::map<uint32_t, NewNode> new_nodes;                            // id -> level + vector
  std::map<std::pair<uint32_t, int>, std::vector<uint32_t>> rows;   // (id, level) -> neighbours
  std::vector<uint32_t> removed;
  std::vector<uint64_t> deleted;  // dense Node.Deleted bitset as last sent / as pending
  bool deleted_dirty = false;
  bool entry_dirty = false;
  uint32_t entry = 0;
  int max_level = -1;
  bool has_pending = false;
  Clock::time_point oldest;
  bool closing = false;
  std::thread clock_thread;
  int last_error = KDBGPU_OK;
  // counters
  uint64_t flushes = 0, rows_applied = 0, rows_queued = 0, nodes_applied = 0, by_rows = 0, by_lag = 0, by_call = 0;
  float last_flush_ms = 0.f;
};

namespace {

void touch(kdbgpu_refresher *r) {  // caller holds r->mu
  if (!r->has_pending) {
    r->has_pending = true;
    r->oldest = Clock::now();
    r->cv.notify_all();
  }
}

int flush_locked(kdbgpu_refresher *r, int why);

// Nothing unwinds across the C boundary or out of the clock thread: a std::bad_alloc while the batch is assembled is
// reported like a failed mirror call (last_error; the batch is gone, see kdbgpu_refresher_flush in the header).
int flush_impl(kdbgpu_refresher *r, int why) {
  std::lock_guard<std::mutex> fl(r->flush_mu);
  try {
    return flush_locked(r, why);
  } catch (...) {
    std::lock_guard<std::mutex> lk(r->mu);
    r->last_error = KDBGPU_ERR_NOMEM;
    return KDBGPU_ERR_NOMEM;
  }
}

int flush_locked(kdbgpu_refresher *r, int why) {
  std::map<uint32_t, NewNode> new_nodes;
  std::map<std::pair<uint32_t, int>, std::vector<uint32_t>> rows;
  std::vector<uint32_t> removed;
  std::vector<uint64_t> deleted;
  bool deleted_dirty, entry_dirty;
  uint32_t entry;
  int max_level;
  {
    std::lock_guard<std::mutex> lk(r->mu);
    if (!r->has_pending) return KDBGPU_OK;
    new_nodes.swap(r->new_nodes);
    rows.swap(r->rows);
    removed.swap(r->removed);
    deleted_dirty = r->deleted_dirty;
    if (deleted_dirty) deleted = r->deleted;
    entry_dirty = r->entry_dirty;
    entry = r->entry;
    max_level = r->max_level;
    r->deleted_dirty = r->entry_dirty = false;
    r->has_pending = false;
  }
  const auto t0 = Clock::now();
  int rc = KDBGPU_OK;
  // 1. new nodes: ids continue the mirror's count; ids the host skipped stay nil (level -1)
  if (!new_nodes.empty()) {
    const uint32_t first = kdbgpu_index_count(r->h) + 1;
    const uint32_t last = new_nodes.rbegin()->first;
    if (new_nodes.begin()->first < first) rc = KDBGPU_ERR_INVALID;  // a node the mirror already holds
    if (rc == KDBGPU_OK) {
      std::vector<int32_t> levels((size_t)(last - first + 1), -1);
      for (auto &kv : new_nodes) levels[kv.first - first] = kv.second.level;
      // vectors travel in runs of consecutive ids
      auto it = new_nodes.begin();
      std::vector<unsigned char> run;
      while (rc == KDBGPU_OK && it != new_nodes.end()) {
        const uint32_t run_first = it->first;
        uint32_t next = run_first;
        run.clear();
        while (it != new_nodes.end() && it->first == next) {
          run.insert(run.end(), it->second.row.begin(), it->second.row.end());
          ++next;
          ++it;
        }
        rc = kdbgpu_upload_rows_raw(r->h, run_first, next - run_first, run.data());
      }
      if (rc == KDBGPU_OK) rc = kdbgpu_register_nodes(r->h, first, (uint32_t)levels.size(), levels.data());
    }
  }
  // 2. adjacency rows, one patch call
  if (rc == KDBGPU_OK && !rows.empty()) {
    std::vector<uint32_t> ids, nbrs;
    std::vector<int32_t> lv;
    std::vector<uint64_t> off(1, 0);
    for (auto &kv : rows) {
      ids.push_back(kv.first.first);
      lv.push_back(kv.first.second);
      nbrs.insert(nbrs.end(), kv.second.begin(), kv.second.end());
      off.push_back(nbrs.size());
    }
    if (nbrs.empty()) nbrs.push_back(0);
    rc = kdbgpu_patch_rows(r->h, (uint32_t)ids.size(), ids.data(), lv.data(), off.data(), nbrs.data());
  }
  // 3. Vacuum's physical cleanup, soft deletes, entry point
  if (rc == KDBGPU_OK && !removed.empty()) rc = kdbgpu_remove_nodes(r->h, (uint32_t)removed.size(), removed.data());
  if (rc == KDBGPU_OK && deleted_dirty) rc = kdbgpu_set_deleted(r->h, deleted.data(), deleted.size());
  if (rc == KDBGPU_OK && entry_dirty) rc = kdbgpu_set_entry(r->h, entry, max_level);
  const float ms = std::chrono::duration<float, std::milli>(Clock::now() - t0).count();
  {
    std::lock_guard<std::mutex> lk(r->mu);
    r->flushes++;
    r->rows_applied += rows.size();
    r->nodes_applied += new_nodes.size();
    (why == 0 ? r->by_rows : why == 1 ? r->by_lag : r->by_call)++;
    r->last_flush_ms = ms;
    if (rc != KDBGPU_OK) r->last_error = rc;
  }
  return rc;
}

void clock_main(kdbgpu_refresher *r) {
  std::unique_lock<std::mutex> lk(r->mu);
  while (!r->closing) {
    if (!r->has_pending) {
      r->cv.wait(lk);
      continue;
    }
    const auto due = r->oldest + std::chrono::milliseconds(r->max_lag_ms);
    if (Clock::now() < due) {
      r->cv.wait_until(lk, due);
      continue;
    }
    lk.unlock();
    (void)flush_impl(r, 1);
    lk.lock();
  }
}

}  // namespace

extern "C" {

int kdbgpu_refresher_create(kdbgpu_index *h, uint32_t max_pending_rows, uint32_t max_lag_ms, kdbgpu_refresher **out) {
  if (!h || !out) return KDBGPU_ERR_INVALID;
  *out = nullptr;
  kdbgpu_refresher *r = new (std::nothrow) kdbgpu_refresher();
  if (!r) return KDBGPU_ERR_NOMEM;
  r->h = h;
  r->max_pending_rows = max_pending_rows ? max_pending_rows : 4096;
  r->max_lag_ms = max_lag_ms;
  const int prec = kdbgpu_index_precision(h);
  r->row_bytes = (size_t)kdbgpu_index_dim(h) * (prec == KDBGPU_PRECISION_F32 ? 4 : prec == KDBGPU_PRECISION_F16 ? 2 : 1);
  try {
    if (max_lag_ms) r->clock_thread = std::thread(clock_main, r);
  } catch (...) {
    delete r;
    return KDBGPU_ERR_NOMEM;
  }
  *out = r;
  return KDBGPU_OK;
}

int kdbgpu_refresher_destroy(kdbgpu_refresher *r) {
  if (!r) return KDBGPU_OK;
  {
    std::lock_guard<std::mutex> lk(r->mu);
    r->closing = true;
  }
  r->cv.notify_all();
  if (r->clock_thread.joinable()) r->clock_thread.join();
  const int rc = flush_impl(r, 2);  // nothing queued is lost
  delete r;
  return rc;
}

int kdbgpu_refresher_add_node(kdbgpu_refresher *r, uint32_t id, int level, const void *row_raw) {
  if (!r || !row_raw || id == 0 || level < 0 || level > 120) return KDBGPU_ERR_INVALID;
  try {
    NewNode nn;
    nn.level = level;
    nn.row.assign(static_cast<const unsigned char *>(row_raw), static_cast<const unsigned char *>(row_raw) + r->row_bytes);
    std::lock_guard<std::mutex> lk(r->mu);
    r->new_nodes[id] = std::move(nn);
    touch(r);
    return KDBGPU_OK;
  } catch (...) {
    return KDBGPU_ERR_NOMEM;
  }
}

int kdbgpu_refresher_set_row(kdbgpu_refresher *r, uint32_t id, int level, const uint32_t *nbrs, uint32_t count) {
  if (!r || id == 0 || level < 0 || (count && !nbrs)) return KDBGPU_ERR_INVALID;
  bool full = false;
  try {
    std::lock_guard<std::mutex> lk(r->mu);
    r->rows[std::make_pair(id, level)].assign(nbrs, nbrs + count);  // last write wins
    r->rows_queued++;
    touch(r);
    full = r->rows.size() >= r->max_pending_rows;
  } catch (...) {
    return KDBGPU_ERR_NOMEM;
  }
  return full ? flush_impl(r, 0) : KDBGPU_OK;
}

int kdbgpu_refresher_remove_node(kdbgpu_refresher *r, uint32_t id) {
  if (!r || id == 0) return KDBGPU_ERR_INVALID;
  try {
    std::lock_guard<std::mutex> lk(r->mu);
    r->removed.push_back(id);
    // rows queued FOR the node die with it
    for (auto it = r->rows.lower_bound(std::make_pair(id, 0)); it != r->rows.end() && it->first.first == id;) it = r->rows.erase(it);
    touch(r);
    return KDBGPU_OK;
  } catch (...) {
    return KDBGPU_ERR_NOMEM;
  }
}

int kdbgpu_refresher_set_deleted(kdbgpu_refresher *r, uint32_t id, int is_deleted) {
  if (!r || id == 0) return KDBGPU_ERR_INVALID;
  try {
    std::lock_guard<std::mutex> lk(r->mu);
    const size_t w = id / 64;
    if (r->deleted.size() <= w) r->deleted.resize(w + 1, 0);
    const uint64_t bit = 1ull << (id % 64);
    const uint64_t old = r->deleted[w];
    r->deleted[w] = is_deleted ? (old | bit) : (old & ~bit);
    if (r->deleted[w] != old) {
      r->deleted_dirty = true;
      touch(r);
    }
    return KDBGPU_OK;
  } catch (...) {
    return KDBGPU_ERR_NOMEM;
  }
}

int kdbgpu_refresher_set_entry(kdbgpu_refresher *r, uint32_t entry, int max_level) {
  if (!r) return KDBGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(r->mu);
  if (!r->entry_dirty || r->entry != entry || r->max_level != max_level) {
    r->entry = entry;
    r->max_level = max_level;
    r->entry_dirty = true;
    touch(r);
  }
  return KDBGPU_OK;
}

int kdbgpu_refresher_flush(kdbgpu_refresher *r) {
  if (!r) return KDBGPU_ERR_INVALID;
  return flush_impl(r, 2);
}

int kdbgpu_refresher_stats(kdbgpu_refresher *r, kdbgpu_refresher_stats_t *out) {
  if (!r || !out) return KDBGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(r->mu);
  memset(out, 0, sizeof *out);
  out->pending_rows = r->rows.size();
  out->pending_nodes = r->new_nodes.size();
  out->flushes = r->flushes;
  out->flushes_by_rows = r->by_rows;
  out->flushes_by_lag = r->by_lag;
  out->flushes_by_call = r->by_call;
  out->rows_queued = r->rows_queued;
  out->rows_applied = r->rows_applied;
  out->nodes_applied = r->nodes_applied;
  out->oldest_pending_ms = r->has_pending ? std::chrono::duration<float, std::milli>(Clock::now() - r->oldest).count() : 0.f;
  out->last_flush_ms = r->last_flush_ms;
  out->last_error = r->last_error;
  return KDBGPU_OK;
}

}  // extern "C"
