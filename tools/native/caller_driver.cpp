// caller_driver.cpp — LOAD GENERATOR (bench / test infrastructure): the reference's call shape at load,
// played natively so that the micro-batcher can be timed without the Python interpreter between the
// callers and the C ABI.
//
// KektorDB answers every search on its own goroutine with one blocking
// idx.SearchWithScores(query, k, allowList, efSearch) call (reference pkg/engine/ops.go:1006).
//
//   kdb_run_callers  n_threads OS threads, each issuing blocking ONE-QUERY calls to kdbgpu_batcher_search — one
//                    OS thread per in-flight query, what goroutines blocked inside cgo would cost;
//   kdb_run_async    the shape INTEGRATION.md §3 gives the Go shim: n_submitters threads play the request
//                    goroutines' cgo calls (kdbgpu_batcher_submit returns at once), ONE dispatcher thread
//                    loops in kdbgpu_batcher_poll and hands every finished ticket to kdbgpu_batcher_take.
//                    `window` queries are kept in flight in total.  OS threads: n_submitters + 1 (+ the
//                    batcher's own workers).
//
// Build: g++ -O2 -shared -fPIC -pthread -I include tools/native/caller_driver.cpp -o tools/native/libcaller_driver.so
// (resolves the kdbgpu_* symbols from the already-loaded libkektordb_gpu.so at run time).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "kektordb_gpu.h"

extern "C" int kdb_run_callers(kdbgpu_batcher *b, const float *queries, uint32_t nq, int dim, int k, int ef_search,
                               int n_threads, uint32_t *out_ids, double *out_scores, uint32_t *out_counts,
                               double *seconds) {
  if (!b || !queries || n_threads <= 0) return -1;
  std::atomic<uint32_t> next{0};
  std::atomic<int> first_error{0};
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> threads;
  threads.reserve((size_t)n_threads);
  for (int t = 0; t < n_threads; ++t)
    threads.emplace_back([&]() {
      ready.fetch_add(1);
      while (!go.load(std::memory_order_acquire)) std::this_thread::yield();  // all callers exist before the clock starts
      for (;;) {
        const uint32_t i = next.fetch_add(1);  // every caller takes the next pending request
        if (i >= nq) break;
        const int rc = kdbgpu_batcher_search(b, queries + (size_t)i * dim, k, ef_search, nullptr, 0,
                                             out_ids + (size_t)i * k, out_scores + (size_t)i * k, out_counts + i);
        if (rc != 0) {
          int expected = 0;
          first_error.compare_exchange_strong(expected, rc);
        }
      }
    });
  while (ready.load() < n_threads) std::this_thread::yield();
  const auto t0 = std::chrono::steady_clock::now();
  go.store(true, std::memory_order_release);
  for (auto &th : threads) th.join();
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return first_error.load();
}

namespace {
struct TicketMap {  // ticket -> query index, sharded so that submitters and the dispatcher rarely meet
  static constexpr int kShards = 64;
  std::mutex mu[kShards];
  std::unordered_map<uint64_t, uint32_t> m[kShards];
  static int shard(uint64_t t) { return (int)((t * 0x9e3779b97f4a7c15ull) >> 58); }
  void put(uint64_t t, uint32_t i) {
    const int s = shard(t);
    std::lock_guard<std::mutex> lk(mu[s]);
    m[s][t] = i;
  }
  bool pop(uint64_t t, uint32_t *i) {
    const int s = shard(t);
    std::lock_guard<std::mutex> lk(mu[s]);
    auto it = m[s].find(t);
    if (it == m[s].end()) return false;
    *i = it->second;
    m[s].erase(it);
    return true;
  }
};
}  // namespace

extern "C" int kdb_run_async(kdbgpu_batcher *b, const float *queries, uint32_t nq, int dim, int k, int ef_search,
                             int n_submitters, uint32_t window, uint32_t *out_ids, double *out_scores,
                             uint32_t *out_counts, double *seconds) {
  if (!b || !queries || n_submitters <= 0 || window == 0) return -1;
  TicketMap tickets;
  std::atomic<uint32_t> next{0}, outstanding{0}, finished{0};
  std::atomic<int> first_error{0};
  std::atomic<bool> go{false};
  auto note = [&](int rc) {
    if (rc != 0) {
      int expected = 0;
      first_error.compare_exchange_strong(expected, rc);
    }
  };
  std::vector<std::thread> threads;
  for (int t = 0; t < n_submitters; ++t)
    threads.emplace_back([&]() {
      while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
      for (;;) {
        if (outstanding.load(std::memory_order_relaxed) >= window) {  // the host's own admission control
          std::this_thread::yield();
          continue;
        }
        const uint32_t i = next.fetch_add(1);
        if (i >= nq) break;
        outstanding.fetch_add(1);
        uint64_t tk = 0;
        const int rc = kdbgpu_batcher_submit(b, queries + (size_t)i * dim, k, ef_search, nullptr, 0, 0, &tk);
        if (rc != 0) {
          note(rc);
          out_counts[i] = 0;
          outstanding.fetch_sub(1);
          finished.fetch_add(1);
          continue;
        }
        tickets.put(tk, i);
      }
    });
  std::thread dispatcher([&]() {
    std::vector<uint64_t> done(4096);
    while (finished.load() < nq) {
      uint32_t n = 0;
      if (kdbgpu_batcher_poll(b, done.data(), (uint32_t)done.size(), 2000, &n) != 0) break;
      for (uint32_t j = 0; j < n; ++j) {
        uint32_t i = 0;
        while (!tickets.pop(done[j], &i)) std::this_thread::yield();  // the submitter has not recorded it yet
        note(kdbgpu_batcher_take(b, done[j], out_ids + (size_t)i * k, out_scores + (size_t)i * k, out_counts + i));
        outstanding.fetch_sub(1);
        finished.fetch_add(1);
      }
    }
  });
  const auto t0 = std::chrono::steady_clock::now();
  go.store(true, std::memory_order_release);
  for (auto &th : threads) th.join();
  dispatcher.join();
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return first_error.load();
}
