// batcher_driver.cpp — TEST / BENCH INFRASTRUCTURE: the reference's call shape at load.
//
// KektorDB answers every search on its own goroutine with one blocking
// idx.SearchWithScores(query, k, allowList, efSearch) call (reference pkg/engine/ops.go:1006).  This
// driver plays those callers natively: n_threads OS threads, each issuing blocking ONE-QUERY calls to
// kdbgpu_batcher_search (what the cgo shim would do), so the micro-batcher can be
// timed without the Python interpreter between the callers and the C ABI.
//
// Build: g++ -O2 -shared -fPIC -pthread -I include tests/native/batcher_driver.cpp -o tests/native/libbatcher_driver.so
// (resolves the kdbgpu_* symbols from the already-loaded libkektordb_gpu.so at run time).
#include <atomic>
#include <cstdint>
#include <chrono>
#include <thread>
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
