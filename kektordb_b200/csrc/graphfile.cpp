// graphfile.cpp — topology hand-off through a flat binary file (second half).
//
// The reference persists the graph inside its gob-encoded .kdb snapshot (pkg/core/core.go:177-306; node data via
// (*Index).SnapshotData, hnsw_index.go:3064-3150) — Go maps that only Go can decode.  At 1 M - 10 M nodes the shim
// should not rebuild the CSR arrays of kdbgpu_set_graph in Go slices on every start either.  So the topology
// travels as a sidecar file next to the arena: the shim (or kdbgpu_save_graph_file, after a build on the
// device) writes it once per snapshot, kdbgpu_set_graph_file maps it and stages it — no Go-side arrays at all.
//
// Layout (little endian), every section 8-byte aligned:
//   header, 64 bytes: u32 magic 0x4742444B ("KDBG"), u32 version 1, u32 n (nodeCounter), u32 entry (entrypointID),
//                     i32 max_level, u32 m, u64 n_rows, u64 n_edges, zero padding
//   levels   i32[n + 1]      len(Connections) - 1 of node i, -1 = nil slot (index 0 is nil)
//   node_row u64[n + 2]      node i owns rows node_row[i] .. node_row[i+1]-1, level 0 first
//   row_off  u64[n_rows + 1] row r lists nbrs[row_off[r] .. row_off[r+1]-1]
//   nbrs     u32[n_edges]    neighbour ids in the reference's order
// i.e. exactly the arguments of kdbgpu_set_graph.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/kektordb_gpu.h"

namespace kdb {
int set_error(int code, const char *fmt, ...);
}
using kdb::set_error;

namespace {

constexpr uint32_t kMagic = 0x4742444Bu;
constexpr uint32_t kVersion = 1;
constexpr size_t kHeader = 64;

struct Header {
  uint32_t magic, version, n, entry;
  int32_t max_level;
  uint32_t m;
  uint64_t n_rows, n_edges;
  unsigned char pad[kHeader - 40];
};
static_assert(sizeof(Header) == kHeader, "graph file header is 64 bytes");

size_t pad8(size_t x) { return (x + 7) & ~(size_t)7; }

struct Sections {
  size_t o_levels, o_node_row, o_row_off, o_nbrs, total;
};
Sections sections(uint32_t n, uint64_t n_rows, uint64_t n_edges) {
  Sections s;
  s.o_levels = kHeader;
  s.o_node_row = s.o_levels + pad8(((size_t)n + 1) * sizeof(int32_t));
  s.o_row_off = s.o_node_row + ((size_t)n + 2) * sizeof(uint64_t);
  s.o_nbrs = s.o_row_off + ((size_t)n_rows + 1) * sizeof(uint64_t);
  s.total = s.o_nbrs + pad8((size_t)n_edges * sizeof(uint32_t));
  return s;
}

bool write_all(int fd, const void *p, size_t n) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  while (n) {
    ssize_t w = write(fd, b, n);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    b += w;
    n -= (size_t)w;
  }
  return true;
}

bool write_padded(int fd, const void *p, size_t n) {
  static const unsigned char zeros[8] = {0};
  return write_all(fd, p, n) && write_all(fd, zeros, pad8(n) - n);
}

}  // namespace

extern "C" {

int kdbgpu_graph_file_write(const char *path, uint32_t n, int m, const int32_t *levels, const uint64_t *node_row,
                            const uint64_t *row_off, const uint32_t *nbrs, uint32_t entry, int max_level) {
  if (!path || !levels || !node_row || !row_off) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  const uint64_t n_rows = node_row[(size_t)n + 1];
  const uint64_t n_edges = row_off[n_rows];
  if (n_edges && !nbrs) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  Header hd;
  memset(&hd, 0, sizeof hd);
  hd.magic = kMagic;
  hd.version = kVersion;
  hd.n = n;
  hd.entry = entry;
  hd.max_level = max_level;
  hd.m = (uint32_t)m;
  hd.n_rows = n_rows;
  hd.n_edges = n_edges;
  const std::string tmp = std::string(path) + ".tmp";
  const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return set_error(KDBGPU_ERR_INVALID, "%s: %s", tmp.c_str(), strerror(errno));
  const bool ok = write_all(fd, &hd, sizeof hd) && write_padded(fd, levels, ((size_t)n + 1) * sizeof(int32_t)) &&
                  write_all(fd, node_row, ((size_t)n + 2) * sizeof(uint64_t)) &&
                  write_all(fd, row_off, ((size_t)n_rows + 1) * sizeof(uint64_t)) &&
                  write_padded(fd, nbrs, (size_t)n_edges * sizeof(uint32_t));
  int e = errno;
  // the rename below must never publish a file whose pages did not reach the disk (a crash would leave a sidecar
  // with a valid header and torn sections)
  bool synced = ok && fsync(fd) == 0;
  if (ok && !synced) e = errno;
  close(fd);
  if (!ok || !synced) {
    unlink(tmp.c_str());
    return set_error(KDBGPU_ERR_INVALID, "%s: %s", tmp.c_str(), strerror(e));
  }
  if (rename(tmp.c_str(), path) != 0) {  // readers never see a half-written file
    unlink(tmp.c_str());
    return set_error(KDBGPU_ERR_INVALID, "%s: %s", path, strerror(errno));
  }
  return KDBGPU_OK;
}

int kdbgpu_save_graph_file(kdbgpu_index *h, const char *path) {
  if (!h || !path) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  uint32_t n = 0, entry = 0;
  uint64_t n_rows = 0, n_edges = 0;
  int max_level = -1;
  int rc = kdbgpu_get_graph_sizes(h, &n, &n_rows, &n_edges, &entry, &max_level);
  if (rc) return rc;
  try {
    std::vector<int32_t> levels((size_t)n + 1);
    std::vector<uint64_t> node_row((size_t)n + 2), row_off((size_t)n_rows + 1);
    std::vector<uint32_t> nbrs((size_t)n_edges + 1);
    rc = kdbgpu_get_graph(h, levels.data(), node_row.data(), row_off.data(), nbrs.data());
    if (rc) return rc;
    return kdbgpu_graph_file_write(path, n, kdbgpu_index_m(h), levels.data(), node_row.data(), row_off.data(), nbrs.data(),
                                   entry, max_level);
  } catch (...) {
    return set_error(KDBGPU_ERR_NOMEM, "out of host memory reading the topology back");
  }
}

int kdbgpu_graph_file_probe(const char *path, uint32_t *n, int *m, uint64_t *n_rows, uint64_t *n_edges, uint32_t *entry,
                            int *max_level) {
  if (!path) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return set_error(KDBGPU_ERR_INVALID, "%s: %s", path, strerror(errno));
  Header hd;
  const ssize_t r = pread(fd, &hd, sizeof hd, 0);
  struct stat st;
  const int sr = fstat(fd, &st);
  close(fd);
  if (r != (ssize_t)sizeof hd || sr != 0) return set_error(KDBGPU_ERR_INVALID, "%s: short header", path);
  if (hd.magic != kMagic) return set_error(KDBGPU_ERR_INVALID, "%s is not a graph file (magic mismatch)", path);
  if (hd.version != kVersion) return set_error(KDBGPU_ERR_INVALID, "%s: unsupported version %u", path, hd.version);
  if ((uint64_t)st.st_size < sections(hd.n, hd.n_rows, hd.n_edges).total)
    return set_error(KDBGPU_ERR_INVALID, "%s is truncated: %lld bytes, %zu expected", path, (long long)st.st_size,
                     sections(hd.n, hd.n_rows, hd.n_edges).total);
  if (n) *n = hd.n;
  if (m) *m = (int)hd.m;
  if (n_rows) *n_rows = hd.n_rows;
  if (n_edges) *n_edges = hd.n_edges;
  if (entry) *entry = hd.entry;
  if (max_level) *max_level = hd.max_level;
  return KDBGPU_OK;
}

int kdbgpu_set_graph_file(kdbgpu_index *h, const char *path) {
  if (!h || !path) return set_error(KDBGPU_ERR_INVALID, "NULL argument");
  uint32_t n = 0, entry = 0;
  uint64_t n_rows = 0, n_edges = 0;
  int m = 0, max_level = -1;
  int rc = kdbgpu_graph_file_probe(path, &n, &m, &n_rows, &n_edges, &entry, &max_level);
  if (rc) return rc;
  if (m != kdbgpu_index_m(h)) return set_error(KDBGPU_ERR_INVALID, "%s was written for M = %d, the index has M = %d", path, m, kdbgpu_index_m(h));
  const Sections s = sections(n, n_rows, n_edges);
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return set_error(KDBGPU_ERR_INVALID, "%s: %s", path, strerror(errno));
  void *map = mmap(nullptr, s.total, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return set_error(KDBGPU_ERR_INVALID, "mmap %s: %s", path, strerror(errno));
  (void)madvise(map, s.total, MADV_SEQUENTIAL);
  const unsigned char *b = static_cast<const unsigned char *>(map);
  const uint64_t *node_row = reinterpret_cast<const uint64_t *>(b + s.o_node_row);
  const uint64_t *row_off = reinterpret_cast<const uint64_t *>(b + s.o_row_off);
  rc = KDBGPU_OK;
  if (node_row[(size_t)n + 1] != n_rows || row_off[n_rows] != n_edges)
    rc = set_error(KDBGPU_ERR_INVALID, "%s: section sizes disagree with the header", path);
  for (uint64_t r = 0; rc == KDBGPU_OK && r < n_rows; ++r)
    if (row_off[r] > row_off[r + 1] || row_off[r + 1] > n_edges)
      rc = set_error(KDBGPU_ERR_INVALID, "%s: row offsets are not monotone at row %llu", path, (unsigned long long)r);
  for (uint32_t i = 0; rc == KDBGPU_OK && i <= n; ++i)
    if (node_row[i] > node_row[i + 1] || node_row[i + 1] > n_rows)
      rc = set_error(KDBGPU_ERR_INVALID, "%s: node rows are not monotone at node %u", path, i);
  if (rc == KDBGPU_OK)
    rc = kdbgpu_set_graph(h, n, reinterpret_cast<const int32_t *>(b + s.o_levels), node_row, row_off,
                          reinterpret_cast<const uint32_t *>(b + s.o_nbrs), entry, max_level);
  munmap(map, s.total);
  return rc;
}

}  // extern "C"
