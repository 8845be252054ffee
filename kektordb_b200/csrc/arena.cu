// arena.cu — arena -> HBM staging: reads the reference's vector arena
// (pkg/storage/mmap/arena.go) straight into the GPU mirror, without a detour through Go slices.
//
// Format restated from arena.go:14-19, :79-118, :307-376, :378-444:
//   <dir>/arena_%04d.bin, each file DefaultChunkSize = 64 MiB;
//   64-byte header: u32 LE magic 0x4B414F4E, u32 version 1, u32 dim, u8 precision (0 f32, 1 f16, 2 int8),
//   rest zero; then vecsPerChk = (64 MiB - 64) / vectorSize vectors of vectorSize = dim * elem bytes;
//   logical internal id -> physical slot through ArenaState.SlotTable (0xFFFFFFFF = unallocated);
//   slot p lives in chunk p / vecsPerChk at byte 64 + (p % vecsPerChk) * vectorSize.
//
// Data path: chunk file -> pinned staging buffer (pread) -> ONE async H2D copy of the whole payload
// (full PCIe rate, two buffers so the next chunk is read while the previous one travels) ->
// arena_scatter_kernel places every row of that chunk at its logical id.  int8 norms follow from the
// staged rows (computeInt8Norm, hnsw_index.go:3371-3377).
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "kdb_internal.cuh"

namespace kdb {

int arena_fail(int code, const char *fmt, ...);  // api.cu: sets the thread's last error

namespace {
constexpr size_t kChunkSize = 64ull * 1024 * 1024;  // DefaultChunkSize
constexpr uint32_t kMagic = 0x4B414F4Eu;            // ArenaMagic
constexpr uint32_t kVersion = 1;                    // ArenaVersion
constexpr size_t kHeader = 64;                      // ArenaHeaderSize

uint32_t le32(const unsigned char *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
}  // namespace

// addChunk's header validation (arena.go:346-364)
int arena_check_header(const unsigned char *hdr, uint32_t dim, int precision, const char *what) {
  if (le32(hdr) != kMagic) return arena_fail(KDBGPU_ERR_INVALID, "%s is not a valid arena (magic mismatch)", what);
  if (le32(hdr + 4) != kVersion) return arena_fail(KDBGPU_ERR_INVALID, "%s unsupported version %u", what, le32(hdr + 4));
  if (le32(hdr + 8) != dim)
    return arena_fail(KDBGPU_ERR_INVALID, "%s dimension mismatch: expected %u, got %u", what, dim, le32(hdr + 8));
  if ((int)hdr[12] != precision)
    return arena_fail(KDBGPU_ERR_INVALID, "%s precision mismatch: expected %d, got %d", what, precision, (int)hdr[12]);
  return KDBGPU_OK;
}

size_t arena_chunk_size() { return kChunkSize; }
size_t arena_header_size() { return kHeader; }
uint32_t arena_vecs_per_chunk(uint32_t vector_bytes) { return (uint32_t)((kChunkSize - kHeader) / vector_bytes); }

// highest chunk id present in dir (loadExistingChunks, arena.go:283-305), -1 if none
int arena_max_chunk(const char *dir, std::string *err) {
  int max_id = -1;
  for (int id = 0; id < 10000; ++id) {
    char name[32];
    snprintf(name, sizeof name, "/arena_%04d.bin", id);
    struct stat st;
    if (stat((std::string(dir) + name).c_str(), &st) == 0 && S_ISREG(st.st_mode)) max_id = id;
    else if (id > max_id + 64) break;  // chunk ids are dense from 0; allow a few holes (dropped chunks)
  }
  (void)err;
  return max_id;
}

// reads up to `want` payload bytes of chunk `id` into dst (pinned); returns bytes read or -1
long arena_read_chunk(const char *dir, int id, unsigned char *hdr, unsigned char *dst, size_t want, std::string *err) {
  char name[32];
  snprintf(name, sizeof name, "/arena_%04d.bin", id);
  const std::string path = std::string(dir) + name;
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    *err = path + ": " + strerror(errno);
    return -1;
  }
  ssize_t r = pread(fd, hdr, kHeader, 0);
  if (r != (ssize_t)kHeader) {
    *err = path + ": short header";
    close(fd);
    return -1;
  }
  size_t got = 0;
  while (got < want) {
    r = pread(fd, dst + got, want - got, (off_t)(kHeader + got));
    if (r < 0) {
      *err = path + ": " + strerror(errno);
      close(fd);
      return -1;
    }
    if (r == 0) break;  // a truncated file: the rest of the payload is zero pages in the reference's mmap
    got += (size_t)r;
  }
  if (got < want) memset(dst + got, 0, want - got);
  close(fd);
  return (long)want;
}

}  // namespace kdb
