// Host build of pecanpy_b200/csrc/b2w_rowout.cuh (test infrastructure; the product runs it inside CUDA kernels):
// the 32 lanes of one warp are executed one after the other, every global store is recorded.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#define B2W_HOST_TEST 1
#define __device__
#define __forceinline__ inline
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct { unsigned x; } threadIdx;
static inline void __syncwarp() {}
struct WalkParams { uint32_t* out; uint64_t ld_out; long long mirror_delta[7]; int n_mirrors; };

// every buffer (local matrix + mirrors) with its per-word store counters
struct Buf { uint32_t* base; size_t words; uint32_t* count; };
static std::vector<Buf> g_bufs;
static int g_stray = 0;
static void record(uint32_t* p, uint32_t v) {
  for (auto& b : g_bufs)
    if (p >= b.base && p < b.base + b.words) { *p = v; ++b.count[p - b.base]; return; }
  ++g_stray;                                               // a store outside every matrix
}
static inline void __stcs(uint32_t* p, uint32_t v) { record(p, v); }
static inline void __stcs(uint4* p, uint4 v) {
  if (reinterpret_cast<uintptr_t>(p) & 15) ++g_stray;      // a misaligned 16-byte store would fault on the device
  uint32_t* q = reinterpret_cast<uint32_t*>(p);
  record(q, v.x); record(q + 1, v.y); record(q + 2, v.z); record(q + 3, v.w);
}
#include "../pecanpy_b200/csrc/b2w_rowout.cuh"

static inline uint32_t value_of(uint32_t row, uint32_t j) { return 0x10000u * (row + 1) + j + 1; }

// One warp writes `n_rows` (<= 32) rows of L + 2 words (leading dimension ld, first row `row0`) into a local matrix and
// `n_mirrors` mirrors whose bases all sit `phase` words past a 32-byte boundary, the way walk_uw_edge_kernel<., true,
// true> drives WarpRowTile.  values/counts: [1 + n_mirrors][total_rows * ld], filled for the caller to check.
extern "C" int h_rowtile_run(uint32_t L, uint32_t ld, uint32_t row0, uint32_t n_rows, uint32_t total_rows, int n_mirrors,
                             uint32_t phase, uint32_t* values, uint32_t* counts) {
  const size_t words = (size_t)total_rows * ld;
  std::vector<uint32_t*> raw;
  g_bufs.clear();
  g_stray = 0;
  for (int b = 0; b <= n_mirrors; ++b) {
    uint32_t* r = static_cast<uint32_t*>(aligned_alloc(64, (words + 16) * 4 + 64));
    raw.push_back(r);
    uint32_t* base = r + phase;
    for (size_t t = 0; t < words; ++t) base[t] = 0xDEADBEEFu;
    g_bufs.push_back(Buf{base, words, counts + (size_t)b * words});
    memset(g_bufs.back().count, 0, words * 4);
  }
  WalkParams P{};
  P.out = g_bufs[0].base;
  P.ld_out = ld;
  P.n_mirrors = n_mirrors;
  for (int q = 0; q < n_mirrors; ++q) P.mirror_delta[q] = g_bufs[q + 1].base - g_bufs[0].base;
  std::vector<uint32_t> smem(33 * 32, 0);
  WarpRowTile tile[32];
  for (uint32_t lane = 0; lane < 32; ++lane) { threadIdx.x = lane; tile[lane].begin(smem.data(), row0, n_rows); }
  auto put_all = [&](uint32_t j) {
    for (uint32_t lane = 0; lane < n_rows; ++lane) { threadIdx.x = lane; tile[lane].put(j, value_of(row0 + lane, j)); }
  };
  auto flush_all = [&](uint32_t have, bool last) {
    for (uint32_t lane = 0; lane < 32; ++lane) { threadIdx.x = lane; tile[lane].flush(P, have, last); }
  };
  put_all(0);
  uint32_t since = 1;
  for (uint32_t j = 1; j <= L; ++j) {
    put_all(j);
    if (++since == MIRROR_PERIOD) { flush_all(j + 1, false); since = 0; }
  }
  put_all(L + 1);
  flush_all(L + 2, true);
  for (int b = 0; b <= n_mirrors; ++b) memcpy(values + (size_t)b * words, g_bufs[b].base, words * 4);
  for (auto r : raw) free(r);
  return g_stray;
}

// The same rows through RowWriter (one lane per row, sector stores), THREADS = 32.
extern "C" int h_rowwriter_run(uint32_t L, uint32_t ld, uint32_t row0, uint32_t n_rows, uint32_t total_rows, uint32_t phase,
                               uint32_t* values, uint32_t* counts) {
  const size_t words = (size_t)total_rows * ld;
  uint32_t* r = static_cast<uint32_t*>(aligned_alloc(64, (words + 16) * 4 + 64));
  uint32_t* base = r + phase;
  for (size_t t = 0; t < words; ++t) base[t] = 0xDEADBEEFu;
  g_bufs.clear();
  g_stray = 0;
  g_bufs.push_back(Buf{base, words, counts});
  memset(counts, 0, words * 4);
  std::vector<uint32_t> smem(8 * 32, 0);
  for (uint32_t lane = 0; lane < n_rows; ++lane) {
    threadIdx.x = lane;
    RowWriter<32> row;
    row.begin(base + (size_t)(row0 + lane) * ld, smem.data());
    for (uint32_t j = 0; j < L + 2; ++j) row.push(j, value_of(row0 + lane, j));
    row.finish(L + 2);
  }
  memcpy(values, base, words * 4);
  free(r);
  return g_stray;
}
