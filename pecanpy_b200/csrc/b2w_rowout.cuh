// b2w_rowout.cuh -- sector-aligned stores of a walk-matrix row by ONE lane (the lane-per-walker kernels).
//
// A lane produces one 4-byte entry of its row per step.  Stored one by one, every 32-byte sector of the matrix is
// written eight times, microseconds apart; under the cache turnover of a table-streaming kernel a partly written
// sector is evicted and written to DRAM more than once (ncu, round 1: 283 MB of DRAM writes for a 164 MB matrix).
// Here the entries are staged in shared memory (8 words per lane, [slot][thread]: conflict free) and leave as two
// 16-byte streaming stores when the sector they belong to is complete.  The sector phase comes from the ADDRESS of
// the row, so any row length / leading dimension works (rows of L + 2 = 82 words start at four different phases);
// the head of the row before its first sector boundary and the tail after the last one are written word by word.
#pragma once
#ifdef B2W_HOST_TEST
// g++ build for tests/test_rowtile_host.py (no GPU needed): tests/rowtile_harness.cpp supplies WalkParams, threadIdx,
// uint4, __stcs and __syncwarp and runs the 32 lanes of a warp one after the other
#else
#include "b2w_common.cuh"
#endif

template <int THREADS>
struct RowWriter {
  uint32_t* out;
  uint32_t* stage;     // this thread's column of the staging tile: slot s at stage[s * THREADS]
  uint32_t ph;         // position of out[0] inside its 32-byte sector, in words

  __device__ __forceinline__ void begin(uint32_t* row, uint32_t* tile) {
    out = row;
    stage = tile + threadIdx.x;
    ph = (uint32_t)((reinterpret_cast<uintptr_t>(row) >> 2) & 7u);
  }
  // entry j of the row (entries must be pushed in order j = 0, 1, 2, ...)
  __device__ __forceinline__ void push(const uint32_t j, const uint32_t v) {
    const uint32_t s = (ph + j) & 7u;
    stage[s * THREADS] = v;
    if (s == 7u) {
      if (j >= 7u) {
        const uint4 a = make_uint4(stage[0], stage[THREADS], stage[2 * THREADS], stage[3 * THREADS]);
        const uint4 b = make_uint4(stage[4 * THREADS], stage[5 * THREADS], stage[6 * THREADS], stage[7 * THREADS]);
        __stcs(reinterpret_cast<uint4*>(out + (j - 7u)), a);          // 32-byte aligned by construction
        __stcs(reinterpret_cast<uint4*>(out + (j - 3u)), b);
      } else {
        for (uint32_t t = 0; t <= j; ++t) out[t] = stage[((ph + t) & 7u) * THREADS];   // head of the row
      }
    }
  }
  // after `count` entries have been pushed: the words staged since the last complete sector
  __device__ __forceinline__ void finish(const uint32_t count) {
    const uint32_t total = ph + count;
    const uint32_t rem = total < 8u ? count : (total & 7u);
    for (uint32_t t = count - rem; t < count; ++t) out[t] = stage[((ph + t) & 7u) * THREADS];
  }
};

// Mirrors: the same rows of the walk matrices of the OTHER GPUs of the job (mapped through CUDA IPC, b2w_shared_open),
// written by the walk kernel itself over NVLink while it walks -- the all-gather fused into the kernel
// (b2w_walk_mirrored).  WalkParams::mirror_delta[q] = (peer q's matrix) - (this GPU's matrix) in 4-byte words, all
// congruent modulo 32 bytes.
//
// 32-byte stores by single lanes (RowWriter) are a poor fit for the link: measured 130-150 GB/s of egress per GPU
// (2 GPUs: +1.6 ms on a 9.8 ms kernel; 8 GPUs: 21.7 ms instead of 2.8).  WarpRowTile stages the rows of the warp's 32
// walkers in shared memory instead (a ring of 32 words per walker, padded to 33: conflict free both ways) and, every
// MIRROR_PERIOD steps, the WHOLE WARP writes each walker's new words as one coalesced store of up to 31 consecutive
// words -- whole 32-byte sectors, cut at the sector boundaries of that row's address -- to the local matrix and to every
// mirror.  Must be called by the 32 converged lanes of the warp.
constexpr uint32_t MIRROR_PERIOD = 24;      // words between flushes; + up to 7 carried words <= 31 < ring of 32

struct WarpRowTile {
  uint32_t* tile;      // this warp's [32][33] words
  uint64_t row0;       // first row of the warp (lane 0's), in rows of the local matrix
  uint32_t n;          // rows of the warp that exist (a prefix of the lanes)
  uint32_t flushed;    // `have` of the last flush (0: none yet)

  __device__ __forceinline__ void begin(uint32_t* warp_tile, const uint64_t first_row, const uint32_t n_rows) {
    tile = warp_tile; row0 = first_row; n = n_rows; flushed = 0;
  }
  __device__ __forceinline__ void put(const uint32_t j, const uint32_t v) { tile[(threadIdx.x & 31) * 33 + (j & 31)] = v; }
  // words [0, have) of every row have been put; `last`: the rows are complete
  __device__ __forceinline__ void flush(const WalkParams& P, const uint32_t have, const bool last) {
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t obase = (uint64_t)(reinterpret_cast<uintptr_t>(P.out) >> 2);
    for (uint32_t w = 0; w < n; ++w) {
      const uint64_t rb = (row0 + w) * P.ld_out;
      const uint32_t ph = (uint32_t)((obase + rb) & 7u);
      const uint32_t a = flushed ? flushed - ((ph + flushed) & 7u) : 0u;      // (flushed >= MIRROR_PERIOD > 7)
      const uint32_t b = last ? have : have - ((ph + have) & 7u);
      const uint32_t t = a + lane;
      if (t < b) {
        const uint32_t v = tile[w * 33 + (t & 31)];
        uint32_t* const o = P.out + rb + t;
        __stcs(o, v);
        for (int q = 0; q < P.n_mirrors; ++q) __stcs(o + P.mirror_delta[q], v);
      }
    }
    flushed = have;
    __syncwarp();
  }
};
