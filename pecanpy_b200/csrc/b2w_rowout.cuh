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
#include "b2w_common.cuh"

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
