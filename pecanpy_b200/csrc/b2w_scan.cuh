// b2w_scan.cuh -- exclusive prefix sum, in place, of one 32-bit field of an array of fixed-size records
// (block sums -> one-block scan -> block-local scan); used by the edge-index builders for their list offsets.
#pragma once
#include "b2w_common.cuh"

namespace b2w_scan {

constexpr int THREADS = 256;
constexpr int ITEMS = 16;                                            // per thread -> 4096 records per block

// field(e) = base[e * stride]  (stride in 32-bit words)
static __global__ void __launch_bounds__(THREADS) block_sums(const uint64_t count, const uint32_t* __restrict__ base,
                                                      const uint32_t stride, unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long s_w[THREADS / 32];
  const uint64_t b0 = (uint64_t)blockIdx.x * THREADS * ITEMS;
  unsigned long long acc = 0;
  for (int it = 0; it < ITEMS; ++it) {
    const uint64_t e = b0 + (uint64_t)it * THREADS + threadIdx.x;
    if (e < count) acc += base[e * stride];
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(B2W_FULL, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < THREADS / 32; ++w) t += s_w[w];
    sums[blockIdx.x] = t;
  }
}

static __global__ void __launch_bounds__(1024) scan_sums(const uint64_t nblocks, unsigned long long* __restrict__ sums,
                                                  unsigned long long* __restrict__ total) {
  // one block; every thread owns a contiguous slice
  __shared__ unsigned long long s_part[1024];
  const uint64_t per = (nblocks + 1023) / 1024;
  const uint64_t lo = min(nblocks, (uint64_t)threadIdx.x * per), hi = min(nblocks, lo + per);
  unsigned long long acc = 0;
  for (uint64_t i = lo; i < hi; ++i) acc += sums[i];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; ++t) { const unsigned long long v = s_part[t]; s_part[t] = run; run += v; }
    *total = run;
  }
  __syncthreads();
  unsigned long long run = s_part[threadIdx.x];
  for (uint64_t i = lo; i < hi; ++i) { const unsigned long long v = sums[i]; sums[i] = run; run += v; }
}

static __global__ void __launch_bounds__(THREADS) apply(const uint64_t count, uint32_t* __restrict__ base, const uint32_t stride,
                                                 const unsigned long long* __restrict__ sums) {
  // items are laid out [it][thread] inside the block, so the block-local order is it-major
  __shared__ uint32_t s_warp[ITEMS][THREADS / 32];
  __shared__ uint32_t s_itbase[ITEMS];
  const uint64_t b0 = (uint64_t)blockIdx.x * THREADS * ITEMS;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t v[ITEMS], incl[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const uint64_t e = b0 + (uint64_t)it * THREADS + threadIdx.x;
    v[it] = e < count ? base[e * stride] : 0u;
    uint32_t x = v[it];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(B2W_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
    incl[it] = x;
    if (lane == 31) s_warp[it][wib] = x;
  }
  __syncthreads();
  if (threadIdx.x < ITEMS) {
    uint32_t run = 0;
    for (int w = 0; w < THREADS / 32; ++w) { const uint32_t t = s_warp[threadIdx.x][w]; s_warp[threadIdx.x][w] = run; run += t; }
    s_itbase[threadIdx.x] = run;                                      // total of this item row
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int it = 0; it < ITEMS; ++it) { const uint32_t t = s_itbase[it]; s_itbase[it] = run; run += t; }
  }
  __syncthreads();
  const uint32_t blockbase = (uint32_t)sums[blockIdx.x];              // total < 2^32 is checked by the caller
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const uint64_t e = b0 + (uint64_t)it * THREADS + threadIdx.x;
    if (e < count) base[e * stride] = blockbase + s_itbase[it] + s_warp[it][wib] + incl[it] - v[it];
  }
}

inline uint64_t blocks(uint64_t count) { return (count + (uint64_t)THREADS * ITEMS - 1) / ((uint64_t)THREADS * ITEMS); }
inline size_t work_bytes(uint64_t count) { return (blocks(count) + 2) * sizeof(unsigned long long); }

// Exclusive scan of field `base[e * stride]`, e < count, in place; *h_total = the sum (synchronises the stream).
// `sums`: work_bytes(count) bytes of device scratch.  Returns a cudaError_t.
inline cudaError_t exclusive_scan(uint64_t count, uint32_t* base, uint32_t stride, unsigned long long* sums,
                                  unsigned long long* h_total, cudaStream_t s) {
  const uint64_t nb = blocks(count);
  unsigned long long* total = sums + nb;
  block_sums<<<(unsigned)nb, THREADS, 0, s>>>(count, base, stride, sums);
  scan_sums<<<1, 1024, 0, s>>>(nb, sums, total);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(h_total, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  if (*h_total >= 0xFFFFFFFFull) return cudaSuccess;                 // caller reports the overflow
  apply<<<(unsigned)nb, THREADS, 0, s>>>(count, base, stride, sums);
  return cudaGetLastError();
}

}  // namespace b2w_scan
