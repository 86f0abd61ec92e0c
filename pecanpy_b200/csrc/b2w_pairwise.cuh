// b2w_pairwise.cuh -- NumPy's pairwise summation and mean/std, restated operation for operation.
//
// The node2vec+ noise thresholds of the reference are plain NumPy (`row.mean() + gamma * row.std()`,
// rw/sparse_rw.py:22-35, rw/dense_rw.py:11-19), and they feed every comparison of the extended bias
// (rw/sparse_rw.py:93-130), so they are part of the bit-exact contract.  NumPy's float add-reduce is
//     0 + pairwise_sum(a, n)                                     (numpy/_core/src/umath/loops_utils.h.src)
//     pairwise_sum(a, n) = sequential                            n < 8
//                        = 8 interleaved accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//                          then the n % 8 tail sequentially      n <= 128
//                        = pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2),  n2 = n/2 - (n/2) % 8
// and mean / std are (numpy/_core/_methods.py: _mean, _var, _std), all in the array's dtype,
//     mean = sum(a) / n;   std = sqrt( sum((a - mean) * (a - mean)) / n ),
// where n is an np.intp scalar, so each quotient is formed in float64 and rounded back to the dtype.
// Leaves are visited left to right and every leaf consumes its elements in index order, so the whole
// reduction runs over a forward-only element stream (`next()`), which is what lets the dense layout skip
// its zero columns without materialising the compressed row.
//
// __host__ __device__ on purpose: tests/test_pairwise_host.py compiles this header with g++ and checks it
// against NumPy itself on the build box; the CUDA kernels in b2w_thresholds.cu use the same code.
// Compile with FMA contraction off (-fmad=false / -ffp-contract=off).
#pragma once
#include <stdint.h>

#include <cmath>

#ifndef __CUDACC__
#define B2W_HD inline
#else
#define B2W_HD __host__ __device__ __forceinline__
#endif

// One leaf of the recursion (n <= 128), consuming n elements of `it` in order.
template <typename T, typename It>
B2W_HD T b2w_pairwise_leaf(It& it, uint32_t n) {
  if (n < 8) {
    T res = (T)0;
    for (uint32_t i = 0; i < n; ++i) res = res + it.next();
    return res;
  }
  T r0 = it.next(), r1 = it.next(), r2 = it.next(), r3 = it.next();
  T r4 = it.next(), r5 = it.next(), r6 = it.next(), r7 = it.next();
  const uint32_t body = n - (n % 8);
  for (uint32_t i = 8; i < body; i += 8) {
    r0 = r0 + it.next(); r1 = r1 + it.next(); r2 = r2 + it.next(); r3 = r3 + it.next();
    r4 = r4 + it.next(); r5 = r5 + it.next(); r6 = r6 + it.next(); r7 = r7 + it.next();
  }
  T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  for (uint32_t i = body; i < n; ++i) res = res + it.next();
  return res;
}

// NumPy's add.reduce of n stream elements: post-order walk of the recursion tree with an explicit stack
// (depth <= log2(n / 128) + 2 <= 32 for n < 2^32).
template <typename T, typename It>
B2W_HD T b2w_pairwise_sum(It& it, uint32_t n) {
  uint32_t len[34];
  T left[34];
  uint8_t stage[34];
  int sp = 0;
  len[0] = n; stage[0] = 0;
  T val = (T)0;
  while (sp >= 0) {
    const uint32_t m = len[sp];
    if (m <= 128) {
      val = b2w_pairwise_leaf<T, It>(it, m);
      --sp;
    } else if (stage[sp] == 0) {
      uint32_t n2 = m / 2;
      n2 -= n2 % 8;
      stage[sp] = 1;
      len[sp + 1] = n2; stage[sp + 1] = 0;
      ++sp;
    } else if (stage[sp] == 1) {
      uint32_t n2 = m / 2;
      n2 -= n2 % 8;
      left[sp] = val;
      stage[sp] = 2;
      len[sp + 1] = m - n2; stage[sp + 1] = 0;
      ++sp;
    } else {
      val = left[sp] + val;
      --sp;
    }
  }
  return (T)0 + val;                                                  // the reduction's identity comes first
}

// `mean + gamma * std` of a row given two fresh streams over it (the second one yields the elements again).
// `Make` creates a stream; T is the row dtype (float for CSR rows, double for dense rows).  An empty row gives
// NaN, as NumPy's 0/0 does.
template <typename T, typename Make>
B2W_HD void b2w_mean_std(const Make& make, uint32_t n, T& mean, T& stdev) {
  using std::sqrt;
  auto s1 = make.raw();
  const T total = b2w_pairwise_sum<T>(s1, n);
  mean = (T)((double)total / (double)n);                              // n == 0: 0/0 = NaN
  auto s2 = make.centered_sq(mean);
  const T ss = b2w_pairwise_sum<T>(s2, n);
  stdev = sqrt((T)((double)ss / (double)n));
}
