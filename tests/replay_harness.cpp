// Host build of pecanpy_b200/csrc/b2w_replay.cuh (test infrastructure; the product runs it inside CUDA kernels).
#define B2W_HOST_TEST 1
#include "../pecanpy_b200/csrc/b2w_replay.cuh"

extern "C" uint32_t h_udiv24(uint32_t a, uint32_t b) { return udiv24(a, b); }
extern "C" float h_upper_float(double u) { return upper_float(u); }

// n additions of fo starting from *cdf at element index *k
extern "C" int h_advance_run(float* cdf, uint32_t* k, uint32_t n, float fo, double u, uint32_t* choice) {
  return advance_run(*cdf, *k, n, fo, upper_float(u), *choice) ? 1 : 0;
}

// the same by n genuine float additions and the reference's comparison `cdf < u` in float64
extern "C" int h_brute_run(float* cdf, uint32_t* k, uint32_t n, float fo, double u, uint32_t* choice) {
  for (uint32_t i = 0; i < n; ++i) {
    *cdf = __fadd_rn(*cdf, fo);
    if (!((double)*cdf < u)) { *choice = *k; return 1; }
    ++*k;
  }
  return 0;
}
