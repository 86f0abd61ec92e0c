// Host stand-ins for the CUDA intrinsics used by pecanpy_b200/csrc/b2w_replay.cuh, so that the header's integer
// arithmetic can be checked with g++ on a box without a GPU (test infrastructure only).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
using std::min;
using std::max;
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __uint2float_rn(uint32_t u) { return (float)u; }
static inline float __fadd_rn(float a, float b) { volatile float s = a + b; return s; }
// reciprocal + multiply, like the device's approximate division (so the fix-ups of udiv24 are exercised)
static inline float __fdividef(float a, float b) { volatile float r = 1.0f / b; volatile float q = a * r; return q; }
static inline float __double2float_ru(double u) {
  float f = (float)u;
  if ((double)f < u) f = std::nextafterf(f, INFINITY);
  return f;
}
static inline double __hiloint2double(int hi, int lo) {
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; std::memcpy(&d, &b, 8); return d;
}
