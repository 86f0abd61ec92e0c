// b2w_replay.cuh -- exact emulation of a sequential float32 cumulative sum over runs of identical addends
// (numba/np/arraymath.py:384-405 as used by pecanpy.py:556-557), shared by the unweighted SparseOTF kernels.
#pragma once
#ifdef B2W_HOST_TEST
#include "../../tests/cuda_host_shim.h"   // g++ build of this header for tests/test_replay_host.py (no GPU needed)
#else
#include "b2w_common.cuh"
#endif

__device__ __forceinline__ double pow2_double(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// floor(a / b) for a < 2^24, 1 <= b < 2^24, a / b < 2^23 or b == 1: both operands are exact floats and the fast
// quotient (reciprocal + multiply, <= 2 ulp) is within 1 of the true one, so one fix-up in each direction makes
// it exact -- a quarter of the instructions of the 32-bit integer division.
__device__ __forceinline__ uint32_t udiv24(const uint32_t a, const uint32_t b) {
  uint32_t q = (uint32_t)__fdividef(__uint2float_rn(a), __uint2float_rn(b));
  const int r = (int)a - (int)(q * b);
  if (r < 0) --q; else if ((uint32_t)r >= b) ++q;
  return q;
}

// The reference compares the f32 prefix with the f64 uniform, `cdf[i] < u` (np.searchsorted, arraymath.py:3841-3929).
// For floats that is a float comparison with ub = the smallest float >= u:  cdf < u  <=>  cdf < ub.
__device__ __forceinline__ float upper_float(const double u) { return __double2float_ru(u); }

// Adds `fo` to the running f32 prefix `cdf` n times, exactly as n sequential __fadd_rn would, but
// jumping through each binade of cdf in O(1): while cdf stays inside one binade its grid is
// g = ulp(cdf), cdf is a multiple of g, and RN(cdf + fo) = cdf + RN_g(fo) whenever fo/g is not a
// rounding tie (ties and binade crossings fall back to genuine single additions).  `k` is the
// index of the next element; returns true and sets `choice` at the first element with !(cdf < ub).
// Precondition: cdf < ub (or cdf == 0).  Everything on the per-binade path is integer arithmetic on the bit patterns.
__device__ __forceinline__ bool advance_run(float& cdf, uint32_t& k, uint32_t n, const float fo, const float ub,
                                            uint32_t& choice) {
  const uint32_t fbits = __float_as_uint(fo);
  const int ef = (int)((fbits >> 23) & 0xFFu);
  const uint32_t Mf = (fbits & 0x7FFFFFu) | 0x800000u;                // fo = Mf 2^(ef - 150)  (ef >= 1: normal)
  const uint32_t ubits = __float_as_uint(ub);
  while (n > 0) {
    const uint32_t bits = __float_as_uint(cdf);
    const int ex = (int)((bits >> 23) & 0xFFu);
    const int s = ex - ef;                                            // fo / g = Mf / 2^s,  g = ulp(cdf) = 2^(ex - 150)
    if (ex >= 1 && ex < 255 && ef >= 1 && s >= 1) {                   // (s <= 0: fo / g >= 2^23, genuine additions)
      uint32_t R = 0;
      bool tie = false;
      if (s <= 25) {                                                  // R = RN(fo / g); s >= 26: fo / g < 1/4, absorbed
        const uint32_t half = 1u << (s - 1), low = Mf & ((half << 1) - 1u);
        R = (s <= 24 ? (Mf >> s) : 0u) + (low > half ? 1u : 0u);
        tie = low == half;
      }
      if (!tie) {
        if (R == 0) { k += n; return false; }                         // fo is absorbed: cdf never moves again
        const uint32_t Cm = (bits & 0x7FFFFFu) | 0x800000u;           // cdf / g in [2^23, 2^24)
        const uint32_t imax = udiv24(0xFFFFFFu - Cm, R);              // additions that stay below 2^24 g
        const uint32_t steps = min(n, imax);
        if (steps > 0) {
          const uint32_t Cn = Cm + steps * R;                         // < 2^24: still in this binade
          if ((ubits >> 23) == (uint32_t)ex) {                        // ub lies in this binade: the run may reach it
            const uint32_t Um = (ubits & 0x7FFFFFu) | 0x800000u;      // > Cm, because cdf < ub
            if (Cn >= Um) {
              choice = k + udiv24(Um - Cm + R - 1u, R) - 1u;          // first i >= 1 with Cm + i R >= Um
              return true;
            }
          }
          cdf = __uint_as_float((bits & 0xFF800000u) | (Cn & 0x7FFFFFu));
          k += steps; n -= steps;
          if (n == 0) return false;
        }
      }
    }
    cdf = __fadd_rn(cdf, fo);                                          // genuine addition
    if (cdf >= ub) { choice = k; return true; }
    ++k; --n;
  }
  return false;
}
