// b2w_replay.cuh -- exact emulation of a sequential float32 cumulative sum over runs of identical addends
// (numba/np/arraymath.py:384-405 as used by pecanpy.py:556-557), shared by the unweighted SparseOTF kernels.
#pragma once
#include "b2w_common.cuh"

__device__ __forceinline__ double pow2_double(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// Adds `fo` to the running f32 prefix `cdf` n times, exactly as n sequential __fadd_rn would, but
// jumping through each binade of cdf in O(1): while cdf stays inside one binade its grid is
// g = ulp(cdf), cdf is a multiple of g, and RN(cdf + fo) = cdf + RN_g(fo) whenever fo/g is not a
// rounding tie (ties and binade crossings fall back to genuine single additions).  `k` is the
// index of the next element; returns true and sets `choice` at the first element with !(cdf < u).
__device__ __forceinline__ bool advance_run(float& cdf, uint32_t& k, uint32_t n, const float fo, const double u,
                                            uint32_t& choice) {
  while (n > 0) {
    const uint32_t bits = __float_as_uint(cdf);
    const int ex = (int)((bits >> 23) & 0xFFu);
    if (ex >= 1 && ex < 255) {
      const double t = (double)fo * pow2_double(150 - ex);            // fo / g, exact
      if (t < 8388608.0) {
        const double tr = rint(t);
        if (fabs(t - tr) != 0.5) {
          const uint32_t R = (uint32_t)tr;
          if (R == 0) { k += n; return false; }                       // fo is absorbed: cdf never moves again
          const uint32_t Cm = (bits & 0x7FFFFFu) | 0x800000u;         // cdf / g in [2^23, 2^24)
          const uint32_t imax = (0xFFFFFFu - Cm) / R;                 // additions that stay below 2^24 g
          const uint32_t steps = min(n, imax);
          if (steps > 0) {
            const uint32_t Cn = Cm + steps * R;
            const float cdf_n = __uint_as_float((bits & 0xFF800000u) | (Cn & 0x7FFFFFu));
            if (!((double)cdf_n < u)) {
              const double U = u * pow2_double(150 - ex);             // u / g, exact scaling
              double di = ceil((U - (double)Cm) / (double)R);
              uint32_t i = di < 1.0 ? 1u : (di > (double)steps ? steps : (uint32_t)di);
              while (i > 1 && (double)(Cm + (i - 1) * R) >= U) --i;
              while ((double)(Cm + i * R) < U) ++i;
              choice = k + i - 1;
              return true;
            }
            cdf = cdf_n; k += steps; n -= steps;
            if (n == 0) return false;
          }
        }
      }
    }
    cdf = __fadd_rn(cdf, fo);                                          // genuine addition
    if (!((double)cdf < u)) { choice = k; return true; }
    ++k; --n;
  }
  return false;
}

