// b2w_membership.cuh -- group-of-G-lanes primitives and the membership phase shared by the SparseOTF kernels.
//
// A "group" is G consecutive lanes of a warp (G = 8, 16 or 32) that own one walker; all collectives
// take the group's lane mask, so the groups of a warp may diverge freely.
#pragma once
#include "b2w_common.cuh"

constexpr uint32_t B2W_NONE = 0xFFFFFFFFu;

template <int G>
struct Tile {
  int lane, tl, base;
  uint32_t mask;
  __device__ __forceinline__ Tile() {
    lane = threadIdx.x & 31;
    tl = lane & (G - 1);
    base = lane & ~(G - 1);
    mask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << base);
  }
  __device__ __forceinline__ uint32_t ballot(bool p) const {
    uint32_t b = __ballot_sync(mask, p);
    return (G == 32) ? b : ((b >> base) & ((1u << G) - 1u));
  }
  template <typename T>
  __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, G); }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ uint32_t incl_scan(uint32_t v) const {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      uint32_t t = __shfl_up_sync(mask, v, o, G);
      if (tl >= o) v += t;
    }
    return v;
  }
  __device__ __forceinline__ uint32_t sum(uint32_t v) const {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
  }
  __device__ __forceinline__ uint32_t minu(uint32_t v) const {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v = min(v, __shfl_xor_sync(mask, v, o, G));
    return v;
  }
};

// Number of elements of the sorted row[0..n) that are < x.  Branch-free, `lg` = 32 - clz(n)
// iterations (uniform across the group): pos only grows over a prefix of elements < x.
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ row, uint32_t n, uint32_t x,
                                                    uint32_t lg) {
  const uint32_t* const rowm1 = row - 1;
  uint32_t pos = 0;
  for (uint32_t step = lg ? (1u << (lg - 1)) : 0u; step; step >>= 1) {
    const uint32_t idx = pos | step;                                  // pos has no bit at or below `step`
    uint32_t v = 0xFFFFFFFFu;
    if (idx <= n) v = __ldg(rowm1 + idx);
    if (v < x) pos = idx;
  }
  return pos;
}


// Membership of the neighbours of `cur` in N(prev) as a BITMAP over the positions of row(cur)
// (rows sorted and duplicate-free, for which the reference's merge `isnotin`, rw/sparse_rw.py:142-230,
// is exactly set membership).  Searches whichever side needs fewer probes:
//   forward: every neighbour of cur looked up in row(prev)  (streams row(cur), writes whole words)
//   reverse: every neighbour of prev, and prev itself, looked up in row(cur) (touches only
//            O(deg(prev) log deg(cur)) words of row(cur); needs the bitmap zeroed first)
// `prev` itself never counts as common (rw/sparse_rw.py:84); its position is returned in `kp`
// (B2W_NONE when prev is not a neighbour of cur: directed graphs, or after a choice == deg overflow).
// Returns the number of common neighbours.  Rows of at most 32 entries return their single bitmap word in
// `word0` (in_regs = true, nothing written to `bm`); otherwise the bitmap is in `bm` and the function ends with a
// group sync so that it is visible to all lanes.
template <int G>
__device__ __forceinline__ uint32_t membership_bitmap(const Tile<G>& T, const uint32_t* __restrict__ crow,
                                                      const uint32_t d, const uint32_t* __restrict__ prow,
                                                      const uint32_t pdeg, const uint32_t prev,
                                                      uint32_t* __restrict__ bm, uint32_t& kp, uint32_t& word0, bool& in_regs) {
  const uint32_t nwords = (d + 31) >> 5;
  const uint32_t lgp = 32 - __clz(pdeg), lgd = 32 - __clz(d);
  const uint32_t fwd_cost = ((d + G - 1) / G) * (lgp + 2);
  const uint32_t rev_cost = ((pdeg + G) / G) * (lgd + 2) + (nwords + G - 1) / G;
  uint32_t m = 0;
  kp = B2W_NONE;
  // rows of cur that fit a few chunks always take the forward direction: the choice then does not depend on
  // deg(prev), so the groups of one warp stay on the same path (no divergence between walkers)
  word0 = 0u;
  in_regs = false;
  if (d <= 32) {
    // single-word row (the common case): the bitmap stays in a register -- no shared-memory store, no group
    // sync; the caller materialises it only if it has to replay
    for (uint32_t c0 = 0; c0 < d; c0 += G) {
      const uint32_t k = c0 + T.tl;
      const bool valid = k < d;
      const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
      const uint32_t pos = lower_bound_u32(prow, pdeg, x, lgp);
      const bool found = pos < pdeg && __ldg(prow + pos) == x;
      const bool isprev = valid && (x == prev);
      const uint32_t bprev = T.ballot(isprev);
      if (bprev) kp = c0 + __ffs(bprev) - 1;
      const uint32_t bal = T.ballot(valid && found && !isprev);
      word0 |= (G == 32) ? bal : (bal << c0);
      m += __popc(bal);
    }
    in_regs = true;
    return m;
  }
  if (d <= 2 * 32 || fwd_cost <= rev_cost) {
    for (uint32_t c0 = 0; c0 < d; c0 += G) {
      const uint32_t k = c0 + T.tl;
      const bool valid = k < d;
      const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
      const uint32_t pos = lower_bound_u32(prow, pdeg, x, lgp);
      const bool found = pos < pdeg && __ldg(prow + pos) == x;
      const bool isprev = valid && (x == prev);
      const uint32_t bprev = T.ballot(isprev);
      if (bprev) kp = c0 + __ffs(bprev) - 1;
      const uint32_t bal = T.ballot(valid && found && !isprev);
      if (T.tl == 0) {
        if (G == 32 || (c0 & 31) == 0) bm[c0 >> 5] = bal; else bm[c0 >> 5] |= bal << (c0 & 31);
      }
      m += __popc(bal);
    }
  } else {
    for (uint32_t w = T.tl; w < nwords; w += G) bm[w] = 0u;
    T.sync();
    const uint32_t nkeys = pdeg + 1;
    uint32_t mloc = 0, kploc = B2W_NONE;
    for (uint32_t c0 = 0; c0 < nkeys; c0 += G) {
      const uint32_t ii = c0 + T.tl;
      const bool valid = ii < nkeys;
      const uint32_t y = valid ? (ii < pdeg ? __ldg(prow + ii) : prev) : B2W_NONE;
      const uint32_t pos = lower_bound_u32(crow, d, y, lgd);
      if (valid && pos < d && __ldg(crow + pos) == y) {
        if (ii == pdeg) kploc = pos;
        else if (y != prev) { atomicOr(&bm[pos >> 5], 1u << (pos & 31)); ++mloc; }
      }
    }
    m = T.sum(mloc);
    kp = T.minu(kploc);
  }
  T.sync();
  return m;
}
