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

// Number of elements of the sorted, duplicate-free row[0..n) (n >= 1) that are < x, and whether x is one
// of them.  `k` = floor(log2 n) must be uniform across the group.  k + 1 dependent loads, none of them out
// of bounds, no data-dependent branch and (fully unrolled through a jump on k) five instructions per probe:
// address, load, compare, predicated add, predicated move.
//   * first probe at row[2^k - 1] picks one of two overlapping windows of 2^k possible answers:
//     [0, 2^k - 1] or [n - 2^k + 1, n]  (the second contains the impossible answers below 2^k, harmless);
//   * invariant: answer in [lo, lo + 2S - 1] with lo + 2S - 1 <= n; probe row[lo + S - 1] (always < n);
//   * x is in the row iff the LAST probe that loaded a value >= x loaded x itself: telling answer a from a + 1
//     needs exactly the probe of row[a] (the top of the first window, row[2^k - 1], is the first probe), probes
//     with v >= x move down the row, and below row[a] every value is < x.
// (The first version of this search, `if (idx <= n) v = row[idx - 1]` inside a counted loop, compiled to a
// real branch with a reconvergence barrier: 13 instructions per probe, a third of the whole kernel --
// profiles/r1_uw_g32_powerlaw_nw1_src.txt.)
#define B2W_LB_PROBE(S)                              \
  {                                                  \
    const uint32_t v_ = __ldg(row + lo + ((S) - 1)); \
    if (v_ < x) lo += (S); else ge = v_;             \
  }
// WARP_UNIFORM (k is the same for all 32 lanes, i.e. 32-lane groups): jump into the unrolled probe sequence.
// Sub-warp groups have different k per group; jumping to different entry points would serialise the groups
// of a warp, so they run a counted loop instead (lanes that finish early wait at the loop exit).  The unrolled
// form costs ~330 instructions of code per inlined copy; the generic weight-streaming kernel (already 2760
// instructions) went from 0.63 to 0.48 G steps/s with it -- instruction-cache misses became its second largest
// stall -- and uses the loop form too.
template <bool WARP_UNIFORM>
__device__ __forceinline__ uint32_t lower_bound_eq(const uint32_t* __restrict__ row, const uint32_t n, const uint32_t x,
                                                   const uint32_t k, bool& found) {
  const uint32_t top = 1u << k;
  const uint32_t v0 = __ldg(row + (top - 1));
  uint32_t ge = (v0 < x) ? B2W_NONE : v0;                            // value of the last probe that was >= x
  uint32_t lo = (v0 < x) ? n - top + 1 : 0u;
  if (WARP_UNIFORM) {
    uint32_t kk = k;
    for (; kk > 16; --kk) B2W_LB_PROBE(1u << (kk - 1))                // rows beyond 2^17 entries
    switch (kk) {
      case 16: B2W_LB_PROBE(32768u)
      case 15: B2W_LB_PROBE(16384u)
      case 14: B2W_LB_PROBE(8192u)
      case 13: B2W_LB_PROBE(4096u)
      case 12: B2W_LB_PROBE(2048u)
      case 11: B2W_LB_PROBE(1024u)
      case 10: B2W_LB_PROBE(512u)
      case 9: B2W_LB_PROBE(256u)
      case 8: B2W_LB_PROBE(128u)
      case 7: B2W_LB_PROBE(64u)
      case 6: B2W_LB_PROBE(32u)
      case 5: B2W_LB_PROBE(16u)
      case 4: B2W_LB_PROBE(8u)
      case 3: B2W_LB_PROBE(4u)
      case 2: B2W_LB_PROBE(2u)
      case 1: B2W_LB_PROBE(1u)
      default: break;
    }
  } else {
    for (uint32_t S = top >> 1; S; S >>= 1) {
      const uint32_t v_ = __ldg(row + (lo + S - 1u));                 // 32-bit index first: one IMAD.WIDE
      if (v_ < x) lo += S; else ge = v_;
    }
  }
  found = ge == x;                                                   // (x == B2W_NONE: callers mask invalid lanes)
  return lo;
}
#undef B2W_LB_PROBE


// N independent searches in the same row, interleaved probe by probe: the dependent-load chain of a search
// (k + 1 loads, each an L1/L2 round trip on a hub row) is what a multi-chunk membership step waits for, and N
// chunks in flight divide the number of such chains per step by N.
template <int N>
__device__ __forceinline__ void lower_bound_eq_xN(const uint32_t* __restrict__ row, const uint32_t n, const uint32_t (&x)[N],
                                                  const uint32_t k, uint32_t (&lo)[N], bool (&found)[N]) {
  const uint32_t top = 1u << k;
  const uint32_t v = __ldg(row + (top - 1));
  uint32_t ge[N];
#pragma unroll
  for (int t = 0; t < N; ++t) {
    ge[t] = (v < x[t]) ? B2W_NONE : v;
    lo[t] = (v < x[t]) ? n - top + 1 : 0u;
  }
  for (uint32_t S = top >> 1; S; S >>= 1) {
    uint32_t a[N];
#pragma unroll
    for (int t = 0; t < N; ++t) a[t] = __ldg(row + (lo[t] + S - 1u));
#pragma unroll
    for (int t = 0; t < N; ++t) { if (a[t] < x[t]) lo[t] += S; else ge[t] = a[t]; }
  }
#pragma unroll
  for (int t = 0; t < N; ++t) found[t] = ge[t] == x[t];
}
// Chunks of G keys searched per iteration.  Measured (G steps/s): BASELINE config #3, 32-lane groups: 1 chunk 1.92,
// 2 chunks 2.06 / 2.04, 3 chunks 2.03, 4 chunks 1.90;  config #2, 8-lane groups: 1 chunk 4.35, 2 chunks 4.62-4.70,
// 3 chunks 5.21, 4 chunks 5.38 (a whole in-register word, 32 positions, in one iteration).
#ifndef B2W_CHUNKS_WARP
#define B2W_CHUNKS_WARP 2      /* 32-lane groups */
#endif
#ifndef B2W_INREG_MEMBERSHIP
#define B2W_INREG_MEMBERSHIP 1 /* 32-lane groups, both rows <= 32 entries: search row(prev) in registers with shuffles.
                                  Round 2 A/B on the B200: parity-green (235 tests), 2.04 -> 2.06 G steps/s on config #3 */
#endif
#ifndef B2W_SUBWARP_REDUX
#define B2W_SUBWARP_REDUX 1   /* sub-warp groups, rows <= 32: assemble the bitmap word with one REDUX.OR.  Round 2 A/B on the
                                  B200: parity-green, 5.38 -> 6.11 G steps/s on config #2 (index-free kernel) */
#endif
#ifndef B2W_CHUNKS_SUBWARP
#define B2W_CHUNKS_SUBWARP 0   /* sub-warp groups: 0 = 32 / G (one bitmap word per iteration) */
#endif


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
template <int G, bool UNROLLED_SEARCH = false>
__device__ __forceinline__ uint32_t membership_bitmap(const Tile<G>& T, const uint32_t* __restrict__ crow,
                                                      const uint32_t d, const uint32_t* __restrict__ prow,
                                                      const uint32_t pdeg, const uint32_t prev,
                                                      uint32_t* __restrict__ bm, uint32_t& kp, uint32_t& word0, bool& in_regs) {
  const uint32_t nwords = (d + 31) >> 5;
  const uint32_t lgp = 32 - __clz(pdeg), lgd = 32 - __clz(d);          // probes per search (cost model)
  const uint32_t kp2 = 31 - __clz(pdeg), kd2 = 31 - __clz(d);          // floor(log2): pdeg >= 1, d >= 1
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
    uint32_t c0 = 0;
#if B2W_INREG_MEMBERSHIP
    if (G == 32 && pdeg <= 32) {
      // both rows fit one register per lane: row(prev) is loaded once (coalesced) and searched with shuffles -- one
      // memory round trip instead of a chain of k + 1 dependent loads (31 % of the steps of config #3,
      // profiles/r1_costmodel_config3.txt).  Padding with B2W_NONE keeps the 32 values sorted.
      const uint32_t pv = (uint32_t)T.tl < pdeg ? __ldg(prow + T.tl) : B2W_NONE;
      const uint32_t k = T.tl;
      const bool valid = k < d;
      const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
      uint32_t base = 0;
#pragma unroll
      for (uint32_t half = 16; half; half >>= 1) {                    // lower_bound over 32 register-resident values
        const uint32_t v = __shfl_sync(B2W_FULL, pv, (int)(base + half - 1));
        if (v < x) base += half;
      }
      const bool found = __shfl_sync(B2W_FULL, pv, (int)base) == x;   // (idle lanes match the padding: masked below)
      const bool isprev = valid && (x == prev);
      kp = __reduce_min_sync(T.mask, isprev ? k : B2W_NONE);
      word0 = T.ballot(valid && found && !isprev);
      in_regs = true;
      return __popc(word0);
    }
#endif
#if B2W_SUBWARP_REDUX
    if (G < 32) {
      // the whole word in one go: every lane searches its 32 / G positions, builds its own bits of the word and one
      // REDUX.OR assembles it (instead of one ballot + shift + popcount per chunk)
      constexpr int NW = 32 / G;
      uint32_t x[NW], p[NW], kk[NW];
      bool f[NW];
#pragma unroll
      for (int t = 0; t < NW; ++t) {
        kk[t] = t * G + T.tl;
        x[t] = kk[t] < d ? __ldg(crow + kk[t]) : B2W_NONE;
      }
      lower_bound_eq_xN<NW>(prow, pdeg, x, kp2, p, f);
      uint32_t mine = 0u, kpl = B2W_NONE;
#pragma unroll
      for (int t = 0; t < NW; ++t) {
        const bool valid = kk[t] < d;
        const bool isprev = valid && (x[t] == prev);
        if (isprev) kpl = kk[t];
        if (valid && f[t] && !isprev) mine |= 1u << kk[t];
      }
      word0 = __reduce_or_sync(T.mask, mine);
      kp = __reduce_min_sync(T.mask, kpl);
      in_regs = true;
      return __popc(word0);
    }
#endif
    if (G < 32) {
      constexpr int N = B2W_CHUNKS_SUBWARP > 1 ? B2W_CHUNKS_SUBWARP : (32 / G > 1 ? 32 / G : 2);
      for (; c0 + G < d; c0 += N * G) {                               // sub-warp groups: N chunks per iteration
        uint32_t x[N], p[N], kk[N];
        bool f[N], valid[N];
#pragma unroll
        for (int t = 0; t < N; ++t) {
          kk[t] = c0 + t * G + T.tl;
          valid[t] = kk[t] < d;
          x[t] = valid[t] ? __ldg(crow + kk[t]) : B2W_NONE;
        }
        lower_bound_eq_xN<N>(prow, pdeg, x, kp2, p, f);
        uint32_t kpl = B2W_NONE;
#pragma unroll
        for (int t = 0; t < N; ++t) {
          const bool isprev = valid[t] && (x[t] == prev);
          if (isprev) kpl = kk[t];
          const uint32_t bal = T.ballot(valid[t] && f[t] && !isprev);
          if (c0 + t * G < 32) word0 |= bal << (c0 + t * G);
          m += __popc(bal);
        }
        kp = min(kp, __reduce_min_sync(T.mask, kpl));
      }
    }
    for (; c0 < d; c0 += G) {
      const uint32_t k = c0 + T.tl;
      const bool valid = k < d;
      const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
      bool found;
      lower_bound_eq<UNROLLED_SEARCH>(prow, pdeg, x, kp2, found);
      const bool isprev = valid && (x == prev);
      kp = min(kp, __reduce_min_sync(T.mask, isprev ? k : B2W_NONE));   // one REDUX instead of ballot + ffs + select
      const uint32_t bal = T.ballot(valid && found && !isprev);
      word0 |= (G == 32) ? bal : (bal << c0);
      m += __popc(bal);
    }
    in_regs = true;
    return m;
  }
  constexpr bool MULTI = (B2W_CHUNKS_WARP > 1) && G == 32;
  constexpr int N = B2W_CHUNKS_WARP > 1 ? B2W_CHUNKS_WARP : 2;
  if (d <= 2 * 32 || fwd_cost <= rev_cost) {
    uint32_t c0 = 0;
    if (MULTI) {
      for (; c0 + G < d; c0 += N * G) {                               // N chunks of the row per iteration
        uint32_t x[N], p[N], kk[N];
        bool f[N], valid[N];
#pragma unroll
        for (int t = 0; t < N; ++t) {
          kk[t] = c0 + t * G + T.tl;
          valid[t] = kk[t] < d;
          x[t] = valid[t] ? __ldg(crow + kk[t]) : B2W_NONE;
        }
        lower_bound_eq_xN<N>(prow, pdeg, x, kp2, p, f);
        uint32_t kpl = B2W_NONE;
#pragma unroll
        for (int t = 0; t < N; ++t) {
          const bool isprev = valid[t] && (x[t] == prev);
          if (isprev) kpl = kk[t];
          const uint32_t bal = T.ballot(valid[t] && f[t] && !isprev);
          if (T.tl == 0 && c0 + t * G < d) bm[(c0 >> 5) + t] = bal;
          m += __popc(bal);
        }
        kp = min(kp, __reduce_min_sync(T.mask, kpl));
      }
    }
    for (; c0 < d; c0 += G) {
      const uint32_t k = c0 + T.tl;
      const bool valid = k < d;
      const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
      bool found;
      lower_bound_eq<UNROLLED_SEARCH>(prow, pdeg, x, kp2, found);
      const bool isprev = valid && (x == prev);
      kp = min(kp, __reduce_min_sync(T.mask, isprev ? k : B2W_NONE));   // one REDUX instead of ballot + ffs + select
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
    uint32_t c0 = 0;
    if (MULTI) {
      for (; c0 + G < nkeys; c0 += N * G) {                           // N chunks of keys per iteration
        uint32_t y[N], p[N], ii[N];
        bool f[N];
#pragma unroll
        for (int t = 0; t < N; ++t) {
          ii[t] = c0 + t * G + T.tl;
          y[t] = ii[t] < nkeys ? (ii[t] < pdeg ? __ldg(prow + ii[t]) : prev) : B2W_NONE;
        }
        lower_bound_eq_xN<N>(crow, d, y, kd2, p, f);
#pragma unroll
        for (int t = 0; t < N; ++t) {
          if (ii[t] < nkeys && f[t]) {
            if (ii[t] == pdeg) kploc = p[t];
            else if (y[t] != prev) { atomicOr(&bm[p[t] >> 5], 1u << (p[t] & 31)); ++mloc; }
          }
        }
      }
    }
    for (; c0 < nkeys; c0 += G) {
      const uint32_t ii = c0 + T.tl;
      const bool valid = ii < nkeys;
      const uint32_t y = valid ? (ii < pdeg ? __ldg(prow + ii) : prev) : B2W_NONE;
      bool found;
      const uint32_t pos = lower_bound_eq<UNROLLED_SEARCH>(crow, d, y, kd2, found);
      if (valid && found) {
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
