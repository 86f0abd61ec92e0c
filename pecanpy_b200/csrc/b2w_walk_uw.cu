// b2w_walk_uw.cu -- SparseOTF node2vec on UNWEIGHTED graphs: the membership-bitmap kernel.
//
// PecanPy's default input is an unweighted edge list (`--weighted` off => data[:] = 1.0,
// graph.py:479-480, cli.py), and BASELINE configs #1-#3 are unweighted.  Then every biased weight
// of a step is one of three constants
//     w_in = 1.0f,   w_out = f32(1/q),   w_ret = f32(1/p)            (rw/sparse_rw.py:86-87)
// and the whole transition distribution of (cur, prev) is determined by
//     deg(cur),  the POSITIONS in row(cur) of the common neighbours N(cur) & N(prev),  pos(prev).
// So a step does not have to stream row(cur) at all:
//   phase 1  membership -> bitmap over the positions of row(cur), searching whichever side is
//            cheaper: every neighbour of cur in row(prev) (lane-parallel lower_bound, coalesced
//            stream of row(cur)), or every neighbour of prev in row(cur) (touches only
//            O(deg(prev) log deg(cur)) words of the hub row -- this is what makes power-law hubs
//            cheap).  T groups of G lanes (G = 8/16/32) each own one walker, so a warp keeps
//            32/G independent gather chains in flight.
//   phase 2  normaliser: the host verified that 1, w_out, w_ret are multiples of one power of two g and
//            (max_degree + 1) * max(w) < 2^24 g, so NO f32 partial sum can round and the reference's sequential
//            f32 sum S equals the exact count-weighted total.  In units of g everything is a small integer:
//            W_k = exact un-normalised prefix up to position k, W_d the total (S = W_d g).
//   phase 3  filter: the reference's cdf_k (sequential f32 cumsum of fdiv(w_i, S), numba/np/arraymath.py:384-405)
//            equals (W_k / W_d)(1 + t) with |t| <= e_k = 1.02 (k + 3) 2^-24.  Popcount prefix over the bitmap words
//            locates the first word, then the first position k, with W_k >= u W_d (1 - e); if also
//            W_k >= u W_d (1 + e + 2 e^2) the choice is proven.  Otherwise (~1 % of the steps on a power-law
//            graph) the group forms the three f32 quotients and replays the f32 recurrence exactly from the
//            bitmap (no memory traffic).
// Bit-exact with the generic kernels and the oracle for every (cur, prev, u); dispatched only when
// the exactness precondition holds (power-of-two-like p, q), else the generic stream kernel runs.
//
// Reference: pecanpy.py:164-210 (_random_walks), :522-561 (SparseOTF.get_move_forward),
//            rw/sparse_rw.py:51-91 (get_normalized_probs), :142-230 (isnotin).
#include <cmath>
#include <cstring>

#include "b2w_membership.cuh"
#include "b2w_replay.cuh"

namespace {

constexpr int UW_THREADS = 256;
constexpr int UW_BW = 32;            // bitmap words of shared memory per group (rows up to 1024)
constexpr uint32_t NONE = B2W_NONE;

struct UwConsts {
  float w_out, w_ret;                // f32(1/q), f32(1/p)
  float g;                           // common grid of the three weights: a power of two with 1, w_out, w_ret in g Z
  uint32_t a_in, a_out, a_ret;       // 1 / g, w_out / g, w_ret / g  (exact integers, (max_degree + 1) max(a) < 2^24)
  uint32_t gbm_stride;               // words of global bitmap scratch per group (0: never needed)
  uint32_t* gbm;                     // global bitmap scratch
};

template <int G>
__device__ __noinline__ uint32_t replay_exact(const uint32_t* __restrict__ bm, const bool has_bm,
                                              const uint32_t nwords, const uint32_t d, const uint32_t kp, const float fa,
                                              const float fo, const float fp, const double u) {
  const Tile<G> T;
  const float ub = upper_float(u);                                    // cdf < u  <=>  cdf < ub
  float cdf = 0.f;
  uint32_t k = 0, choice = d;
  for (uint32_t w0 = 0; w0 < nwords; w0 += G) {
    const uint32_t w = w0 + T.tl;
    uint32_t bits = (has_bm && w < nwords) ? bm[w] : 0u;
    if (kp != NONE && (kp >> 5) == w) bits |= 1u << (kp & 31);
    uint32_t nz = T.ballot(bits != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      uint32_t wb = T.shfl(bits, src);
      const uint32_t wbase = (w0 + src) << 5;
      while (wb) {
        const uint32_t pos = wbase + __ffs(wb) - 1;
        wb &= wb - 1;
        if (advance_run(cdf, k, pos - k, fo, ub, choice)) return choice;
        cdf = __fadd_rn(cdf, pos == kp ? fp : fa);                    // the special element at `pos`
        if (cdf >= ub) return pos;
        k = pos + 1;
      }
    }
  }
  if (advance_run(cdf, k, d - k, fo, ub, choice)) return choice;
  return d;                                                           // cdf[-1] < u: the reference's overflow
}

// One SparseOTF step of one walker by a group of G lanes: membership bitmap, exact normaliser, filter,
// exact replay.  All arguments are group-uniform; `bm` is a group-private bitmap of >= ceil(d / 32) words.
// Returns the reference's `choice` (== d for its cdf[-1] < u overflow).  Ends with a group sync.
template <int G>
__device__ __forceinline__ uint32_t uw_step(const Tile<G>& T, const uint32_t flags, const uint32_t a_in,
                                            const uint32_t a_out, const uint32_t a_ret, const float grid,
                                            const uint32_t* __restrict__ crow, const uint32_t d,
                                            const uint32_t* __restrict__ prow, const uint32_t pdeg, const uint32_t prev,
                                            const bool has_prev, const double u, uint32_t* __restrict__ bm,
                                            uint32_t& st_replays, uint32_t& st_overflow) {
  const uint32_t nwords = (d + 31) >> 5;
  // ---------------- phase 1: membership bitmap over the positions of row(cur)
  uint32_t m = 0, kp = NONE, word0 = 0;
  bool in_regs = false;
  if (has_prev) m = membership_bitmap<G, G == 32>(T, crow, d, prow, pdeg, prev, bm, kp, word0, in_regs);

  // ---------------- phase 2: exact normaliser, in integer units of the common weight grid g
  // W_k = (k+1) a_o + n_in(k) (a_in - a_o) + [kp <= k] (a_ret - a_o): the exact un-normalised prefix (< 2^24, host
  // verified), W_d = W_{d-1} the exact total; S = W_d g is the reference's sequential f32 sum (no partial sum rounds).
  // Differences are taken modulo 2^32; every true value is a small non-negative integer.
  const uint32_t a_o = has_prev ? a_out : a_in;                    // first step: every weight is 1
  const uint32_t da = a_in - a_o, dr = a_ret - a_o;
  const uint32_t h = (kp != NONE) ? 1u : 0u;
  const uint32_t Wd = d * a_o + m * da + h * dr;

  // ---------------- phase 3: filter.  The reference's cdf_k = f32 sequential cumsum of fdiv(w_i, S) satisfies
  //   cdf_k = (W_k / W_d)(1 + t),  |t| <= e_k := 1.02 (k + 3) 2^-24      ((k+3) 2^-24 <= 0.01; one fdiv + k additions)
  // so with A = u W_d:   W_k < A (1 - e)          => cdf_k < u   (1 - e <= 1 / (1 + e))
  //                      W_k >= A (1 + e + 2 e^2) => cdf_k >= u  (1 / (1 - e) <= 1 + e + 2 e^2)
  // for any e >= e_k.  W_k is an exact integer (exact in f64 too); the two thresholds are uniform per row (word
  // level) / per word (position level) and are pushed outwards by 2^-45 relative, far more than the f64 rounding
  // of the three products, so a lane only converts its W_k and compares.  The first k that is "possible" is the
  // reference's choice if it is also "sure"; otherwise (probability ~ e W_d per step) the exact replay decides.
  const double EC = 1.02 * 5.9604644775390625e-08;                    // 1.02 * 2^-24
  const double A = u * (double)Wd;
  uint32_t choice = d;                                            // default: cdf[-1] < u (overflow)
  bool replay = (flags & B2W_FLAG_FORCE_EXACT_REPLAY) != 0 || d > 160000u;
  if (!replay) {
    uint32_t wsel = 0, bits_sel = 0, nc_before = 0;
    if (nwords > 1) {
      // word level: first word whose last position possibly reaches u (row-uniform, most conservative threshold)
      const double e_row = EC * (double)(d + 2);
      const double t_row = A * (1.0 - e_row - 2.9e-14);
      uint32_t carry = 0;
      wsel = NONE;
      for (uint32_t w0 = 0; w0 < nwords; w0 += G) {
        const uint32_t w = w0 + T.tl;
        const bool valid = w < nwords;
        const uint32_t bits = (valid && has_prev) ? bm[w] : 0u;
        const uint32_t cnt = __popc(bits);
        const uint32_t incl = T.incl_scan(cnt) + carry;
        const uint32_t kend = min(d, (w + 1) << 5) - 1;
        const uint32_t Wk = (kend + 1) * a_o + incl * da + (kp <= kend ? dr : 0u);   // NONE compares false
        const uint32_t bal = T.ballot(valid && (double)Wk >= t_row);
        if (bal) {
          const int src = __ffs(bal) - 1;
          wsel = w0 + src;
          bits_sel = T.shfl(bits, src);
          nc_before = T.shfl(incl - cnt, src);
          break;
        }
        carry = T.shfl(incl, G - 1);
      }
    } else if (has_prev) {
      bits_sel = in_regs ? word0 : bm[0];
    }
    if (wsel != NONE) {
      // position level inside the selected word
      const uint32_t nb = min(32u, d - (wsel << 5));
      const double e_w = EC * (double)((wsel << 5) + nb + 2);
      const double t_poss = A * (1.0 - e_w - 2.9e-14);
      const double t_sure = A * (1.0 + e_w + 2.0 * e_w * e_w + 2.9e-14);
      bool decided = false;
      for (uint32_t r0 = 0; r0 < nb && !decided; r0 += G) {
        const uint32_t b = r0 + T.tl;
        const uint32_t k = (wsel << 5) + b;
        const bool valid = b < nb;
        const uint32_t nc = nc_before + __popc(bits_sel & (0xFFFFFFFFu >> (31 - (b & 31))));
        const uint32_t Wk = (k + 1) * a_o + nc * da + (kp <= k ? dr : 0u);
        const double Wkd = (double)Wk;
        const uint32_t bp = T.ballot(valid && Wkd >= t_poss);
        if (bp) {
          const int f = __ffs(bp) - 1;
          const bool sure = T.shfl((Wkd >= t_sure) ? 1 : 0, f) != 0;
          if (sure) choice = (wsel << 5) + r0 + f; else replay = true;
          decided = true;
        }
      }
      if (!decided) replay = true;                               // rounding at the word boundary
    } else {
      replay = true;                                               // no word reaches the row threshold (u ~ 1)
    }
  }
  if (replay) {
    if (in_regs) { if (T.tl == 0) bm[0] = word0; T.sync(); }      // the replay reads the bitmap from memory
    // the reference's probabilities: three exact f32 quotients by S = W_d g (rw/sparse_rw.py:89)
    const float S = __fmul_rn(__uint2float_rn(Wd), grid);          // exact: W_d < 2^24, g a power of two
    const float fa = __fdiv_rn(1.0f, S);                           // w_out = a_out g, w_ret = a_ret g, exactly
    const float fo = has_prev ? __fdiv_rn(__fmul_rn(__uint2float_rn(a_out), grid), S) : fa;
    const float fp = __fdiv_rn(__fmul_rn(__uint2float_rn(a_ret), grid), S);
    choice = replay_exact<G>(bm, has_prev, nwords, d, kp, fa, fo, fp, u);
    ++st_replays;
  }
  if (choice == d) ++st_overflow;
  if (!in_regs || replay) T.sync();                               // a bitmap in memory is rewritten by the next step

  return choice;
}

// (A restructured main loop -- one outer iteration per block of G output entries, Philox and the block store hoisted
// out of the step loop, hub rows through an out-of-line copy of the step so that the common path only sees a
// shared-memory bitmap -- executes ~40 fewer instructions per step but needs more live registers than the 48 that
// 5 CTAs/SM allow: measured 1.59 (out-of-line hub step, 76 B of spills) and 1.79 (inline) vs 1.91 G steps/s for the
// flat loop below on BASELINE config #3.  Occupancy beats instruction count here: the kernel waits on dependent loads.)
template <int G, int MINB>
__global__ void __launch_bounds__(UW_THREADS, MINB) walk_uw_kernel(const WalkParams P, const UwConsts C) {
  constexpr int GROUPS = UW_THREADS / G;
  __shared__ uint32_t s_bm[GROUPS][UW_BW];
  const Tile<G> T;
  const int gib = threadIdx.x / G;                                   // group in block
  const uint32_t L = P.L;
  uint32_t st_steps = 0, st_replays = 0, st_overflow = 0;

  for (;;) {
    unsigned long long i = 0;
    if (T.tl == 0) i = atomicAdd(P.counter, 1ull);
    i = T.shfl(i, 0);
    if (i >= P.n_rows) break;
    uint32_t* const out = P.out + i * P.ld_out;
    uint32_t cur = __ldg(P.start + i);
    uint32_t prev = 0, pdeg = 0;
    const uint32_t* prow = P.indices;
    uint32_t cs = __ldg(P.indptr + cur);
    uint32_t ce = __ldg(P.indptr + cur + 1);
    uint32_t eff = L + 1;
    uint32_t myval = (T.tl == 0) ? cur : 0u;                          // lane (e mod G) holds output entry e
    double my_u = 0.0;
    uint32_t j = 1;
    for (; j <= L; ++j) {
      const uint32_t d = ce - cs;
      if (d == 0) { eff = j; break; }                                 // pecanpy.py:194-196, 204-206
      if (((j - 1) & (G - 1)) == 0) {
        const uint32_t sj = j + T.tl;
        if (sj <= L) my_u = step_uniform(P, i, sj);
      }
      const double u = T.shfl(my_u, (j - 1) & (G - 1));
      const bool has_prev = j > 1;
      const uint32_t nwords = (d + 31) >> 5;
      uint32_t* bm = s_bm[gib];
      if (nwords > UW_BW)                                             // hub row: this group's global scratch (rare)
        bm = C.gbm + (size_t)(blockIdx.x * GROUPS + gib) * C.gbm_stride;
      const uint32_t* const crow = P.indices + cs;

      const uint32_t choice = uw_step<G>(T, P.flags, C.a_in, C.a_out, C.a_ret, C.g, crow, d, prow, pdeg, prev, has_prev, u, bm, st_replays, st_overflow);

      const uint32_t nxt = __ldg(crow + choice);                      // unchecked, as pecanpy.py:559
      if (T.tl == (j & (G - 1))) myval = nxt;
      if ((j & (G - 1)) == G - 1) {
        __stcs(out + ((j & ~(uint32_t)(G - 1)) + T.tl), myval);   // streaming store: keep the graph in L2
        myval = 0u;
      }
      prev = cur; prow = crow; pdeg = d;
      cur = nxt;
      cs = __ldg(P.indptr + cur);
      ce = __ldg(P.indptr + cur + 1);
      ++st_steps;
    }
    // tail: the G-block holding entry j (first entry not produced), then zeros, then eff at L+1
    const uint32_t blk = j & ~(uint32_t)(G - 1);
    for (uint32_t b0 = blk; b0 < L + 2; b0 += G) {
      const uint32_t e = b0 + T.tl;
      uint32_t v = (b0 == blk) ? myval : 0u;
      if (e == L + 1) v = eff;
      if (e < L + 2) __stcs(out + e, v);
    }
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS) && T.tl == 0) {
    if (st_steps) atomicAdd((unsigned long long*)&P.stats->steps, (unsigned long long)st_steps);
    if (st_replays) atomicAdd((unsigned long long*)&P.stats->exact_replays, (unsigned long long)st_replays);
    if (st_overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)st_overflow);
  }
}

// ---------------------------------------------------------------------------------------------------
// Cooperative variant for G < 32.  Small groups keep 32/G walkers per warp in flight and share every
// instruction of the (dominant) short steps, but a hub row handled by G lanes alone would stall the other
// groups of the warp.  Here the warp runs an explicit per-group state machine in lockstep; each iteration
//   1. groups without a walker fetch the next row (or retire),
//   2. steps that touch a long row (deg(cur) or deg(prev) > BIG) are executed ONE AT A TIME BY ALL 32 LANES
//      (the owner's state is broadcast with full-warp shuffles, uw_step<32> runs, the owner keeps the result),
//   3. all remaining (short) steps run concurrently, one per group, with uw_step<G>.
// Results are identical to every other kernel (same Philox counters, same exact arithmetic).
template <int G, int MINB>
__global__ void __launch_bounds__(UW_THREADS, MINB) walk_uw_coop_kernel(const WalkParams P, const UwConsts C,
                                                                        const uint32_t BIG) {
  static_assert(G < 32, "the cooperative kernel is for sub-warp groups");
  constexpr int GROUPS = UW_THREADS / G;
  constexpr int WARPS = UW_THREADS / 32;
  constexpr int SMALL_WORDS = 4;                                      // short steps: deg(cur) <= BIG <= 128
  __shared__ uint32_t s_bm[GROUPS][SMALL_WORDS];
  __shared__ uint32_t s_wbm[WARPS][UW_BW];
  const Tile<G> T;
  const Tile<32> TW;
  const int gib = threadIdx.x / G;
  const int wib = threadIdx.x >> 5;
  const uint32_t wgid = blockIdx.x * WARPS + wib;                     // global warp id
  uint32_t* const gbm = C.gbm + (size_t)wgid * C.gbm_stride;          // long-row bitmap scratch of this warp
  const uint32_t L = P.L;
  const uint32_t gmask = T.mask;
  uint32_t st_steps = 0, st_replays = 0, st_overflow = 0;

  bool have = false, done = false;
  unsigned long long i = 0;
  uint32_t* out = nullptr;
  uint32_t cur = 0, prev = 0, pdeg = 0, cs = 0, ce = 0, eff = 0, myval = 0, j = 1;
  const uint32_t* prow = P.indices;
  double my_u = 0.0;

  auto finish = [&]() {
    // tail: the G-block holding entry j (first entry not produced), then zeros, then eff at L+1
    const uint32_t blk = j & ~(uint32_t)(G - 1);
    for (uint32_t b0 = blk; b0 < L + 2; b0 += G) {
      const uint32_t e = b0 + T.tl;
      uint32_t v = (b0 == blk) ? myval : 0u;
      if (e == L + 1) v = eff;
      if (e < L + 2) __stcs(out + e, v);
    }
    have = false;
  };

  for (;;) {
    if (!have && !done) {
      unsigned long long r = 0;
      if (T.tl == 0) r = atomicAdd(P.counter, 1ull);
      r = T.shfl(r, 0);
      if (r >= P.n_rows) {
        done = true;
      } else {
        i = r;
        out = P.out + i * P.ld_out;
        cur = __ldg(P.start + i);
        prev = 0; pdeg = 0; prow = P.indices;
        cs = __ldg(P.indptr + cur);
        ce = __ldg(P.indptr + cur + 1);
        eff = L + 1;
        myval = (T.tl == 0) ? cur : 0u;
        j = 1;
        have = true;
      }
    }
    if (__all_sync(B2W_FULL, done)) break;
    uint32_t d = have ? ce - cs : 0u;
    if (have && d == 0) { eff = j; finish(); }                         // dead end (pecanpy.py:194-196, 204-206)
    const bool active = have;
    if (active && ((j - 1) & (G - 1)) == 0) {
      const uint32_t sj = j + T.tl;
      if (sj <= L) my_u = step_uniform(P, i, sj);
    }
    const double u = T.shfl(my_u, (j - 1) & (G - 1));
    const bool has_prev = j > 1;
    const uint32_t* const crow = P.indices + cs;
    const bool big = active && (d > BIG || (has_prev && pdeg > BIG));
    uint32_t bigm = __ballot_sync(B2W_FULL, big);
    uint32_t choice = 0, rep = 0, ovf = 0;
    while (bigm) {                                                    // warp-uniform loop
      const int src = __ffs(bigm) - 1;                                // a lane of the owning group
      const uint32_t b_cs = __shfl_sync(B2W_FULL, cs, src);
      const uint32_t b_d = __shfl_sync(B2W_FULL, d, src);
      const unsigned long long b_prow = __shfl_sync(B2W_FULL, (unsigned long long)(uintptr_t)prow, src);
      const uint32_t b_pdeg = __shfl_sync(B2W_FULL, pdeg, src);
      const uint32_t b_prev = __shfl_sync(B2W_FULL, prev, src);
      const bool b_hp = __shfl_sync(B2W_FULL, j, src) > 1;
      const double b_u = __shfl_sync(B2W_FULL, u, src);
      uint32_t* const bmw = (((b_d + 31) >> 5) <= (uint32_t)UW_BW) ? s_wbm[wib] : gbm;
      uint32_t r2 = 0, o2 = 0;
      const uint32_t c = uw_step<32>(TW, P.flags, C.a_in, C.a_out, C.a_ret, C.g, P.indices + b_cs, b_d, reinterpret_cast<const uint32_t*>((uintptr_t)b_prow),
                                     b_pdeg, b_prev, b_hp, b_u, bmw, r2, o2);
      const uint32_t om = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << (src & ~(G - 1)));
      if (gmask == om) { choice = c; rep = r2; ovf = o2; }
      bigm &= ~om;
    }
    if (active && !big) choice = uw_step<G>(T, P.flags, C.a_in, C.a_out, C.a_ret, C.g, crow, d, prow, pdeg, prev, has_prev, u, s_bm[gib], rep, ovf);
    if (active) {
      st_replays += rep; st_overflow += ovf;
      const uint32_t nxt = __ldg(crow + choice);                      // unchecked, as pecanpy.py:559
      if (T.tl == (j & (G - 1))) myval = nxt;
      if ((j & (G - 1)) == G - 1) {
        __stcs(out + ((j & ~(uint32_t)(G - 1)) + T.tl), myval);   // streaming store: keep the graph in L2
        myval = 0u;
      }
      prev = cur; prow = crow; pdeg = d;
      cur = nxt;
      cs = __ldg(P.indptr + cur);
      ce = __ldg(P.indptr + cur + 1);
      ++st_steps;
      ++j;
      if (j > L) finish();
    }
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS) && T.tl == 0) {
    if (st_steps) atomicAdd((unsigned long long*)&P.stats->steps, (unsigned long long)st_steps);
    if (st_replays) atomicAdd((unsigned long long*)&P.stats->exact_replays, (unsigned long long)st_replays);
    if (st_overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)st_overflow);
  }
}

template <int G, int MINB>
int grid_blocks_coop(const b2w_graph* g) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_uw_coop_kernel<G, MINB>, UW_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  return per_sm * g->num_sms;
}

template <int G, int MINB>
int grid_blocks(const b2w_graph* g) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_uw_kernel<G, MINB>, UW_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  return per_sm * g->num_sms;
}

int pick_group(const b2w_graph* g, uint32_t flags) {
  uint32_t forced = (flags >> 8) & 0xFF;                             // debug/tuning: bits 8..15 = group size
  if (forced == 8 || forced == 16 || forced == 32) return (int)forced;
  double avg = g->n ? (double)g->nnz / g->n : 0.0;
  // measured (B200): 8 lanes per walker win on flat low-degree graphs (ER, deg 20: 3.72 / 3.44 / 2.97 G steps/s
  // for 8 / 16 / 32 lanes with the first-session kernel; 5.38 / 4.31 for 8 / 16 with the interleaved searches),
  // 32 lanes win as soon as there are hub rows (power law: 1.77 vs 1.39 for 16)
  return (avg <= 32.0 && (double)g->max_degree <= 8.0 * avg + 16.0) ? 8 : 32;
}

uint32_t max_groups(const b2w_graph* g) {
  return (uint32_t)g->num_sms * 8u * (UW_THREADS / 8);               // upper bound: 8 resident CTAs of 8-lane groups
}

uint32_t gbm_stride(const b2w_graph* g) {
  uint32_t words = (g->max_degree + 31) / 32;
  return words > (uint32_t)UW_BW ? ((words + 3) & ~3u) : 0u;
}

}  // namespace

// Exactness precondition of the analytic normaliser: 1, f32(1/q), f32(1/p) are multiples of one power of two g
// and (max_degree + 1) * max(w) < 2^24 g.  Returns the exponent of g through `grid_exp` when eligible.
bool b2w_uw_grid(const b2w_graph* g, double p, double q, int* grid_exp) {
  if (!(g->flags & B2W_GRAPH_UNWEIGHTED)) return false;
  const float w[3] = {1.0f, (float)(1.0 / q), (float)(1.0 / p)};
  int low = 1000;
  float mx = 0.f;
  for (float v : w) {
    if (!(v > 0.f) || !(v < 3.0e38f)) return false;
    uint32_t bits;
    memcpy(&bits, &v, 4);
    uint32_t expo = (bits >> 23) & 0xFF, man = bits & 0x7FFFFF;
    if (expo == 0) return false;                                      // denormal
    man |= 0x800000;
    int tz = __builtin_ctz(man);
    int e = (int)expo - 127 - 23 + tz;                                // exponent of the lowest set bit
    if (e < low) low = e;
    if (v > mx) mx = v;
  }
  if (low < -100 || low > 100) return false;
  double lim = ldexp(1.0, 24 + low);
  if (!(((double)g->max_degree + 1.0) * (double)mx < lim)) return false;
  if (grid_exp) *grid_exp = low;
  return true;
}

bool b2w_uw_eligible(const b2w_graph* g, double p, double q) { return b2w_uw_grid(g, p, q, nullptr); }

size_t b2w_uw_work_bytes(const b2w_graph* g) {
  return 256 + (size_t)max_groups(g) * gbm_stride(g) * sizeof(uint32_t);
}

int b2w_launch_uw(const b2w_graph* g, const WalkParams& P_in, cudaStream_t s) {
  WalkParams P = P_in;
  char* base = reinterpret_cast<char*>(P_in.work);
  P.counter = reinterpret_cast<unsigned long long*>(base);
  UwConsts C;
  C.w_out = (float)(1.0 / P.q);
  C.w_ret = (float)(1.0 / P.p);
  int gexp = 0;
  if (!b2w_uw_grid(g, P.p, P.q, &gexp)) { b2w_set_error("walk_uw_kernel: graph / p / q not eligible"); return B2W_ERR_INVALID; }
  C.g = ldexpf(1.0f, gexp);
  C.a_in = (uint32_t)ldexp(1.0, -gexp);                              // exact integers by construction of g
  C.a_out = (uint32_t)ldexp((double)C.w_out, -gexp);
  C.a_ret = (uint32_t)ldexp((double)C.w_ret, -gexp);
  C.gbm = reinterpret_cast<uint32_t*>(base + 256);
  C.gbm_stride = gbm_stride(g);
  B2W_CUDA(cudaMemsetAsync(P.counter, 0, 8, s));
  const int G = pick_group(g, P.flags);
  const int MB = (int)((P.flags >> 16) & 0xF);                        // tuning: min resident CTAs per SM (0 = default)
  uint64_t groups_per_block = UW_THREADS / G;
  uint64_t need = (P.n_rows + groups_per_block - 1) / groups_per_block;
  // (An L2 persisting access-policy window over `indices` -- cudaLaunchAttributeAccessPolicyWindow, hitProp
  // persisting / missProp streaming -- was measured and dropped: 1.913 vs 1.919 G steps/s on config #3, 4.351 vs
  // 4.354 on config #2.  The hot rows already live in L2; what the kernel waits for is L2 latency, not DRAM.)
#define B2W_UW_LAUNCH(GG, BB)                                                        \
  do {                                                                               \
    int blocks = grid_blocks<GG, BB>(g);                                             \
    if ((uint64_t)blocks > need) blocks = (int)(need ? need : 1);                    \
    walk_uw_kernel<GG, BB><<<blocks, UW_THREADS, 0, s>>>(P, C);                      \
  } while (0)
  // G < 32 with B2W_FLAG_COOP: the cooperative kernel (long rows by the whole warp); bits 20..23 tune BIG = 16 << x
  const uint32_t bigx = (P.flags >> 20) & 0xF;
  const uint32_t BIG = bigx ? (16u << (bigx - 1)) : 64u;
  const bool coop = (P.flags & B2W_FLAG_COOP) != 0;   // opt-in: measured slower than the plain kernels
#define B2W_UW_COOP(GG, BB)                                                          \
  do {                                                                               \
    int blocks = grid_blocks_coop<GG, BB>(g);                                        \
    if ((uint64_t)blocks > need) blocks = (int)(need ? need : 1);                    \
    walk_uw_coop_kernel<GG, BB><<<blocks, UW_THREADS, 0, s>>>(P, C, BIG > 128u ? 128u : BIG); \
  } while (0)
  if (G == 8) { if (coop) B2W_UW_COOP(8, 4); else B2W_UW_LAUNCH(8, 4); }
  else if (G == 16) {
    if (coop) B2W_UW_COOP(16, 4);
    else if (MB == 4) B2W_UW_LAUNCH(16, 4); else if (MB == 6) B2W_UW_LAUNCH(16, 6); else B2W_UW_LAUNCH(16, 5);
  }
  else { if (MB == 4) B2W_UW_LAUNCH(32, 4); else if (MB == 6) B2W_UW_LAUNCH(32, 6); else B2W_UW_LAUNCH(32, 5); }   // measured: 5 CTAs/SM (48 regs) best
#undef B2W_UW_COOP
#undef B2W_UW_LAUNCH
  return b2w_cuda_fail(cudaGetLastError(), "walk_uw_kernel launch");
}
