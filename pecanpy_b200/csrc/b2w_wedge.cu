// b2w_wedge.cu -- SparseOTF on WEIGHTED graphs (node2vec and node2vec+, any p and q) through a weighted per-edge
// index: one lane per walker, O(log deg) per step, no row is streamed.
//
// For a step taken from `cur` after arriving over the stored edge e = (prev -> cur) the reference forms biased
// weights w'_k of row(cur) (rw/sparse_rw.py:51-130), S = their sequential f32 sum, probs = w' / S, a sequential f32
// cumsum and searchsorted(cdf, u) (pecanpy.py:546-559).  Everything but u is a function of the edge and of the bias
// parameters (p, q, node2vec+ thresholds), so it is prepared once per (graph, parameters):
//   * per row: the BASE biased weight of every slot, b_k = what the slot weighs when its neighbour is neither prev
//     nor a common neighbour (node2vec: w/q; node2vec+: w * (1/q, or min(1, 1/q) when w < thr[cur])) -- `bw`, f32 --
//     and the running f64 prefix sums of b -- `bq`;
//   * per edge, 32 bytes: next node, its degree and row start, where prev sits in row(cur) and its biased weight
//     w/p, the reference's exact f32 normaliser S of that (prev, cur) pair, and offsets of the two lists below;
//   * per edge, the EXCEPTIONS: the common neighbours whose biased weight differs from the base (node2vec: w;
//     node2vec+: w or w * alpha(t)), with position, exact f32 value, and the f64 prefix of the deviations;
//   * per edge into a row of >= 32 slots (whose return edge exists), CHECKPOINTS of the reference's exact f32 cdf
//     every 32 positions, laid out [node][slot of prev][j].
// A step then is: un-normalised prefix P(k) = bq[k] + deviations up to k (+ the return-edge deviation), the
// reference's cdf_k = (P(k) / S)(1 + t), |t| <= e_k = 1.02 (k + 3) 2^-24 + f64 slack (S is the reference's own f32
// sum, so only the divisions and the cumsum round), and the first k with P(k) >= u S (1 + e) found by a bisection
// over the exception list and a bisection over bq inside one segment.  If the element before it is provably below
// u S (1 - e) the choice is proven; otherwise (~1 % of the steps) the reference's recurrence is replayed exactly from
// the nearest checkpoint: at most 32 + window sequential additions instead of deg(cur).
//
// The first step of a walker has no edge (raw weights): evaluated like the oracle, by its lane on rows of <= 64
// slots and by the whole warp on longer ones (b2w_offedge.cuh); so is the step after the reference's unchecked
// choice == deg read (pecanpy.py:559).
//
// Reference: pecanpy.py:164-210, :522-561; rw/sparse_rw.py:51-130, :142-295.
#include <cmath>

#include "b2w_membership.cuh"
#include "b2w_offedge.cuh"
#include "b2w_probs.cuh"
#include "b2w_replay.cuh"
#include "b2w_rowout.cuh"
#include "b2w_scan.cuh"

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t KPF_POS_MASK = 0x3FFFFFFFu;
constexpr uint32_t KPF_NOTFOUND = 0x40000000u;
constexpr uint32_t KPF_HAS_EXC = 0x80000000u;
constexpr int WI_THREADS = 256;
constexpr uint32_t CKP = 32;                                          // checkpoint spacing (positions)

struct __align__(16) WRec {
  uint32_t nxt, kpf, exc, deg;       // exc: offset of the exception list (entries); in the count pass: its length
  uint32_t cs;                       // indptr[nxt]
  float S;                           // the reference's sequential f32 sum of the biased weights of row(nxt) given prev
  float vkp;                         // biased weight of the return edge, f32(w / p)
  float bkp;                         // base weight of the same slot (what vkp replaces)
};
static_assert(sizeof(WRec) == 32, "record size");

struct __align__(8) WExc {
  uint32_t pos;                      // position in row(cur)  (header entry of a list: the number of exceptions)
  float v;                           // exact biased weight of that slot for this edge
  double Pat;                        // un-normalised prefix AT pos: bq[pos] + D          (without the return edge)
  double D;                          // sum of (v - b) over the exceptions up to and including this one
};
static_assert(sizeof(WExc) == 24, "exception entry size");

// base biased weight of a slot (its neighbour neither prev nor common)
template <bool EXTEND>
__device__ __forceinline__ float base_weight(const WalkParams& P, const float wt, const float thr_cur) {
  if (!EXTEND) return div_by(wt, P.q, P.invq_f, P.q_pow2);           // rw/sparse_rw.py:86
  double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), 0.0));   // t = 0 (:119)
  if (wt < thr_cur) alpha = P.supp;                                   // (:122-124)
  return (float)__dmul_rn((double)wt, alpha);                         // (:125)
}

// biased weight of a COMMON neighbour slot (x != prev): wt = w(cur, x), wp = w(prev, x)
template <bool EXTEND>
__device__ __forceinline__ float common_weight(const WalkParams& P, const float wt, const float wp, const float thx,
                                               const float thr_cur) {
  if (!EXTEND) return wt;                                             // in-edge: unchanged
  if (wp >= thx) return wt;                                           // tight in-edge (:273-274)
  const float t = __fdiv_rn(wp, thx);                                 // (:276)
  double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), (double)t));
  if (wt < thr_cur) alpha = P.supp;
  return (float)__dmul_rn((double)wt, alpha);
}

// ---- pass 1: per row, base weights and their f64 prefix sums (one lane per row)
template <bool EXTEND>
__global__ void __launch_bounds__(WI_THREADS) wrow_kernel(const WalkParams P, float* __restrict__ bw, double* __restrict__ bq) {
  for (uint64_t r = blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x; r < P.n; r += (uint64_t)gridDim.x * WI_THREADS) {
    const uint32_t s = P.indptr[r], e = P.indptr[r + 1];
    const float thr_cur = EXTEND ? P.thr[r] : 0.f;
    double run = 0.0;
    for (uint32_t k = s; k < e; ++k) {
      const float b = base_weight<EXTEND>(P, P.data[k], thr_cur);
      bw[k] = b;
      run = __dadd_rn(run, (double)b);
      bq[k] = run;
    }
  }
}

__global__ void __launch_bounds__(WI_THREADS) wsrc_kernel(const uint32_t n, const uint32_t* __restrict__ indptr,
                                                          uint32_t* __restrict__ src) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * WI_THREADS) >> 5;
  for (uint64_t r = warp; r < n; r += nwarps) {
    const uint32_t s = __ldg(indptr + r), e = __ldg(indptr + r + 1);
    for (uint32_t k = s + lane; k < e; k += 32) src[k] = (uint32_t)r;
  }
}

// ---- pass 2 / 3: per edge, the exceptions (one warp per block of 32 edges, as b2w_edge_index.cu).
// FILL = false: records (exc = list length incl. header, ckp = number of checkpoints); FILL = true: the lists.
template <bool EXTEND, bool FILL>
__global__ void __launch_bounds__(WI_THREADS) wedge_kernel(const WalkParams P, const uint64_t nnz,
                                                           const uint32_t* __restrict__ src, WRec* __restrict__ rec,
                                                           const float* __restrict__ bw, const double* __restrict__ bq,
                                                           WExc* __restrict__ exc) {
  const Tile<32> T;
  const uint32_t lane = T.lane;
  const uint32_t* __restrict__ indptr = P.indptr;
  const uint32_t* __restrict__ indices = P.indices;
  const uint64_t warp = (blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * WI_THREADS) >> 5;
  const uint64_t nblk = (nnz + 1 + 31) >> 5;                          // the pad record [nnz] included
  for (uint64_t blk = warp; blk < nblk; blk += nwarps) {
    const uint64_t e = (blk << 5) + lane;
    uint32_t a = 0, b = 0, as = 0, ad = 0, bs = 0, bd = 0, kpf = 0, off = 0;
    const bool real = e < nnz;
    if (e <= nnz) {
      b = __ldg(indices + e);
      if (b < P.n) { bs = __ldg(indptr + b); bd = __ldg(indptr + b + 1) - bs; }
      if (real) { a = __ldg(src + e); as = __ldg(indptr + a); ad = __ldg(indptr + a + 1) - as; }
      if (FILL) { kpf = rec[e].kpf; off = rec[e].exc; }
    }
    uint32_t work = __ballot_sync(B2W_FULL, FILL ? (real && (kpf & KPF_HAS_EXC)) : (real && bd > 0));
    uint32_t my_cnt = 0, my_kpf = KPF_NOTFOUND;
    float my_vkp = 0.f, my_bkp = 0.f;
    while (work) {
      const int t = __ffs(work) - 1;
      work &= work - 1;
      const uint32_t ta = __shfl_sync(B2W_FULL, a, t), tb = __shfl_sync(B2W_FULL, b, t);
      const uint32_t tas = __shfl_sync(B2W_FULL, as, t), tad = __shfl_sync(B2W_FULL, ad, t);
      const uint32_t tbs = __shfl_sync(B2W_FULL, bs, t), tbd = __shfl_sync(B2W_FULL, bd, t);
      const uint32_t toff = __shfl_sync(B2W_FULL, off, t);
      const uint32_t* const arow = indices + tas;                     // row(prev)
      const uint32_t* const brow = indices + tbs;                     // row(cur)
      const uint32_t ka = 31 - __clz(tad), kb = 31 - __clz(tbd);
      const float thr_cur = EXTEND ? __ldg(P.thr + tb) : 0.f;
      if (!FILL) {
        bool found;
        const uint32_t pos = lower_bound_eq<true>(brow, tbd, ta, kb, found);
        if (lane == (uint32_t)t) {
          my_kpf = pos | (found ? 0u : KPF_NOTFOUND);
          if (found) {
            my_vkp = div_by(__ldg(P.data + tbs + pos), P.p, P.invp_f, P.p_pow2);   // rw/sparse_rw.py:87 / :126
            my_bkp = __ldg(bw + tbs + pos);
          }
        }
      }
      const uint32_t fwd_cost = ((tbd + 31) >> 5) * (ka + 3);
      const uint32_t rev_cost = ((tad + 31) >> 5) * (kb + 3);
      const bool fwd = fwd_cost <= rev_cost;
      const uint32_t nkeys = fwd ? tbd : tad;
      uint32_t m = 0;
      double carry = 0.0;                                             // deviations of the exceptions written so far
      for (uint32_t c0 = 0; c0 < nkeys; c0 += 32) {
        const uint32_t idx = c0 + lane;
        const bool valid = idx < nkeys;
        uint32_t kcur, iprev;                                         // position in row(cur) / in row(prev)
        bool found;
        uint32_t x;
        if (fwd) {                                                    // every neighbour of cur looked up in row(prev)
          x = valid ? __ldg(brow + idx) : B2W_NONE;
          iprev = lower_bound_eq<true>(arow, tad, x, ka, found);
          kcur = idx;
        } else {                                                      // every neighbour of prev looked up in row(cur)
          x = valid ? __ldg(arow + idx) : B2W_NONE;
          kcur = lower_bound_eq<true>(brow, tbd, x, kb, found);
          iprev = idx;
        }
        bool hit = valid && found && x != ta;
        float v = 0.f, bk = 0.f;
        if (hit) {
          const float wt = __ldg(P.data + tbs + kcur);
          const float wp = EXTEND ? __ldg(P.data + tas + iprev) : 0.f;
          const float thx = EXTEND ? __ldg(P.thr + x) : 0.f;
          v = common_weight<EXTEND>(P, wt, wp, thx, thr_cur);
          bk = __ldg(bw + tbs + kcur);
          hit = __float_as_uint(v) != __float_as_uint(bk);            // same weight as the base: not an exception
        }
        const uint32_t bal = __ballot_sync(B2W_FULL, hit);
        if (FILL) {
          double dev = hit ? __dsub_rn((double)v, (double)bk) : 0.0;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {                          // inclusive scan of the deviations, lane order
            const double y = __shfl_up_sync(B2W_FULL, dev, o);
            if (lane >= (uint32_t)o) dev = __dadd_rn(dev, y);
          }
          const double D = __dadd_rn(carry, dev);
          if (hit) {
            WExc x2;
            x2.pos = kcur; x2.v = v; x2.D = D;
            x2.Pat = __dadd_rn(__ldg(bq + tbs + kcur), D);
            exc[toff + 1 + m + __popc(bal & ((1u << lane) - 1u))] = x2;
          }
          carry = __shfl_sync(B2W_FULL, D, 31);
        }
        m += __popc(bal);
      }
      if (FILL) {
        if (lane == 0) { WExc h; h.pos = m; h.v = 0.f; h.Pat = 0.0; h.D = 0.0; exc[toff] = h; }
      } else if (lane == (uint32_t)t) {
        my_cnt = m;
      }
    }
    if (!FILL && e <= nnz) {
      if (my_cnt) my_kpf |= KPF_HAS_EXC;
      WRec r;
      r.nxt = b; r.kpf = my_kpf; r.exc = my_cnt ? my_cnt + 1 : 0u; r.deg = bd; r.cs = bs;
      r.S = 0.f; r.vkp = my_vkp; r.bkp = my_bkp;
      rec[e] = r;
    }
  }
}

// biased weight of slot k of the edge's row, exceptions merged in (xi = index of the next exception, advanced here)
struct EdgeRow {
  const float* __restrict__ bw;       // base weights of row(cur)
  const WExc* __restrict__ lst;       // exceptions of the edge
  uint32_t m, kp, xi, xpos;
  float vkp;
  __device__ __forceinline__ void seek(const uint32_t k0) {           // first exception with pos >= k0
    uint32_t lo = 0, hi = m;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(&lst[mid].pos) < k0) lo = mid + 1; else hi = mid; }
    xi = lo;
    xpos = xi < m ? __ldg(&lst[xi].pos) : NONE;
  }
  __device__ __forceinline__ float weight(const uint32_t k) {         // k must not decrease between calls
    if (k == xpos) {
      const float v = __ldg(&lst[xi].v);
      ++xi;
      xpos = xi < m ? __ldg(&lst[xi].pos) : NONE;
      return v;
    }
    return k == kp ? vkp : __ldg(bw + k);
  }
};

// ---- pass 4: per edge, the reference's exact normaliser and cdf checkpoints (one lane per edge, sequential)
__global__ void __launch_bounds__(WI_THREADS) wsum_kernel(const uint64_t nnz, WRec* __restrict__ rec,
                                                          const float* __restrict__ bw, const WExc* __restrict__ exc,
                                                          const uint32_t* __restrict__ ckb, float* __restrict__ ckpt) {
  for (uint64_t e = blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x; e < nnz; e += (uint64_t)gridDim.x * WI_THREADS) {
    const WRec r = rec[e];
    if (r.deg == 0) continue;
    EdgeRow row;
    row.bw = bw + r.cs;
    row.m = 0; row.lst = exc;
    if (r.kpf & KPF_HAS_EXC) { row.m = exc[r.exc].pos; row.lst = exc + r.exc + 1; }
    row.kp = (r.kpf & KPF_NOTFOUND) ? NONE : (r.kpf & KPF_POS_MASK);
    row.vkp = r.vkp;
    row.seek(0);
    float S = 0.f;
    for (uint32_t k = 0; k < r.deg; ++k) S = __fadd_rn(S, row.weight(k));   // sequential f32 sum (arraymath.py:161-174)
    rec[e].S = S;
    const uint32_t nck = r.deg / CKP;
    if (nck && row.kp != NONE) {                                      // checkpoints live at [node][slot of prev][j]
      row.seek(0);
      float cdf = 0.f;
      float* const out = ckpt + ((size_t)ckb[r.nxt] + (size_t)row.kp * nck);
      const uint32_t last = nck * CKP;
      for (uint32_t k = 0; k < last; ++k) {
        cdf = __fadd_rn(cdf, __fdiv_rn(row.weight(k), S));            // probs = w / S; sequential f32 cumsum
        if ((k & (CKP - 1)) == CKP - 1) out[k / CKP] = cdf;
      }
    }
  }
}

// =================================================================================================== the walk
struct WConsts {
  const WRec* __restrict__ rec;
  const WExc* __restrict__ exc;
  const float* __restrict__ bw;
  const double* __restrict__ bq;
  const float* __restrict__ ckpt;
  const uint32_t* __restrict__ ckb;  // per node: offset of its checkpoint block
  double slack;                      // relative slack for the f64 evaluation of P(k)
  int extend;
};

struct WStep {
  const double* __restrict__ q;      // bq of row(cur)
  const WExc* __restrict__ lst;
  uint32_t m, kp, d;
  double dk;                         // deviation of the return edge: vkp - b[kp]
  double inv_slope;                  // d / S: slots per unit of un-normalised weight, on average

  // un-normalised prefix at an exception / inside a segment
  __device__ __forceinline__ double at_exc(const uint32_t i) const {
    const uint32_t p = __ldg(&lst[i].pos);
    return __ldg(&lst[i].Pat) + (kp <= p ? dk : 0.0);
  }
  __device__ __forceinline__ double in_seg(const uint32_t k, const double Dseg) const {
    return __ldg(q + k) + Dseg + (kp <= k ? dk : 0.0);
  }
  // first k with P(k) >= T (d if none); Pk = P(k), Pprev = P(k - 1) (or -1 when k == 0)
  __device__ __forceinline__ uint32_t first_at_least(const double T, double& Pk, double& Pprev) const {
    uint32_t lo = 0, hi = m;                                          // first exception whose prefix reaches T
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (at_exc(mid) >= T) hi = mid; else lo = mid + 1;
    }
    const uint32_t i = lo;
    const double Dseg = i ? __ldg(&lst[i - 1].D) : 0.0;
    const uint32_t klo = i ? __ldg(&lst[i - 1].pos) + 1 : 0u;
    const uint32_t khi = i < m ? __ldg(&lst[i].pos) : d;              // the segment is [klo, khi)
    // first k in [klo, khi) with P(k) >= T.  Invariant: P(k) < T for every k < a;  P(b) >= T or b == khi.
    uint32_t a = klo, b = khi;
    if (b - a > 16u) {
      // The prefix grows by about S / d per slot: two interpolation probes land next to the answer, a gallop from
      // there closes the bracket within a cache line or two -- a plain bisection of a long row touches a new 64-byte
      // line of the f64 prefix array at all but its last three probes (444 DRAM bytes per step on BASELINE #3).
      const double Pa = i ? at_exc(i - 1) : 0.0;                      // P(klo - 1)
      double est = (double)a + (T - Pa) * inv_slope;
      uint32_t g = est <= (double)a ? a : (est >= (double)(b - 1) ? b - 1 : (uint32_t)est);
      double Pg = in_seg(g, Dseg);
      bool ge = Pg >= T;
      if (ge) b = g; else a = g + 1;
      if (a < b) {
        est = (double)g + (T - Pg) * inv_slope;
        const uint32_t g2 = est <= (double)a ? a : (est >= (double)(b - 1) ? b - 1 : (uint32_t)est);
        Pg = in_seg(g2, Dseg);
        ge = Pg >= T;
        if (ge) b = g2; else a = g2 + 1;
        if (ge) {                                                     // gallop towards smaller k from b
          for (uint32_t step = 1; a < b; step <<= 1) {
            const uint32_t c = (b - a > step) ? b - step : a;
            if (in_seg(c, Dseg) >= T) b = c; else { a = c + 1; break; }
          }
        } else {                                                      // gallop towards larger k from a
          for (uint32_t step = 1; a < b; step <<= 1) {
            const uint32_t c = (b - a > step) ? a + step - 1 : b - 1;
            if (in_seg(c, Dseg) >= T) { b = c; break; } else a = c + 1;
          }
        }
      }
    }
    while (a < b) {
      const uint32_t mid = (a + b) >> 1;
      if (in_seg(mid, Dseg) >= T) b = mid; else a = mid + 1;
    }
    const uint32_t k = a;
    if (k >= d) { Pk = 0.0; Pprev = 0.0; return d; }
    Pk = (k == khi) ? at_exc(i) : in_seg(k, Dseg);                   // k == khi < d: the exception itself
    if (k == 0) Pprev = -1.0;
    else if (k > klo) Pprev = in_seg(k - 1, Dseg);
    else Pprev = at_exc(i - 1);                                       // k == klo > 0: the previous exception sits at k - 1
    return k;
  }
};

// exact replay of the reference's recurrence from the checkpoint at or before k0
__device__ __noinline__ uint32_t wreplay(const WConsts& C, const WRec& r, const uint32_t k0, const double u) {
  EdgeRow row;
  row.bw = C.bw + r.cs;
  row.m = 0; row.lst = C.exc;
  if (r.kpf & KPF_HAS_EXC) { row.m = __ldg(&C.exc[r.exc].pos); row.lst = C.exc + r.exc + 1; }
  row.kp = (r.kpf & KPF_NOTFOUND) ? NONE : (r.kpf & KPF_POS_MASK);
  row.vkp = r.vkp;
  const uint32_t nck = r.deg / CKP;
  uint32_t j = row.kp != NONE ? k0 / CKP : 0u;                        // (no return edge: no checkpoints for this edge)
  if (j > nck) j = nck;
  float cdf = j ? __ldg(C.ckpt + ((size_t)__ldg(C.ckb + r.nxt) + (size_t)row.kp * nck + (j - 1))) : 0.f;   // after element j CKP - 1
  const uint32_t s = j * CKP;
  row.seek(s);
  const float ub = upper_float(u);                                    // cdf < u  <=>  cdf < ub
  for (uint32_t k = s; k < r.deg; ++k) {
    cdf = __fadd_rn(cdf, __fdiv_rn(row.weight(k), r.S));
    if (!(cdf < ub)) return k;                                        // (NaN compares false: choice k, like the reference)
  }
  return r.deg;                                                       // cdf[-1] < u: the reference's overflow
}

__device__ __noinline__ uint32_t woff_edge(const WalkParams& P, const int extend, const uint32_t cur, const bool has_prev,
                                           const uint32_t prev, const double u) {
  return extend ? otf_choice_seq<true>(P, cur, has_prev, prev, u) : otf_choice_seq<false>(P, cur, has_prev, prev, u);
}

// first step of a walker on a short row: sequential, one lane (longer rows: the whole warp, b2w_offedge.cuh)
__device__ __noinline__ uint32_t wfirst_lane(const WalkParams& P, const uint32_t cur, const double u) {
  return otf_choice_seq<false>(P, cur, false, 0, u);
}

__device__ __forceinline__ uint32_t wedge_step(const WConsts& C, const uint32_t flags, const WRec& r, const double u,
                                               uint32_t& st_replays) {
  const uint32_t d = r.deg;
  WStep W;
  W.q = C.bq + r.cs;
  W.d = d;
  W.m = 0; W.lst = C.exc;
  if (r.kpf & KPF_HAS_EXC) { W.m = __ldg(&C.exc[r.exc].pos); W.lst = C.exc + r.exc + 1; }
  W.kp = (r.kpf & KPF_NOTFOUND) ? NONE : (r.kpf & KPF_POS_MASK);
  W.dk = W.kp != NONE ? __dsub_rn((double)r.vkp, (double)r.bkp) : 0.0;
  const double S = (double)r.S;
  W.inv_slope = (double)d / S;
  const bool sane = r.S > 0.f && r.S < 3.0e38f && d <= 160000u;       // (a zero / overflowing sum: replay, like the reference)
  uint32_t k_replay = 0;
  if (sane) {
    const bool forced = (flags & B2W_FLAG_FORCE_EXACT_REPLAY) != 0;   // test hook: replay (from a checkpoint) every step
    const double EC = 1.02 * 5.9604644775390625e-08;                  // 1.02 * 2^-24
    const double uS = u * S;
    const double e_row = EC * (double)(d + 2) + C.slack;
    double Pk, Pprev;
    if (!forced) {
      // upper bound k1 of the answer from the most conservative "sure" threshold, thresholds at that position
      const uint32_t k1 = W.first_at_least(uS * (1.0 + e_row + 2.0 * e_row * e_row), Pk, Pprev);
      if (k1 < d) {
        const double e = EC * (double)(k1 + 3) + C.slack;
        const double t_poss = uS * (1.0 - e), t_sure = uS * (1.0 + e + 2.0 * e * e);
        if (Pprev < t_poss && Pk >= t_sure) return k1;                // (P(k1) >= t_hi >= t_sure) proven
        const uint32_t k2 = W.first_at_least(t_sure, Pk, Pprev);      // k2 <= k1, same e is valid
        if (k2 < d && Pprev < t_poss && Pk >= t_sure) return k2;
      }
    }
    // ambiguous: the answer is at or after the first k with P(k) >= u S (1 - e_row)
    k_replay = W.first_at_least(uS * (1.0 - e_row), Pk, Pprev);
    if (k_replay >= d) k_replay = d - 1;
  }
  ++st_replays;
  return wreplay(C, r, k_replay, u);
}

// The lanes of a warp stay converged at loop level, so that the steps without an edge -- the first step of a walker
// on a long row, the step after the reference's unchecked choice == deg read -- are evaluated by the whole warp.
constexpr uint32_t FIRST_COOP_DEG = 64;

template <int MINB, bool COOP>
__global__ void __launch_bounds__(WI_THREADS, MINB) walk_wedge_kernel(const WalkParams P, const WConsts C) {
  __shared__ uint32_t s_stage[8 * WI_THREADS];
  const uint32_t L = P.L;
  const uint32_t lane = threadIdx.x & 31;
  unsigned long long st_steps = 0;
  uint32_t st_replays = 0, st_overflow = 0;
  if (!COOP) {                                                        // plain per-lane loops (the default: 10.1 vs 8.7 G steps/s)
    for (uint64_t i = blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x; i < P.n_rows; i += (uint64_t)gridDim.x * WI_THREADS) {
      RowWriter<WI_THREADS> row;
      row.begin(P.out + i * P.ld_out, s_stage);
      uint32_t cur = __ldg(P.start + i), prev = 0;
      WRec r;
      r.cs = __ldg(P.indptr + cur);
      r.deg = __ldg(P.indptr + cur + 1) - r.cs;
      uint32_t eff = L + 1;
      bool edge_ok = false;                                             // the first step has no edge
      row.push(0, cur);
      uint32_t j = 1;
      for (; j <= L; ++j) {
        const uint32_t d = r.deg;
        if (d == 0) { eff = j; break; }                                 // pecanpy.py:194-196, 204-206
        const double u = step_uniform(P, i, j);
        uint32_t choice;
        if (edge_ok) choice = wedge_step(C, P.flags, r, u, st_replays);
        else choice = woff_edge(P, C.extend, cur, j > 1, prev, u);      // first step / after an unchecked choice == deg read
        if (choice == d) ++st_overflow;
        edge_ok = choice < d;
        const uint4* rp = reinterpret_cast<const uint4*>(C.rec + (r.cs + choice));   // [cs + d]: next row's first edge (:559)
        const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        prev = cur;
        r.nxt = r0.x; r.kpf = r0.y; r.exc = r0.z; r.deg = r0.w;
        r.cs = r1.x; r.S = __uint_as_float(r1.y); r.vkp = __uint_as_float(r1.z); r.bkp = __uint_as_float(r1.w);
        cur = r.nxt;
        row.push(j, cur);
      }
      st_steps += eff - 1;
      for (uint32_t z = j; z <= L; ++z) row.push(z, 0u);                // zero tail (np.zeros, pecanpy.py:182)
      row.push(L + 1, eff);
      row.finish(L + 2);
    }
  }
  for (uint64_t i = blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x; COOP && __any_sync(B2W_FULL, i < P.n_rows);
       i += (uint64_t)gridDim.x * WI_THREADS) {
    const bool alive = i < P.n_rows;
    RowWriter<WI_THREADS> row;
    uint32_t cur = 0, prev = 0;
    WRec r;
    r.cs = 0; r.deg = 0;
    if (alive) {
      row.begin(P.out + i * P.ld_out, s_stage);
      cur = __ldg(P.start + i);
      r.cs = __ldg(P.indptr + cur);
      r.deg = __ldg(P.indptr + cur + 1) - r.cs;
      row.push(0, cur);
    }
    uint32_t eff = L + 1;
    bool walking = alive;
    bool edge_ok = false;                                             // the first step has no edge
    for (uint32_t j = 1; j <= L; ++j) {
      const uint32_t d = r.deg;
      if (walking && d == 0) { eff = j; walking = false; }            // pecanpy.py:194-196, 204-206
      uint32_t choice = 0;
      double u = 0.0;
      bool coop = false;
      if (walking) {
        u = step_uniform(P, i, j);
        if (edge_ok) choice = wedge_step(C, P.flags, r, u, st_replays);
        else if (j == 1 && d <= FIRST_COOP_DEG) choice = wfirst_lane(P, cur, u);
        else coop = true;
      }
      __syncwarp();
      uint32_t off = __ballot_sync(B2W_FULL, coop);
      while (off) {
        const int src = __ffs(off) - 1;
        off &= off - 1;
        const uint32_t bcur = __shfl_sync(B2W_FULL, cur, src), bprev = __shfl_sync(B2W_FULL, prev, src);
        const double bu = __shfl_sync(B2W_FULL, u, src);
        const bool bhp = j > 1;
        const uint32_t c = C.extend ? offedge_w_warp<true>(P, bcur, bhp, bprev, bu) : offedge_w_warp<false>(P, bcur, bhp, bprev, bu);
        if (lane == (uint32_t)src) choice = c;
      }
      if (walking) {
        if (choice == d) ++st_overflow;
        edge_ok = choice < d;
        const uint4* rp = reinterpret_cast<const uint4*>(C.rec + (r.cs + choice));   // [cs + d]: next row's first edge (:559)
        const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        prev = cur;
        r.nxt = r0.x; r.kpf = r0.y; r.exc = r0.z; r.deg = r0.w;
        r.cs = r1.x; r.S = __uint_as_float(r1.y); r.vkp = __uint_as_float(r1.z); r.bkp = __uint_as_float(r1.w);
        cur = r.nxt;
      }
      if (alive) row.push(j, walking ? cur : 0u);                     // zero tail after a dead end (np.zeros, pecanpy.py:182)
    }
    if (alive) {
      st_steps += eff - 1;
      row.push(L + 1, eff);
      row.finish(L + 2);
    }
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS)) {
    for (int o = 16; o; o >>= 1) {
      st_steps += __shfl_xor_sync(B2W_FULL, st_steps, o);
      st_replays += __shfl_xor_sync(B2W_FULL, st_replays, o);
      st_overflow += __shfl_xor_sync(B2W_FULL, st_overflow, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (st_steps) atomicAdd((unsigned long long*)&P.stats->steps, st_steps);
      if (st_replays) atomicAdd((unsigned long long*)&P.stats->exact_replays, (unsigned long long)st_replays);
      if (st_overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)st_overflow);
    }
  }
}

// per node: floats of its checkpoint block = deg * (deg / CKP), to be prefix-summed
__global__ void __launch_bounds__(WI_THREADS) wckpt_count_kernel(const uint32_t n, const uint32_t* __restrict__ indptr,
                                                                 uint32_t* __restrict__ ckb, unsigned int* __restrict__ too_big) {
  for (uint64_t v = blockIdx.x * (uint64_t)WI_THREADS + threadIdx.x; v <= n; v += (uint64_t)gridDim.x * WI_THREADS) {
    uint32_t c = 0;
    if (v < n) {
      const uint32_t d = indptr[v + 1] - indptr[v];
      const unsigned long long t = (unsigned long long)d * (d / CKP);
      if (t >= 0xFFFFFFFFull) atomicOr(too_big, 1u);
      c = (uint32_t)t;
    }
    ckb[v] = c;
  }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

void b2w_fill_bias_params(WalkParams& P, double p, double q);

static WalkParams wparams(const b2w_graph* g, double p, double q, int extend, const float* d_thr) {
  WalkParams P{};
  P.n = g->n; P.indptr = g->indptr; P.indices = g->indices; P.data = g->data; P.thr = d_thr;
  P.extend = extend ? 1 : 0;
  b2w_fill_bias_params(P, p, q);
  return P;
}

extern "C" size_t b2w_windex_work_bytes(const b2w_graph* g) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) return 0;
  return align256(((size_t)g->nnz + 1) * sizeof(uint32_t)) + b2w_scan::work_bytes(g->nnz + 1 + g->n) + 512;
}

static int wcheck(const b2w_graph* g, const void* d_rec, const void* d_bw, const void* d_bq, double p, double q, int extend,
                  const float* d_thr, const char* what) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("%s: CSR graph handle required", what); return B2W_ERR_INVALID; }
  if (!d_rec || !d_bw || !d_bq) { b2w_set_error("%s: null array", what); return B2W_ERR_INVALID; }
  if ((reinterpret_cast<uintptr_t>(d_rec) & 31) != 0 || (reinterpret_cast<uintptr_t>(d_bq) & 7) != 0) { b2w_set_error("%s: record array must be 32-byte aligned", what); return B2W_ERR_INVALID; }
  if (!(p > 0.0) || !(q > 0.0) || !std::isfinite(p) || !std::isfinite(q)) { b2w_set_error("%s: p and q must be finite and > 0", what); return B2W_ERR_INVALID; }
  if (extend && !d_thr) { b2w_set_error("%s: node2vec+ needs the noise thresholds", what); return B2W_ERR_INVALID; }
  if (g->max_degree > KPF_POS_MASK - 1) { b2w_set_error("%s: max degree beyond 2^30", what); return B2W_ERR_UNSUPPORTED; }
  return B2W_OK;
}

extern "C" int b2w_windex_prepare(const b2w_graph* g, double p, double q, int extend, const float* d_thr, void* d_rec,
                                  float* d_bw, double* d_bq, uint32_t* d_ckb, void* d_work, size_t work_bytes,
                                  uint64_t* h_exc_entries, uint64_t* h_ckpt_floats, void* stream) {
  int rc = wcheck(g, d_rec, d_bw, d_bq, p, q, extend, d_thr, "b2w_windex_prepare");
  if (rc) return rc;
  if (!h_exc_entries || !h_ckpt_floats || !d_ckb) { b2w_set_error("b2w_windex_prepare: null output"); return B2W_ERR_INVALID; }
  if (!d_work || work_bytes < b2w_windex_work_bytes(g)) { b2w_set_error("b2w_windex_prepare: scratch too small"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  const WalkParams P = wparams(g, p, q, extend, d_thr);
  uint32_t* src = reinterpret_cast<uint32_t*>(d_work);
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_work) + align256(((size_t)g->nnz + 1) * sizeof(uint32_t)));
  WRec* rec = reinterpret_cast<WRec*>(d_rec);
  const unsigned grid = (unsigned)g->num_sms * 8;
  wsrc_kernel<<<grid, WI_THREADS, 0, s>>>(g->n, g->indptr, src);
  if (extend) {
    wrow_kernel<true><<<grid, WI_THREADS, 0, s>>>(P, d_bw, d_bq);
    wedge_kernel<true, false><<<grid, WI_THREADS, 0, s>>>(P, g->nnz, src, rec, d_bw, d_bq, nullptr);
  } else {
    wrow_kernel<false><<<grid, WI_THREADS, 0, s>>>(P, d_bw, d_bq);
    wedge_kernel<false, false><<<grid, WI_THREADS, 0, s>>>(P, g->nnz, src, rec, d_bw, d_bq, nullptr);
  }
  B2W_CUDA(cudaGetLastError());
  const uint64_t count = g->nnz + 1;
  unsigned long long tot_exc = 0, tot_ck = 0;
  B2W_CUDA(b2w_scan::exclusive_scan(count, reinterpret_cast<uint32_t*>(d_rec) + 2, 8, sums, &tot_exc, s));   // .exc
  unsigned int* flag = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(d_work) + work_bytes - 256);
  B2W_CUDA(cudaMemsetAsync(flag, 0, 4, s));
  wckpt_count_kernel<<<(unsigned)g->num_sms * 4, WI_THREADS, 0, s>>>(g->n, g->indptr, d_ckb, flag);
  B2W_CUDA(cudaGetLastError());
  B2W_CUDA(b2w_scan::exclusive_scan((uint64_t)g->n + 1, d_ckb, 1, sums, &tot_ck, s));
  unsigned int h_flag = 0;
  B2W_CUDA(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, s));
  B2W_CUDA(cudaStreamSynchronize(s));
  *h_exc_entries = tot_exc;
  *h_ckpt_floats = tot_ck;
  if (h_flag) tot_ck = 0xFFFFFFFFull;
  if (tot_exc >= 0xFFFFFFFFull || tot_ck >= 0xFFFFFFFFull) {
    b2w_set_error("b2w_windex_prepare: lists do not fit 32-bit offsets (%llu exceptions, %llu checkpoints)", tot_exc, tot_ck);
    return B2W_ERR_UNSUPPORTED;
  }
  return B2W_OK;
}

extern "C" int b2w_windex_finish(b2w_graph* g, double p, double q, int extend, const float* d_thr, void* d_rec,
                                 float* d_bw, double* d_bq, const uint32_t* d_ckb, void* d_exc, uint64_t exc_entries,
                                 float* d_ckpt, uint64_t ckpt_floats, void* d_work, size_t work_bytes, void* stream) {
  int rc = wcheck(g, d_rec, d_bw, d_bq, p, q, extend, d_thr, "b2w_windex_finish");
  if (rc) return rc;
  if ((exc_entries && !d_exc) || (ckpt_floats && !d_ckpt) || !d_ckb) { b2w_set_error("b2w_windex_finish: null list array"); return B2W_ERR_INVALID; }
  if (!d_work || work_bytes < b2w_windex_work_bytes(g)) { b2w_set_error("b2w_windex_finish: scratch too small"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  const WalkParams P = wparams(g, p, q, extend, d_thr);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(d_work);
  WRec* rec = reinterpret_cast<WRec*>(d_rec);
  const unsigned grid = (unsigned)g->num_sms * 8;
  if (exc_entries) {
    if (extend) wedge_kernel<true, true><<<grid, WI_THREADS, 0, s>>>(P, g->nnz, src, rec, d_bw, d_bq, reinterpret_cast<WExc*>(d_exc));
    else wedge_kernel<false, true><<<grid, WI_THREADS, 0, s>>>(P, g->nnz, src, rec, d_bw, d_bq, reinterpret_cast<WExc*>(d_exc));
  }
  wsum_kernel<<<(unsigned)g->num_sms * 16, WI_THREADS, 0, s>>>(g->nnz, rec, d_bw, reinterpret_cast<const WExc*>(d_exc), d_ckb, d_ckpt);
  B2W_CUDA(cudaGetLastError());
  B2W_CUDA(cudaStreamSynchronize(s));                                 // complete before any walk may use it
  g->w_rec = d_rec; g->w_exc = d_exc; g->w_bw = d_bw; g->w_bq = d_bq; g->w_ckpt = d_ckpt; g->w_ckb = d_ckb;
  g->w_p = p; g->w_q = q; g->w_extend = extend ? 1 : 0; g->w_thr = extend ? d_thr : nullptr;
  g->flags |= B2W_GRAPH_HAS_WINDEX;
  return B2W_OK;
}

extern "C" int b2w_graph_clear_windex(b2w_graph* g) {
  if (!g) { b2w_set_error("clear_windex: null graph"); return B2W_ERR_INVALID; }
  g->w_rec = nullptr; g->w_exc = nullptr; g->w_bw = nullptr; g->w_bq = nullptr; g->w_ckpt = nullptr; g->w_thr = nullptr;
  g->flags &= ~B2W_GRAPH_HAS_WINDEX;
  return B2W_OK;
}

bool b2w_windex_matches(const b2w_graph* g, double p, double q, int extend, const float* d_thr) {
  return (g->flags & B2W_GRAPH_HAS_WINDEX) && g->w_p == p && g->w_q == q && g->w_extend == (extend ? 1 : 0) &&
         (!extend || g->w_thr == d_thr);
}

int b2w_launch_wedge(const b2w_graph* g, const WalkParams& P, cudaStream_t s) {
  if (!b2w_windex_matches(g, P.p, P.q, P.extend, P.thr)) { b2w_set_error("walk_wedge_kernel: no weighted edge index for these parameters"); return B2W_ERR_INVALID; }
  WConsts C;
  C.rec = reinterpret_cast<const WRec*>(g->w_rec);
  C.exc = reinterpret_cast<const WExc*>(g->w_exc);
  C.bw = g->w_bw; C.bq = g->w_bq; C.ckpt = g->w_ckpt; C.ckb = g->w_ckb;
  C.extend = P.extend;
  // f64 evaluation of P(k): (deg + #exceptions + 2) additions of magnitude <= R * P(k), R = the largest ratio between
  // a base weight and the weight that replaces it (the deviations may cancel most of a prefix)
  const double ratios[4] = {P.p, P.q, 1.0 / P.p, 1.0 / P.q};
  double R = 1.0;
  for (double v : ratios) if (v > R) R = v;
  C.slack = ((double)g->max_degree * 2.0 + 16.0) * 1.1102230246251565e-16 * R * R * 4.0;
  const uint64_t want = (P.n_rows + WI_THREADS - 1) / WI_THREADS;
  const uint64_t cap = (uint64_t)g->num_sms * 64;
  unsigned blocks = (unsigned)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  const int mb = (int)((P.flags >> 16) & 0xF);                        // tuning: resident CTAs per SM (0 = default)
  const bool coop = (P.flags & B2W_FLAG_OFFEDGE_WARP) != 0;          // opt-in: measured slower here (plain loops are the default)
  if (coop) walk_wedge_kernel<4, true><<<blocks, WI_THREADS, 0, s>>>(P, C);
  else if (mb == 5) walk_wedge_kernel<5, false><<<blocks, WI_THREADS, 0, s>>>(P, C);
  else if (mb == 3) walk_wedge_kernel<3, false><<<blocks, WI_THREADS, 0, s>>>(P, C);
  else walk_wedge_kernel<4, false><<<blocks, WI_THREADS, 0, s>>>(P, C);
  return b2w_cuda_fail(cudaGetLastError(), "walk_wedge_kernel launch");
}
