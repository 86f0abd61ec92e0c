// b2w_offedge.cuh -- a SparseOTF step evaluated from the rows by ALL 32 lanes of a warp, exactly as the reference.
//
// The lane-per-walker kernels (b2w_walk_edge.cu, b2w_wedge.cu) lean on what the per-edge index knows about the edge a
// walker arrived over.  Two kinds of steps have no such edge: the first step of a walker, and the step after the
// reference's unchecked `indices[indptr[cur] + choice]` read with choice == deg (pecanpy.py:559, ~1e-7 of the steps,
// almost all of them on hub rows, and the node that read lands on -- the smallest id of the next row -- is a hub
// again on graphs whose low ids are the high degrees).  Evaluated by one lane, the sorted merge of two ~4000-entry
// rows took ~1 ms and set the tail of every launch.  Here the walker's warp does it together: the
// other 31 lanes are waiting for that lane anyway.  Membership by lane-parallel binary search, then the reference's
// own recurrences -- sequential f32 sum, fdiv, sequential f32 cumsum -- carried redundantly by every lane over
// shuffled weights, so the result is exact by construction and warp-uniform.
//
// All functions must be called by the 32 converged lanes of a warp with uniform arguments.
#pragma once
#include "b2w_membership.cuh"
#include "b2w_replay.cuh"

// Unweighted graph, biases on the exact grid (b2w_uw_grid): S is the exact count-weighted total.
static __device__ __noinline__ uint32_t offedge_uw_warp(const uint32_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                                 const int a_in, const int a_out, const int a_ret, const float g,
                                                 const uint32_t cur, const uint32_t prev, const double u) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cs = __ldg(indptr + cur), d = __ldg(indptr + cur + 1) - cs;
  const uint32_t ps = __ldg(indptr + prev), pd = __ldg(indptr + prev + 1) - ps;
  const uint32_t* const crow = indices + cs;
  const uint32_t* const prow = indices + ps;
  const uint32_t kpd = 31 - __clz(pd | 1u);
  uint32_t m = 0;
  bool has_kp = false;
  for (uint32_t c0 = 0; c0 < d; c0 += 32) {                           // pass 1: counts -> the exact normaliser
    const uint32_t k = c0 + lane;
    const bool valid = k < d;
    const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
    bool found = false;
    if (pd) lower_bound_eq<true>(prow, pd, x, kpd, found);
    const bool isprev = valid && x == prev;
    m += __popc(__ballot_sync(B2W_FULL, valid && found && !isprev));
    has_kp = has_kp || __any_sync(B2W_FULL, isprev);
  }
  const int Wd = (int)d * a_out + (int)m * (a_in - a_out) + (has_kp ? a_ret - a_out : 0);
  const float S = __fmul_rn(__int2float_rn(Wd), g);                   // exact: W_d < 2^24, g a power of two
  const float fa = __fdiv_rn(__fmul_rn(__int2float_rn(a_in), g), S);
  const float fo = __fdiv_rn(__fmul_rn(__int2float_rn(a_out), g), S);
  const float fp = __fdiv_rn(__fmul_rn(__int2float_rn(a_ret), g), S);
  const float ub = upper_float(u);                                    // cdf < u  <=>  cdf < ub
  float cdf = 0.f;
  for (uint32_t c0 = 0; c0 < d; c0 += 32) {                           // pass 2: the sequential f32 cumsum
    const uint32_t k = c0 + lane;
    const bool valid = k < d;
    const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
    bool found = false;
    if (pd) lower_bound_eq<true>(prow, pd, x, kpd, found);
    const float w = (valid && x == prev) ? fp : (found ? fa : fo);
    const uint32_t nv = min(32u, d - c0);
    for (uint32_t t = 0; t < nv; ++t) {
      cdf = __fadd_rn(cdf, __shfl_sync(B2W_FULL, w, (int)t));
      if (cdf >= ub) return c0 + t;
    }
  }
  return d;                                                           // cdf[-1] < u: the reference's overflow
}

// Any CSR graph: node2vec (EXTEND = false) or node2vec+ weights as rw/sparse_rw.py:51-130, in the reference's order.
template <bool EXTEND>
__device__ __forceinline__ float offedge_weight(const WalkParams& P, const uint32_t cs, const uint32_t ps,
                                                const uint32_t pd, const uint32_t kpd, const uint32_t prev,
                                                const bool has_prev, const float thr_cur, const uint32_t k,
                                                const bool valid) {
  // (every lane runs the search: its trip count is uniform)
  const uint32_t x = (valid && has_prev) ? __ldg(P.indices + cs + k) : B2W_NONE;
  bool common = false;
  uint32_t pos = 0;
  if (has_prev && pd) pos = lower_bound_eq<true>(P.indices + ps, pd, x, kpd, common);
  if (!valid) return 0.f;
  const float wt = __ldg(P.data + cs + k);
  if (!has_prev) return wt;
  if (x == prev) return div_by(wt, P.p, P.invp_f, P.p_pow2);          // return bias (:87 / :126)
  if (!EXTEND) return common ? wt : div_by(wt, P.q, P.invq_f, P.q_pow2);   // out bias (:86)
  float t = 0.f;
  if (common) {
    const float wp = __ldg(P.data + ps + pos);
    const float th = __ldg(P.thr + x);
    if (wp >= th) return wt;                                          // tight in-edge (:273-274)
    t = __fdiv_rn(wp, th);                                            // (:276)
  }
  double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), (double)t));   // (:119)
  if (wt < thr_cur) alpha = P.supp;                                   // (:122-124)
  return (float)__dmul_rn((double)wt, alpha);                         // (:125)
}

template <bool EXTEND>
__device__ __noinline__ uint32_t offedge_w_warp(const WalkParams& P, const uint32_t cur, const bool has_prev,
                                                const uint32_t prev, const double u) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cs = __ldg(P.indptr + cur), d = __ldg(P.indptr + cur + 1) - cs;
  uint32_t ps = 0, pd = 0;
  float thr_cur = 0.f;
  if (has_prev) {
    ps = __ldg(P.indptr + prev);
    pd = __ldg(P.indptr + prev + 1) - ps;
    if (EXTEND) thr_cur = __ldg(P.thr + cur);
  }
  const uint32_t kpd = 31 - __clz(pd | 1u);
  float S = 0.f;
  for (uint32_t c0 = 0; c0 < d; c0 += 32) {                           // sequential f32 sum (arraymath.py:161-174)
    const float w = offedge_weight<EXTEND>(P, cs, ps, pd, kpd, prev, has_prev, thr_cur, c0 + lane, c0 + lane < d);
    const uint32_t nv = min(32u, d - c0);
    for (uint32_t t = 0; t < nv; ++t) S = __fadd_rn(S, __shfl_sync(B2W_FULL, w, (int)t));
  }
  const float ub = upper_float(u);
  float cdf = 0.f;
  for (uint32_t c0 = 0; c0 < d; c0 += 32) {                           // probs = w / S; sequential f32 cumsum
    const float w = offedge_weight<EXTEND>(P, cs, ps, pd, kpd, prev, has_prev, thr_cur, c0 + lane, c0 + lane < d);
    const float pr = __fdiv_rn(w, S);
    const uint32_t nv = min(32u, d - c0);
    for (uint32_t t = 0; t < nv; ++t) {
      cdf = __fadd_rn(cdf, __shfl_sync(B2W_FULL, pr, (int)t));
      if (!(cdf < ub)) return c0 + t;                                 // (NaN compares false: choice, like the reference)
    }
  }
  return d;
}
