// b2w_probs.cuh -- one-lane (sequential) evaluation of the biased transition weights.
//
// Used where the unit of work has no row-wide parallelism to offer: the alias-table builder
// (one lane per (node, neighbour-slot) table), the first step of PreComp walkers and the
// lane-per-walker SparseOTF kernel.  The order of floating-point operations is the
// reference's own (sequential left-to-right f32), so results are bit-identical by construction.
//
// Reference: rw/sparse_rw.py:51-91 (node2vec), :93-130 + :233-295 (node2vec+), :133-139 (get_nbrs).
#pragma once
#include "b2w_common.cuh"

// Streams the biased, UN-normalised weights w_0 .. w_{deg-1} of `cur` given `prev`.
// Rows are sorted and duplicate-free (validated at graph creation), for which the reference's
// merge (`isnotin`, rw/sparse_rw.py:142-230) is exactly set membership in N(prev).
template <bool EXTEND>
struct BiasStream {
  const WalkParams& P;
  uint32_t cs, deg, ps, pdeg, prev, idx2;
  bool has_prev;
  float thr_cur;

  __device__ __forceinline__ BiasStream(const WalkParams& P_, uint32_t cur, bool has_prev_, uint32_t prev_)
      : P(P_), prev(prev_), idx2(0), has_prev(has_prev_), thr_cur(0.f) {
    cs = P.indptr[cur];
    deg = P.indptr[cur + 1] - cs;
    ps = 0; pdeg = 0;
    if (has_prev) {
      ps = P.indptr[prev];
      pdeg = P.indptr[prev + 1] - ps;
      if (EXTEND) thr_cur = P.thr[cur];
    }
  }
  __device__ __forceinline__ void rewind() { idx2 = 0; }

  // must be called with k = 0, 1, 2, ... in order (the merge pointer only moves forward)
  __device__ __forceinline__ float weight(uint32_t k) {
    float wt = P.data[cs + k];
    if (!has_prev) return wt;
    uint32_t x = P.indices[cs + k];
    if (x == prev) return div_by(wt, P.p, P.invp_f, P.p_pow2);       // return bias (:87 / :126)
    while (idx2 < pdeg && P.indices[ps + idx2] < x) ++idx2;
    bool common = (idx2 < pdeg) && (P.indices[ps + idx2] == x);
    if (!EXTEND) {
      return common ? wt : div_by(wt, P.q, P.invq_f, P.q_pow2);       // out bias (:86)
    } else {
      float t = 0.f;
      if (common) {
        float wp = P.data[ps + idx2];
        float th = P.thr[x];
        if (wp >= th) return wt;                                      // tight in-edge (:273-274)
        t = __fdiv_rn(wp, th);                                        // f32 / f32 (:276)
      }
      double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), (double)t));   // (:119)
      if (wt < thr_cur) alpha = P.supp;                               // (:122-124)
      return (float)__dmul_rn((double)wt, alpha);                     // (:125)
    }
  }
};

// cumsum + searchsorted(left) of the normalised probabilities (pecanpy.py:556-557), one lane.
template <bool EXTEND>
__device__ __forceinline__ uint32_t otf_choice_seq(const WalkParams& P, uint32_t cur, bool has_prev,
                                                  uint32_t prev, double u) {
  BiasStream<EXTEND> bs(P, cur, has_prev, prev);
  float sum = 0.f;
  for (uint32_t k = 0; k < bs.deg; ++k) sum = __fadd_rn(sum, bs.weight(k));
  bs.rewind();
  float cdf = 0.f;
  for (uint32_t k = 0; k < bs.deg; ++k) {
    cdf = __fadd_rn(cdf, __fdiv_rn(bs.weight(k), sum));
    if (!((double)cdf < u)) return k;
  }
  return bs.deg;   // the reference's unchecked overflow ("choice == deg")
}
