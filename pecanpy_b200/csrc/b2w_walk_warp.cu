// b2w_walk_warp.cu -- SparseOTF: one warp per walker, persistent CTAs, dynamic row queue.
//
// Per step the warp streams row(cur) (indices + weights) with coalesced loads, resolves
// membership of every neighbour in row(prev) by a lane-parallel binary search, forms the biased
// weights, and must then reproduce -- bit for bit -- the reference's
//     probs = w / w.sum();  cdf = np.cumsum(probs);  choice = np.searchsorted(cdf, u)
// (pecanpy.py:546-559, rw/sparse_rw.py:51-130) whose sum and cumsum are SEQUENTIAL f32
// recurrences (numba/np/arraymath.py:161-174, :384-405).  A warp scan re-associates, so it is
// used as a FILTER with a rigorous error bound, never as the answer:
//
//  (1) normaliser S.  If every weight is a multiple of one power of two g and the total is
//      below 2^24 g, no f32 partial sum in ANY order can round, so the exact f64 warp reduction
//      equals the sequential f32 sum (checked per step; always true for unweighted graphs with
//      power-of-two p, q).  Otherwise S is accumulated sequentially from shared memory.
//  (2) probs_k = fdiv_rn(w_k, S) elementwise -- exact.
//  (3) cdf.  A_k = warp-scan prefix (f32).  Both the reference's sequential prefix and A_k are
//      floating-point summations of the same non-negative terms, so each is within
//      gamma_m * T_k of the exact prefix T_k (m = number of additions on the longest path):
//      |cdf_k - A_k| <= E_k := (k + chunk + 8) * 1.01 * 2^-24 * A_k.  If the first k with
//      A_k + E_k >= u also satisfies A_k - E_k >= u, then choice = k, provably.  Otherwise
//      (probability ~ deg^2 * 4e-8 per step) the warp replays the f32 recurrence exactly.
//
// Reference: pecanpy.py:164-210 (_random_walks), :522-561 (SparseOTF.get_move_forward).
#include "b2w_common.cuh"

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int CAP = 1024;   // floats of shared memory per warp (rows above this use global scratch)

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(B2W_FULL, v, o));
  return v;
}

struct WarpStats { uint32_t steps, replays, seqsums, overflow; };

template <bool EXTEND>
__device__ __forceinline__ uint32_t otf_choice_warp(const WalkParams& P, const int lane, const uint32_t cur,
                                                    const uint32_t cs, const uint32_t deg, const bool has_prev,
                                                    const uint32_t prev, const uint32_t ps, const uint32_t pdeg,
                                                    const double u, float* __restrict__ wbuf, WarpStats& st) {
  const uint32_t nchunks = (deg + 31) >> 5;
  const uint32_t nit = has_prev ? (32 - __clz(pdeg)) : 0;   // lower_bound iterations for pdeg elements
  float thr_cur = 0.f;
  if (EXTEND && has_prev) thr_cur = __ldg(P.thr + cur);

  // ---- pass A: biased weights -> wbuf, exactness statistics for the normaliser
  double acc = 0.0;
  unsigned long long orbits = 0ull;
  bool bad = false;
  for (uint32_t c = 0; c < nchunks; ++c) {
    const uint32_t k = (c << 5) + lane;
    const bool valid = k < deg;
    uint32_t x = 0xFFFFFFFFu;
    float wt = 0.f;
    if (valid) { x = __ldg(P.indices + cs + k); wt = __ldg(P.data + cs + k); }
    float w = wt;
    if (has_prev) {
      uint32_t lo = 0, hi = pdeg;
      bool common = false;
      for (uint32_t it = 0; it < nit; ++it) {
        if (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          uint32_t v = __ldg(P.indices + ps + mid);
          common |= (v == x);
          if (v < x) lo = mid + 1; else hi = mid;
        }
      }
      if (valid) {
        if (x == prev) {
          w = div_by(wt, P.p, P.invp_f, P.p_pow2);                     // return bias (sparse_rw.py:87/126)
        } else if (!EXTEND) {
          if (!common) w = div_by(wt, P.q, P.invq_f, P.q_pow2);        // out bias (:86)
        } else {
          bool out = true;
          float t = 0.f;
          if (common) {
            float wp = __ldg(P.data + ps + lo);
            float th = __ldg(P.thr + x);
            if (wp >= th) out = false; else t = __fdiv_rn(wp, th);     // (:273-276)
          }
          if (out) {
            double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), (double)t));   // (:119)
            if (wt < thr_cur) alpha = P.supp;                          // (:122-124)
            w = (float)__dmul_rn((double)wt, alpha);                   // (:125)
          }
        }
      }
    }
    if (valid) {
      wbuf[k] = w;
      double wd = (double)w;
      acc = __dadd_rn(acc, wd);
      // multiples of 2^-40 below 2^13 keep every f64 partial sum exact (53 bits)
      if (w == 0.f || (w >= 7.62939453125e-06f && w < 8192.f))
        orbits |= (unsigned long long)__double2ll_rn(wd * 1099511627776.0);
      else
        bad = true;
    }
  }
  // zero pad to a multiple of 4 for the float4 replay reads
  if (lane < 4) {
    uint32_t k = deg + lane;
    if (k < ((deg + 3) & ~3u)) wbuf[k] = 0.f;
  }
  __syncwarp();

  // ---- normaliser S (sequential f32 sum of the reference)
  const double T = warp_sum_f64(acc);
  const uint32_t or_lo = __reduce_or_sync(B2W_FULL, (uint32_t)orbits);
  const uint32_t or_hi = __reduce_or_sync(B2W_FULL, (uint32_t)(orbits >> 32));
  const bool anybad = __any_sync(B2W_FULL, bad);
  float S;
  bool exactS = false;
  if (!anybad && T < 8192.0 && T > 0.0) {
    int tz = or_lo ? (__ffs(or_lo) - 1) : (32 + __ffs(or_hi) - 1);
    // all weights are multiples of g = 2^(tz-40); exact if T < 2^24 * g
    double lim = scalbn(1.0, 24 + tz - 40);
    exactS = T < lim;
  }
  const uint32_t n4 = (deg + 3) >> 2;
  const float4* w4 = reinterpret_cast<const float4*>(wbuf);
  if (exactS) {
    S = (float)T;
  } else {
    float s = 0.f;
    for (uint32_t i = 0; i < n4; ++i) {
      float4 v = w4[i];
      s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
    }
    S = s;
    st.seqsums++;
  }

  // ---- cdf: warp-scan filter
  const bool force = (P.flags & B2W_FLAG_FORCE_EXACT_REPLAY) || !(S > 0.f) || !(S < 3.0e38f) || deg > (1u << 20);
  uint32_t choice = deg;       // default: no k with cdf_k >= u -> the reference's choice == deg overflow
  bool replay = force;
  uint32_t c = 0;
  if (!force) {
    float carry = 0.f;
    for (; c < nchunks; ++c) {
      const uint32_t k = (c << 5) + lane;
      const bool valid = k < deg;
      float pr = 0.f;
      if (valid) { pr = __fdiv_rn(wbuf[k], S); wbuf[k] = pr; }
      float a = pr;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(B2W_FULL, a, o);
        if (lane >= o) a = __fadd_rn(a, t);
      }
      const float A = __fadd_rn(carry, a);
      const float E = __fmul_rn(__fmul_rn((float)(k + c + 8), 6.0201e-08f), A);   // 1.01 * 2^-24
      const double lo_b = (double)A - (double)E, hi_b = (double)A + (double)E;
      const uint32_t bp = __ballot_sync(B2W_FULL, valid && (hi_b >= u));   // possibly  cdf_k >= u
      if (bp) {
        const uint32_t bd = __ballot_sync(B2W_FULL, valid && (lo_b >= u)); // certainly cdf_k >= u
        const int fp = __ffs(bp) - 1;
        if (bd && (__ffs(bd) - 1) == fp) choice = (c << 5) + fp; else replay = true;
        ++c;   // chunk c has been converted to probabilities
        break;
      }
      carry = __shfl_sync(B2W_FULL, A, 31);
    }
  }
  if (replay) {
    // ---- exact replay of the f32 recurrence
    for (; c < nchunks; ++c) {
      const uint32_t k = (c << 5) + lane;
      if (k < deg) wbuf[k] = __fdiv_rn(wbuf[k], S);
    }
    __syncwarp();
    float cdf = 0.f;
    choice = deg;
    for (uint32_t i = 0; i < n4; ++i) {
      float4 v = w4[i];
      cdf = __fadd_rn(cdf, v.x); if (!((double)cdf < u)) { choice = 4 * i; break; }
      cdf = __fadd_rn(cdf, v.y); if (!((double)cdf < u)) { choice = 4 * i + 1; break; }
      cdf = __fadd_rn(cdf, v.z); if (!((double)cdf < u)) { choice = 4 * i + 2; break; }
      cdf = __fadd_rn(cdf, v.w); if (!((double)cdf < u)) { choice = 4 * i + 3; break; }
    }
    if (choice > deg) choice = deg;   // a hit on a zero pad element means cdf[deg-1] < u was false earlier
    st.replays++;
  }
  if (choice == deg) st.overflow++;
  __syncwarp();
  return choice;
}

template <bool EXTEND>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) walk_sparse_warp_kernel(const WalkParams P) {
  __shared__ __align__(16) float smem_w[WARPS_PER_CTA][CAP];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t warp_gid = blockIdx.x * WARPS_PER_CTA + wib;
  float* gbuf = P.work + (size_t)warp_gid * P.work_stride;
  const uint32_t L = P.L;
  WarpStats st = {0, 0, 0, 0};

  for (;;) {
    unsigned long long i = 0;
    if (lane == 0) i = atomicAdd(P.counter, 1ull);
    i = __shfl_sync(B2W_FULL, i, 0);
    if (i >= P.n_rows) break;
    uint32_t* out = P.out + i * P.ld_out;

    uint32_t cur = __ldg(P.start + i);
    uint32_t prev = 0, ps = 0, pdeg = 0;
    uint32_t cs = __ldg(P.indptr + cur);
    uint32_t ce = __ldg(P.indptr + cur + 1);
    uint32_t eff = L + 1;
    uint32_t myval = (lane == 0) ? cur : 0u;   // lane (e & 31) holds output entry e until the block is flushed
    double my_u = 0.0;
    uint32_t j = 1;
    for (; j <= L; ++j) {
      const uint32_t deg = ce - cs;
      if (deg == 0) { eff = j; break; }                               // dead end (pecanpy.py:194-196,204-206)
      if (((j - 1) & 31) == 0) {
        // 32 steps' worth of uniforms, one Philox block per lane (counter = row, step, block 0)
        uint32_t sj = j + lane;
        if (sj <= L) my_u = step_uniform(P, i, sj);
      }
      const double u = __shfl_sync(B2W_FULL, my_u, (j - 1) & 31);
      float* wbuf = (deg + 4 <= CAP) ? smem_w[wib] : gbuf;
      const uint32_t choice = otf_choice_warp<EXTEND>(P, lane, cur, cs, deg, j > 1, prev, ps, pdeg, u, wbuf, st);
      const uint32_t nxt = __ldg(P.indices + cs + choice);            // unchecked, as pecanpy.py:559
      if (lane == (j & 31)) myval = nxt;
      if ((j & 31) == 31) {                                           // entries [j-31, j] complete: flush
        out[(j & ~31u) + lane] = myval;
        myval = 0u;
      }
      prev = cur; ps = cs; pdeg = deg;
      cur = nxt;
      cs = __ldg(P.indptr + cur);
      ce = __ldg(P.indptr + cur + 1);
      st.steps++;
    }
    // tail: entries of the current 32-block (zeros past the walk), later blocks all zero, then eff
    // j == first entry index not produced (dead end at step j, or L + 1): its 32-block is the
    // one still held in `myval` (all zero if the loop ended right after a flush)
    const uint32_t blk = j & ~31u;
    for (uint32_t base = blk; base < L + 2; base += 32) {
      uint32_t e = base + lane;
      uint32_t v = (base == blk) ? myval : 0u;
      if (e == L + 1) v = eff;
      if (e < L + 2) out[e] = v;
    }
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS) && lane == 0) {
    if (st.steps) atomicAdd((unsigned long long*)&P.stats->steps, (unsigned long long)st.steps);
    if (st.replays) atomicAdd((unsigned long long*)&P.stats->exact_replays, (unsigned long long)st.replays);
    if (st.seqsums) atomicAdd((unsigned long long*)&P.stats->seq_sums, (unsigned long long)st.seqsums);
    if (st.overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)st.overflow);
  }
}

template <bool EXTEND>
int grid_for(const b2w_graph* g) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_sparse_warp_kernel<EXTEND>, WARPS_PER_CTA * 32, 0);
  if (per_sm < 1) per_sm = 1;
  return per_sm * g->num_sms;
}

}  // namespace

uint32_t b2w_sparse_warp_total_warps(const b2w_graph* g) {
  int a = grid_for<false>(g), b = grid_for<true>(g);
  return (uint32_t)((a > b ? a : b) * WARPS_PER_CTA);
}

size_t b2w_sparse_warp_work_bytes(const b2w_graph* g) {
  // [0,256): row-queue counter; then one scratch row per resident warp for degrees above CAP-4
  size_t stride = (g->max_degree + 4 > (uint32_t)CAP) ? (((size_t)g->max_degree + 4 + 3) & ~(size_t)3) : 0;
  return 256 + (size_t)b2w_sparse_warp_total_warps(g) * stride * sizeof(float);
}

int b2w_launch_sparse_warp(const b2w_graph* g, const WalkParams& P_in, cudaStream_t s) {
  WalkParams P = P_in;
  char* base = reinterpret_cast<char*>(P_in.work);
  P.counter = reinterpret_cast<unsigned long long*>(base);
  P.work = reinterpret_cast<float*>(base + 256);
  P.work_stride = (g->max_degree + 4 > (uint32_t)CAP) ? (uint32_t)(((size_t)g->max_degree + 4 + 3) & ~(size_t)3) : 0;
  B2W_CUDA(cudaMemsetAsync(P.counter, 0, 8, s));
  const bool extend = P.extend != 0;
  int grid = extend ? grid_for<true>(g) : grid_for<false>(g);
  uint64_t need = (P.n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  if ((uint64_t)grid > need) grid = (int)(need ? need : 1);
  if (extend)
    walk_sparse_warp_kernel<true><<<grid, WARPS_PER_CTA * 32, 0, s>>>(P);
  else
    walk_sparse_warp_kernel<false><<<grid, WARPS_PER_CTA * 32, 0, s>>>(P);
  return b2w_cuda_fail(cudaGetLastError(), "walk_sparse_warp_kernel launch");
}
