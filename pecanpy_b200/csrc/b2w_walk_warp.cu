// b2w_walk_warp.cu -- generic SparseOTF (weighted graphs, node2vec+, any p and q): one warp per
// walker, persistent CTAs, dynamic row queue.
//
// Per step the warp must reproduce -- bit for bit -- the reference's
//     probs = w / w.sum();  cdf = np.cumsum(probs);  choice = np.searchsorted(cdf, u)
// (pecanpy.py:546-559, rw/sparse_rw.py:51-130) whose sum and cumsum are SEQUENTIAL f32 recurrences
// (numba/np/arraymath.py:161-174, :384-405).  Parallel arithmetic is used as a FILTER with a rigorous
// error bound, never as the answer:
//
//  phase 1  (node2vec) membership bitmap of N(cur) & N(prev) over the positions of row(cur), searching
//           the cheaper side (b2w_membership.cuh); (node2vec+) per-element lower_bound in row(prev),
//           because each common neighbour also needs w(prev, x) / thr[x].
//  phase 2  stream the weights of row(cur) once (coalesced), form the biased weights w_k, stage them
//           in shared memory (rows above 1020 entries: a per-warp scratch row that stays in L2); then every
//           lane sums one contiguous segment of the staged row (odd pitch: bank-conflict free) and a single
//           warp scan yields the segment prefixes.
//  filter   Let P_k be the exact prefix sums of w.  The reference computes S = fl-sum(w) (relative error
//           gamma_{d-1}), probs_k = fl(w_k / S) (2^-24 each) and cdf_k = fl-cumsum (gamma_k); all terms are
//           non-negative, so cdf_k = (P_k / P_{d-1}) (1 +- (gamma_k + gamma_{d-1} + 2^-24)).  The warp's own
//           segment sums / scans are float summations of the same terms with depth <= mseg + 16 (mseg =
//           elements per lane).  Hence |cdf_k - A_k / T| <= e_k A_k / T with e_k = (k + d + 2 mseg + 48) 1.01 2^-24, and the
//           first k with A_k (1 + e_k) >= u T is the reference's choice whenever A_k (1 - e_k) >= u T.
//  replay   otherwise (probability ~ d^2 1e-7 per step) the warp re-runs the reference's recurrences
//           exactly: sequential f32 sum, fdiv per element, sequential f32 cumsum (float4 broadcast reads).
//
// Reference: pecanpy.py:164-210 (_random_walks), :522-561 (SparseOTF.get_move_forward).
#include "b2w_membership.cuh"

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int CAP = 1024;   // staged weights per warp in shared memory (longer rows use global scratch)

struct __align__(16) WarpBuf {
  float w[CAP];
  uint32_t bm[CAP / 32];
};

struct WarpStats { uint32_t steps, replays, seqsums, overflow; };

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(B2W_FULL, v, o));
  return v;
}

__device__ __forceinline__ float warp_incl_scan_f32(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(B2W_FULL, v, o);
    if (lane >= o) v = __fadd_rn(v, t);
  }
  return v;
}

template <bool EXTEND>
__device__ __forceinline__ uint32_t otf_choice_warp(const WalkParams& P, const Tile<32>& T, const uint32_t cur,
                                                    const uint32_t cs, const uint32_t d, const bool has_prev,
                                                    const uint32_t prev, const uint32_t ps, const uint32_t pdeg,
                                                    const double u, float* __restrict__ wbuf,
                                                    uint32_t* __restrict__ bm,
                                                    WarpStats& st) {
  const int lane = T.lane;
  const uint32_t nchunks = (d + 31) >> 5;
  const uint32_t* const crow = P.indices + cs;
  const float* const cdat = P.data + cs;
  const uint32_t* const prow = P.indices + ps;
  const float* const pdat = P.data + ps;

  // ---- phase 1: membership (node2vec)
  uint32_t kp = B2W_NONE;
  if (!EXTEND && has_prev) {
    uint32_t word0 = 0;
    bool in_regs = false;
    membership_bitmap<32>(T, crow, d, prow, pdeg, prev, bm, kp, word0, in_regs);
    if (in_regs) { if (lane == 0) bm[0] = word0; __syncwarp(); }
  }

  // ---- phase 2: stream the weights, stage w, one partial sum per chunk
  const uint32_t kp2 = 31 - __clz(pdeg | 1u);                                 // floor(log2 pdeg) (first step: pdeg = 0, unused)
  float thr_cur = 0.f;
  if (EXTEND && has_prev) thr_cur = __ldg(P.thr + cur);
  for (uint32_t c = 0; c < nchunks; ++c) {
    const uint32_t k = (c << 5) + lane;
    const bool valid = k < d;
    const float wt = valid ? __ldg(cdat + k) : 0.f;
    float w = wt;
    if (has_prev) {
      if (!EXTEND) {
        const uint32_t bits = bm[c];
        // (invalid lanes keep w = 0: a zero numerator would take the f64 division's slow-path subroutine)
        if (valid) {
          if (k == kp) w = div_by(wt, P.p, P.invp_f, P.p_pow2);                // return bias (sparse_rw.py:87)
          else if (!((bits >> lane) & 1u)) w = div_by(wt, P.q, P.invq_f, P.q_pow2);   // out bias (:86)
        }
      } else {
        const uint32_t x = valid ? __ldg(crow + k) : B2W_NONE;
        bool common;
        const uint32_t pos = lower_bound_eq<false>(prow, pdeg, x, kp2, common);
        if (valid) {
          if (x == prev) {
            w = div_by(wt, P.p, P.invp_f, P.p_pow2);                           // (:126)
          } else {
            bool out = true;
            float t = 0.f;
            if (common) {
              const float wp = __ldg(pdat + pos);
              const float th = __ldg(P.thr + x);
              if (wp >= th) out = false; else t = __fdiv_rn(wp, th);           // (:273-276)
            }
            if (out) {
              double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), (double)t));   // (:119)
              if (wt < thr_cur) alpha = P.supp;                                // (:122-124)
              w = (float)__dmul_rn((double)wt, alpha);                         // (:125)
            }
          }
        }
      }
    }
    if (valid) wbuf[k] = w;
  }
  if (lane < 4) {                                                               // zero pad for float4 replay reads
    const uint32_t k = d + lane;
    if (k < ((d + 3) & ~3u)) wbuf[k] = 0.f;
  }
  __syncwarp();

  // ---- lane-contiguous partial sums: lane l owns elements [l * mseg, (l + 1) * mseg) of the staged row
  // (mseg odd: the strided shared-memory reads are bank-conflict free), one FADD per element and a
  // single warp scan per step instead of a reduction per chunk.
  const uint32_t mseg = nchunks | 1u;
  const uint32_t seg_lo = min(d, (uint32_t)lane * mseg), seg_hi = min(d, seg_lo + mseg);
  float seg = 0.f;
  for (uint32_t i = seg_lo; i < seg_hi; ++i) seg = __fadd_rn(seg, wbuf[i]);
  const float incl_l = warp_incl_scan_f32(seg, lane);
  const float total = __shfl_sync(B2W_FULL, incl_l, 31);

  // ---- filter in the un-normalised domain: compare prefix sums with u * total
  uint32_t choice = d;                       // default: every bound below u -> the reference's choice == deg
  bool replay = (P.flags & B2W_FLAG_FORCE_EXACT_REPLAY) || !(total > 0.f) || !(total < 3.0e38f) || d > (1u << 20);
  if (!replay) {
    const double uT = u * (double)total;
    const double EC = 1.01 * 5.9604644775390625e-08;
    const double ebase = (double)(d + 2 * mseg + 48);
    // lane level: first lane whose segment end possibly reaches u
    const double Al = (double)incl_l;
    const uint32_t bal = __ballot_sync(B2W_FULL, seg_hi > seg_lo && (fma(Al, EC * (ebase + (double)seg_hi), Al) >= uT));
    if (bal) {
      const int fl = __ffs(bal) - 1;
      const uint32_t flo = __shfl_sync(B2W_FULL, seg_lo, fl), fhi = __shfl_sync(B2W_FULL, seg_hi, fl);
      float carry = __shfl_sync(B2W_FULL, __fadd_rn(incl_l, -seg), fl);   // prefix before the segment (within bound)
      bool decided = false;
      for (uint32_t k0 = flo; k0 < fhi && !decided; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool valid = k < fhi;
        const float w = valid ? wbuf[k] : 0.f;
        const float sc = __fadd_rn(carry, warp_incl_scan_f32(w, lane));
        const double A = (double)sc;
        const double E = A * (EC * (ebase + (double)k));
        const uint32_t bp = __ballot_sync(B2W_FULL, valid && (A + E >= uT));
        if (bp) {
          const int f = __ffs(bp) - 1;
          const bool sure = __shfl_sync(B2W_FULL, (A - E >= uT) ? 1 : 0, f) != 0;
          if (sure) choice = k0 + f; else replay = true;
          decided = true;
        }
        carry = __shfl_sync(B2W_FULL, sc, 31);
      }
      if (!decided) replay = true;                                              // rounding at the segment boundary
    }
  }
  if (replay) {
    // ---- exact replay of the reference's recurrences
    const uint32_t n4 = (d + 3) >> 2;
    const float4* w4 = reinterpret_cast<const float4*>(wbuf);
    float s = 0.f;
    for (uint32_t i = 0; i < n4; ++i) {
      const float4 v = w4[i];
      s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
    }
    // probabilities in place, one fdiv per lane per chunk (a 0/0 on an all-zero row gives the reference's NaN)
    for (uint32_t c = 0; c < nchunks; ++c) {
      const uint32_t k = (c << 5) + lane;
      if (k < d) wbuf[k] = __fdiv_rn(wbuf[k], s);
    }
    __syncwarp();
    float cdf = 0.f;
    choice = d;
    for (uint32_t i = 0; i < n4; ++i) {
      const float4 v = w4[i];
      const float c0 = __fadd_rn(cdf, v.x), c1 = __fadd_rn(c0, v.y), c2 = __fadd_rn(c1, v.z), c3 = __fadd_rn(c2, v.w);
      if (!((double)c3 < u)) {                                                  // cdf is non-decreasing (or NaN)
        choice = 4 * i + (!((double)c0 < u) ? 0 : !((double)c1 < u) ? 1 : !((double)c2 < u) ? 2 : 3);
        break;
      }
      cdf = c3;
    }
    if (choice > d) choice = d;
    st.replays++;
    st.seqsums++;
  }
  if (choice == d) st.overflow++;
  __syncwarp();
  return choice;
}

template <bool EXTEND>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4) walk_sparse_warp_kernel(const WalkParams P) {
  __shared__ WarpBuf sbuf[WARPS_PER_CTA];
  const Tile<32> T;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t warp_gid = blockIdx.x * WARPS_PER_CTA + wib;
  // global scratch row of this warp: [w: work_stride floats][bitmap: work_stride/32 words]
  float* const gw = P.work + (size_t)warp_gid * (P.work_stride + P.work_stride / 32);
  uint32_t* const gbm = reinterpret_cast<uint32_t*>(gw + P.work_stride);
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
      const bool small = deg + 4 <= CAP;
      const uint32_t choice = otf_choice_warp<EXTEND>(P, T, cur, cs, deg, j > 1, prev, ps, pdeg, u,
                                                      small ? sbuf[wib].w : gw,
                                                      small ? sbuf[wib].bm : gbm, st);
      const uint32_t nxt = __ldg(P.indices + cs + choice);            // unchecked, as pecanpy.py:559
      if (lane == (j & 31)) myval = nxt;
      if ((j & 31) == 31) {                                           // entries [j-31, j] complete: flush
        __stcs(out + ((j & ~31u) + lane), myval);      // streaming: the 3 GB walk matrix must not evict the graph from L2
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
      if (e < L + 2) __stcs(out + e, v);
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
  size_t stride = (g->max_degree + 4 > (uint32_t)CAP) ? (((size_t)g->max_degree + 4 + 127) & ~(size_t)127) : 0;
  return 256 + (size_t)b2w_sparse_warp_total_warps(g) * (stride + stride / 32) * sizeof(float);
}

int b2w_launch_sparse_warp(const b2w_graph* g, const WalkParams& P_in, cudaStream_t s) {
  WalkParams P = P_in;
  char* base = reinterpret_cast<char*>(P_in.work);
  P.counter = reinterpret_cast<unsigned long long*>(base);
  P.work = reinterpret_cast<float*>(base + 256);
  P.work_stride = (g->max_degree + 4 > (uint32_t)CAP) ? (uint32_t)(((size_t)g->max_degree + 4 + 127) & ~(size_t)127) : 0;
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
