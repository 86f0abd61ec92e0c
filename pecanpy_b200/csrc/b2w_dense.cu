// b2w_dense.cu -- DenseOTF: one CTA per walker, persistent CTAs, dynamic row queue.
//
// Each step streams row(cur) of the f64 adjacency matrix (plus the bool mask of cur and either
// the mask of prev (node2vec) or row(prev) + thresholds (node2vec+)) exactly ONCE with
// coalesced loads.  The reference compresses the row by the mask, normalises by a sequential
// f64 sum, takes a sequential f64 cumsum and bisects it (pecanpy.py:596-612,
// rw/dense_rw.py:34-118).  Here the row is cut into 128-column super tiles; pass 1 leaves one
// f64 partial sum per super tile in shared memory, a warp scan over those finds the super tile
// in which the prefix crosses u * total, and pass 2 re-reads only that super tile (L1/L2 hot) to
// pick the column.  As in the sparse kernel the parallel sums are a FILTER: every quantity is
// within (2N + 256) * 2^-53 (relative) of the reference's sequentially rounded value, so the
// choice is proven unless u falls inside that window around a boundary (probability ~1e-8 per
// step), in which case one lane replays the reference's recurrences exactly.
//
// choice == number-of-neighbours (cdf[-1] < u) indexes past the end of a temporary in the
// reference (pecanpy.py:610-612, undefined behaviour); engine and oracle clamp to the last
// neighbour.
#include "b2w_common.cuh"

namespace {

constexpr int DWARPS = 8;
constexpr int DTHREADS = DWARPS * 32;
constexpr int SUPER = 128;   // columns per super tile (4 per lane)

template <bool EXTEND>
__device__ __forceinline__ double dense_weight(const WalkParams& P, const double* __restrict__ rcur,
                                               const double* __restrict__ rprev, const uint8_t* __restrict__ nzprev,
                                               bool has_prev, uint32_t prev, uint32_t k, double thr_cur) {
  double w = __ldg(rcur + k);
  if (!has_prev) return w;
  if (k == prev) return __ddiv_rn(w, P.p);                            // dense_rw.py:67 / :113
  if (!EXTEND) {
    if (!__ldg(nzprev + k)) w = __ddiv_rn(w, P.q);                    // :63-66
  } else {
    double wp = __ldg(rprev + k);
    double th = (double)__ldg(P.thr + k);
    if (wp < th) {                                                    // :94
      double t = __ddiv_rn(wp, th);                                   // :101
      double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), t));   // :106
      if (w < thr_cur) alpha = P.supp;                                // :109-111
      w = __dmul_rn(w, alpha);                                        // :112
    }
  }
  return w;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(B2W_FULL, v, o));
  return v;
}

__device__ __forceinline__ double warp_incl_scan_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(B2W_FULL, v, o);
    if (lane >= o) v = __dadd_rn(v, t);
  }
  return v;
}

template <bool EXTEND>
__global__ void __launch_bounds__(DTHREADS) walk_dense_kernel(const WalkParams P, const uint32_t n_super) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tile_sum = reinterpret_cast<double*>(smem_raw);             // [n_super] inclusive prefix after the scan
  uint32_t* out_row = reinterpret_cast<uint32_t*>(tile_sum + n_super);   // [L + 2]
  __shared__ unsigned long long s_row;
  __shared__ uint32_t s_cnt[DWARPS], s_last[DWARPS];
  __shared__ uint32_t s_choice;   // chosen column, or 0xFFFFFFFF: replay needed, 0xFFFFFFFE: overflow

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t N = P.n, L = P.L;
  const double EPS = (2.0 * N + 256.0) * 1.01 * 1.1102230246251565e-16;   // (2N+256) * 1.01 * 2^-53
  uint32_t st_steps = 0, st_replays = 0, st_overflow = 0;

  for (;;) {
    if (threadIdx.x == 0) s_row = atomicAdd(P.counter, 1ull);
    __syncthreads();
    const unsigned long long i = s_row;
    if (i >= P.n_rows) break;
    for (uint32_t e = threadIdx.x; e < L + 2; e += DTHREADS) out_row[e] = 0u;
    uint32_t cur = __ldg(P.start + i), prev = 0;
    uint32_t eff = L + 1;
    __syncthreads();
    if (threadIdx.x == 0) out_row[0] = cur;

    for (uint32_t j = 1; j <= L; ++j) {
      const bool has_prev = j > 1;
      const double* rcur = P.dense + (size_t)cur * N;
      const uint8_t* nzcur = P.nonzero + (size_t)cur * N;
      const double* rprev = P.dense + (size_t)prev * N;
      const uint8_t* nzprev = P.nonzero + (size_t)prev * N;
      double thr_cur = 0.0;
      if (EXTEND && has_prev) thr_cur = (double)__ldg(P.thr + cur);

      // ---- pass 1: one partial sum per 128-column super tile
      uint32_t cnt = 0, last = 0;
      for (uint32_t t = warp; t < n_super; t += DWARPS) {
        const uint32_t base = t * SUPER + lane;
        uint8_t nz[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint32_t k = base + 32 * r;
          nz[r] = (k < N) ? __ldg(nzcur + k) : 0;
        }
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint32_t k = base + 32 * r;
          if (nz[r]) {
            acc = __dadd_rn(acc, dense_weight<EXTEND>(P, rcur, rprev, nzprev, has_prev, prev, k, thr_cur));
            ++cnt; last = k;
          }
        }
        acc = warp_sum_f64(acc);
        if (lane == 0) tile_sum[t] = acc;
      }
      cnt = __reduce_add_sync(B2W_FULL, cnt);
      last = __reduce_max_sync(B2W_FULL, last);
      if (lane == 0) { s_cnt[warp] = cnt; s_last[warp] = last; }
      __syncthreads();
      uint32_t tot_cnt = 0, last_col = 0;
#pragma unroll
      for (int w = 0; w < DWARPS; ++w) { tot_cnt += s_cnt[w]; last_col = max(last_col, s_last[w]); }
      if (tot_cnt == 0) { eff = j; break; }                           // has_nbrs(cur) == False (dense_rw.py:25-30)

      const double u = step_uniform(P, i, j);

      // ---- warp 0: inclusive scan of the super-tile sums, locate the crossing, resolve the column
      if (warp == 0) {
        const uint32_t per = (n_super + 31) / 32;
        const uint32_t b = lane * per, e = min(n_super, b + per);
        double loc = 0.0;
        for (uint32_t t = b; t < e; ++t) loc = __dadd_rn(loc, tile_sum[t]);
        double incl = warp_incl_scan_f64(loc, lane);
        double run = __shfl_up_sync(B2W_FULL, incl, 1);   // exclusive prefix of this lane's segment
        if (lane == 0) run = 0.0;
        const double total = __shfl_sync(B2W_FULL, incl, 31);
        for (uint32_t t = b; t < e; ++t) { run = __dadd_rn(run, tile_sum[t]); tile_sum[t] = run; }
        __syncwarp();
        // first super tile whose inclusive prefix possibly reaches u
        uint32_t found = 0xFFFFFFFFu;
        for (uint32_t t0 = 0; t0 < n_super && found == 0xFFFFFFFFu; t0 += 32) {
          uint32_t t = t0 + lane;
          bool poss = false;
          if (t < n_super) {
            double A = __ddiv_rn(tile_sum[t], total);
            poss = (A + EPS * A) >= u;
          }
          uint32_t bal = __ballot_sync(B2W_FULL, poss);
          if (bal) found = t0 + __ffs(bal) - 1;
        }
        uint32_t result;
        if (P.flags & B2W_FLAG_FORCE_EXACT_REPLAY) {
          result = 0xFFFFFFFFu;
        } else if (found == 0xFFFFFFFFu) {
          result = 0xFFFFFFFEu;                                       // every prefix certainly < u: overflow
        } else {
          // ---- pass 2: the located super tile, column order = (r, lane)
          const double excl = found ? tile_sum[found - 1] : 0.0;
          const uint32_t base = found * SUPER + lane;
          double carry = excl;
          result = 0xFFFFFFFFu;
          bool done = false;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            uint32_t k = base + 32 * r;
            bool nzk = (k < N) && __ldg(nzcur + k);
            double w = nzk ? dense_weight<EXTEND>(P, rcur, rprev, nzprev, has_prev, prev, k, thr_cur) : 0.0;
            double sc = warp_incl_scan_f64(w, lane);
            double A = __ddiv_rn(__dadd_rn(carry, sc), total);
            double E = EPS * A;
            uint32_t bp = __ballot_sync(B2W_FULL, nzk && (A + E >= u));
            uint32_t bd = __ballot_sync(B2W_FULL, nzk && (A - E >= u));
            if (!done && bp) {
              int fp = __ffs(bp) - 1;
              if (bd && (__ffs(bd) - 1) == fp) result = found * SUPER + 32 * r + fp;
              done = true;   // first possible column seen: either proven or ambiguous
            }
            carry = __dadd_rn(carry, __shfl_sync(B2W_FULL, sc, 31));
          }
        }
        if (lane == 0) s_choice = result;
      }
      __syncthreads();
      uint32_t nxt = s_choice;
      if (nxt == 0xFFFFFFFFu) {
        // ---- exact replay by one lane (dense_rw.py:69-70, pecanpy.py:608-612 verbatim order)
        if (threadIdx.x == 0) {
          double S = 0.0;
          for (uint32_t k = 0; k < N; ++k)
            if (__ldg(nzcur + k)) S = __dadd_rn(S, dense_weight<EXTEND>(P, rcur, rprev, nzprev, has_prev, prev, k, thr_cur));
          double cdf = 0.0;
          uint32_t pick = 0xFFFFFFFEu;
          for (uint32_t k = 0; k < N; ++k) {
            if (!__ldg(nzcur + k)) continue;
            double w = dense_weight<EXTEND>(P, rcur, rprev, nzprev, has_prev, prev, k, thr_cur);
            cdf = __dadd_rn(cdf, __ddiv_rn(w, S));
            if (!(cdf < u)) { pick = k; break; }
          }
          s_choice = pick;
          ++st_replays;
        }
        __syncthreads();
        nxt = s_choice;
      }
      if (nxt == 0xFFFFFFFEu) { nxt = last_col; if (threadIdx.x == 0) ++st_overflow; }
      if (threadIdx.x == 0) { out_row[j] = nxt; ++st_steps; }
      prev = cur;
      cur = nxt;
      __syncthreads();   // s_choice / tile_sum / s_cnt are rewritten by the next step
    }
    __syncthreads();
    if (threadIdx.x == 0) out_row[L + 1] = eff;
    __syncthreads();
    uint32_t* out = P.out + i * P.ld_out;
    for (uint32_t e = threadIdx.x; e < L + 2; e += DTHREADS) out[e] = out_row[e];
    __syncthreads();
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS) && threadIdx.x == 0) {
    if (st_steps) atomicAdd((unsigned long long*)&P.stats->steps, (unsigned long long)st_steps);
    if (st_replays) atomicAdd((unsigned long long*)&P.stats->exact_replays, (unsigned long long)st_replays);
    if (st_overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)st_overflow);
  }
}

}  // namespace

int b2w_launch_dense(const b2w_graph* g, int extend, const WalkParams& P_in, cudaStream_t s) {
  WalkParams P = P_in;
  P.counter = reinterpret_cast<unsigned long long*>(P_in.work);
  B2W_CUDA(cudaMemsetAsync(P.counter, 0, 8, s));
  const uint32_t n_super = (g->n + SUPER - 1) / SUPER;
  size_t smem = (size_t)n_super * sizeof(double) + ((size_t)P.L + 2) * sizeof(uint32_t);
  if (smem > 200 * 1024) { b2w_set_error("dense walk: row too wide / walk too long for shared memory (%zu bytes)", smem); return B2W_ERR_UNSUPPORTED; }
  int per_sm = 0;
  if (extend) {
    B2W_CUDA(cudaFuncSetAttribute(walk_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2W_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_dense_kernel<true>, DTHREADS, smem));
  } else {
    B2W_CUDA(cudaFuncSetAttribute(walk_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2W_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_dense_kernel<false>, DTHREADS, smem));
  }
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = (uint64_t)per_sm * g->num_sms;
  if (grid > P.n_rows) grid = P.n_rows ? P.n_rows : 1;
  if (extend)
    walk_dense_kernel<true><<<(unsigned)grid, DTHREADS, smem, s>>>(P, n_super);
  else
    walk_dense_kernel<false><<<(unsigned)grid, DTHREADS, smem, s>>>(P, n_super);
  return b2w_cuda_fail(cudaGetLastError(), "walk_dense_kernel launch");
}
