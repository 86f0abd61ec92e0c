// b2w_dense.cu -- DenseOTF: one CTA per walker, persistent CTAs, dynamic row queue.
//
// Each step streams row(cur) of the f64 adjacency matrix (plus the bool mask of cur and either
// the mask of prev (node2vec) or row(prev) + thresholds (node2vec+)) exactly ONCE with
// coalesced loads.  The reference compresses the row by the mask, normalises by a sequential
// f64 sum, takes a sequential f64 cumsum and bisects it (pecanpy.py:596-612,
// rw/dense_rw.py:34-118).  Here the row is cut into 128-column super tiles; pass 1 leaves one
// f64 partial sum per super tile in shared memory, a warp scan over those finds the super tile
// in which the prefix crosses u * total, and pass 2 re-reads only that super tile (L1/L2 hot) to
// pick the column.  Each lane owns 4 adjacent columns of a super tile: one 32-bit mask load and two
// 16-byte loads per f64 row (1 KiB contiguous per warp instruction).  As in the sparse kernel the parallel sums are a FILTER: every quantity is
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

// One lane's share of a 128-column super tile: 4 ADJACENT columns k0 .. k0+3 (k0 = 128 t + 4 lane), so the
// mask is one 32-bit load and each f64 row contributes two 16-byte loads per lane (1 KiB contiguous per warp
// instruction).  VEC requires N % 4 == 0 (then every row base is 32-byte aligned); otherwise scalar loads.
struct LaneTile {
  double w[4];        // biased, un-normalised weights (0 for non-neighbours)
  uint32_t nzmask;    // bit r set <=> column k0 + r is a neighbour of cur
};

template <bool EXTEND>
__device__ __forceinline__ double dense_weight(const WalkParams& P, double w, const bool has_prev, const bool isprev,
                                               const bool nzp, const double wp, const float th, const double thr_cur) {
  if (!has_prev) return w;
  if (isprev) return __ddiv_rn(w, P.p);                               // dense_rw.py:67 / :113
  if (!EXTEND) {
    if (!nzp) w = P.q_pow2 ? __dmul_rn(w, (double)P.invq_f) : __ddiv_rn(w, P.q);   // :63-66 (exact for 2^k)
  } else {
    const double thd = (double)th;
    if (wp < thd) {                                                   // :94
      // :101.  0 / thd == +0 exactly (70 % of the columns: non-neighbours of prev); a zero numerator would
      // send the whole warp through the IEEE division's slow-path subroutine, so those lanes divide thd / thd
      // (the asm makes the substituted numerator opaque, otherwise the optimiser folds it back to wp / thd)
      double num = (wp == 0.0) ? thd : wp;
      asm volatile("" : "+d"(num));
      double t = __ddiv_rn(num, thd);
      if (wp == 0.0) t = 0.0;
      double alpha = __dadd_rn(P.invq, __dmul_rn(__dsub_rn(1.0, P.invq), t));   // :106
      if (w < thr_cur) alpha = P.supp;                                // :109-111
      w = __dmul_rn(w, alpha);                                        // :112
    }
  }
  return w;
}

template <bool EXTEND, bool VEC>
__device__ __forceinline__ void load_lane_tile(const WalkParams& P, const double* __restrict__ rcur,
                                               const uint8_t* __restrict__ nzcur, const double* __restrict__ rprev,
                                               const uint8_t* __restrict__ nzprev, const bool has_prev,
                                               const uint32_t prev, const uint32_t k0, const uint32_t N,
                                               const double thr_cur, LaneTile& out) {
  uint32_t m4 = 0, mp4 = 0;
  double c[4] = {0.0, 0.0, 0.0, 0.0}, pv[4] = {0.0, 0.0, 0.0, 0.0};
  float th[4] = {0.f, 0.f, 0.f, 0.f};
  out.nzmask = 0;
  out.w[0] = out.w[1] = out.w[2] = out.w[3] = 0.0;
  if (VEC) {
    // all loads of the lane are issued together (no dependence on the mask): one round trip per tile
    if (k0 < N) {
      m4 = __ldg(reinterpret_cast<const uint32_t*>(nzcur + k0));
      const double2 a = __ldg(reinterpret_cast<const double2*>(rcur + k0));
      const double2 b = __ldg(reinterpret_cast<const double2*>(rcur + k0 + 2));
      c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
      if (has_prev) {
        if (!EXTEND) {
          mp4 = __ldg(reinterpret_cast<const uint32_t*>(nzprev + k0));
        } else {
          const double2 e = __ldg(reinterpret_cast<const double2*>(rprev + k0));
          const double2 f = __ldg(reinterpret_cast<const double2*>(rprev + k0 + 2));
          pv[0] = e.x; pv[1] = e.y; pv[2] = f.x; pv[3] = f.y;
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(P.thr + k0));
          th[0] = t4.x; th[1] = t4.y; th[2] = t4.z; th[3] = t4.w;
        }
      }
    }
    if (m4 == 0) return;
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (k0 + r < N && __ldg(nzcur + k0 + r)) m4 |= 0xFFu << (8 * r);
    if (m4 == 0) return;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if ((m4 >> (8 * r)) & 0xFFu) {
        c[r] = __ldg(rcur + k0 + r);
        if (has_prev) {
          if (!EXTEND) { if (__ldg(nzprev + k0 + r)) mp4 |= 0xFFu << (8 * r); }
          else { pv[r] = __ldg(rprev + k0 + r); th[r] = __ldg(P.thr + k0 + r); }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if ((m4 >> (8 * r)) & 0xFFu) {
      out.nzmask |= 1u << r;
      out.w[r] = dense_weight<EXTEND>(P, c[r], has_prev, k0 + r == prev, ((mp4 >> (8 * r)) & 0xFFu) != 0, pv[r], th[r],
                                      thr_cur);
    }
  }
}

// ---------------------------------------------------------------- TMA (cp.async.bulk) staging
// Pass 1 can stage the rows through shared memory with 1-D bulk tensor copies (SASS: UBLKCP) instead
// of per-lane LDG: one elected thread issues, per 1024-column tile, the copies of row(cur), and of
// row(prev) + thresholds (node2vec+) or mask(prev) (node2vec), plus mask(cur); an mbarrier with an
// expected-transaction byte count signals arrival; a 3-stage ring keeps two tiles in flight while the
// warps reduce the third.  Requires N % 16 == 0 (every copy is a multiple of 16 B at a 16 B-aligned address).
constexpr int TMA_TILE = 1024;      // columns per stage
constexpr int TMA_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();          // never hang the device: a lost copy becomes a launch failure
  }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool EXTEND>
struct TmaStage {
  static constexpr int A_OFF = 0;                                     // row(cur)   f64[1024]
  static constexpr int B_OFF = A_OFF + TMA_TILE * 8;                  // row(prev)  f64[1024]   (node2vec+)
  static constexpr int M_OFF = EXTEND ? B_OFF + TMA_TILE * 8 : A_OFF + TMA_TILE * 8;   // mask(cur) u8[1024]
  static constexpr int X_OFF = M_OFF + TMA_TILE;                      // thr f32[1024] (n2v+) | mask(prev) u8[1024]
  static constexpr int BYTES = X_OFF + (EXTEND ? TMA_TILE * 4 : TMA_TILE);
};

// one lane's 4 adjacent columns, read from a staged tile (same arithmetic as load_lane_tile)
template <bool EXTEND>
__device__ __forceinline__ void smem_lane_tile(const WalkParams& P, const unsigned char* __restrict__ st,
                                               const bool has_prev, const uint32_t prev, const uint32_t k0,
                                               const uint32_t col, const uint32_t N, const double thr_cur, LaneTile& out) {
  out.nzmask = 0;
  out.w[0] = out.w[1] = out.w[2] = out.w[3] = 0.0;
  if (k0 >= N) return;
  const uint32_t m4 = *reinterpret_cast<const uint32_t*>(st + TmaStage<EXTEND>::M_OFF + col);
  if (m4 == 0) return;
  const double2 a = *reinterpret_cast<const double2*>(st + TmaStage<EXTEND>::A_OFF + 8 * col);
  const double2 b = *reinterpret_cast<const double2*>(st + TmaStage<EXTEND>::A_OFF + 8 * col + 16);
  const double c[4] = {a.x, a.y, b.x, b.y};
  double pv[4] = {0.0, 0.0, 0.0, 0.0};
  float th[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t mp4 = 0;
  if (has_prev) {
    if (!EXTEND) {
      mp4 = *reinterpret_cast<const uint32_t*>(st + TmaStage<EXTEND>::X_OFF + col);
    } else {
      const double2 e = *reinterpret_cast<const double2*>(st + TmaStage<EXTEND>::B_OFF + 8 * col);
      const double2 f = *reinterpret_cast<const double2*>(st + TmaStage<EXTEND>::B_OFF + 8 * col + 16);
      pv[0] = e.x; pv[1] = e.y; pv[2] = f.x; pv[3] = f.y;
      const float4 t4 = *reinterpret_cast<const float4*>(st + TmaStage<EXTEND>::X_OFF + 4 * col);
      th[0] = t4.x; th[1] = t4.y; th[2] = t4.z; th[3] = t4.w;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if ((m4 >> (8 * r)) & 0xFFu) {
      out.nzmask |= 1u << r;
      out.w[r] = dense_weight<EXTEND>(P, c[r], has_prev, k0 + r == prev, ((mp4 >> (8 * r)) & 0xFFu) != 0, pv[r], th[r],
                                      thr_cur);
    }
  }
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

template <bool EXTEND, int LD>   // LD: 0 scalar loads, 1 vector loads, 2 TMA-staged pass 1 (+ vector loads in pass 2)
__global__ void __launch_bounds__(DTHREADS) walk_dense_kernel(const WalkParams P, const uint32_t n_super,
                                                              const uint32_t ts_cap) {
  constexpr bool VEC = LD >= 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [stages (TMA only)] [tile_sum: ts_cap doubles] [out_row: L + 2]
  unsigned char* const stage_base = smem_raw;
  double* tile_sum = reinterpret_cast<double*>(smem_raw + (LD == 2 ? TMA_STAGES * TmaStage<EXTEND>::BYTES : 0));
  uint32_t* out_row = reinterpret_cast<uint32_t*>(tile_sum + ts_cap);   // [L + 2]
  __shared__ __align__(8) uint64_t s_full[TMA_STAGES];
  uint32_t phases = 0;                                                 // parity bit per stage (TMA)
  if (LD == 2) {
    if (threadIdx.x == 0) {
      for (int st = 0; st < TMA_STAGES; ++st) mbar_init(&s_full[st], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
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
      if (LD == 2) {
        const uint32_t n_tiles = (N + TMA_TILE - 1) / TMA_TILE;
        auto issue = [&](uint32_t tile) {                              // elected thread only
          const uint32_t st = tile % TMA_STAGES;
          const uint32_t c0 = tile * TMA_TILE;
          const uint32_t cols = min((uint32_t)TMA_TILE, N - c0);
          unsigned char* dst = stage_base + st * TmaStage<EXTEND>::BYTES;
          uint32_t bytes = cols * 8 + cols;
          if (has_prev) bytes += EXTEND ? cols * 8 + cols * 4 : cols;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this stage are done
          mbar_expect_tx(&s_full[st], bytes);
          tma_load_1d(dst + TmaStage<EXTEND>::A_OFF, rcur + c0, cols * 8, &s_full[st]);
          tma_load_1d(dst + TmaStage<EXTEND>::M_OFF, nzcur + c0, cols, &s_full[st]);
          if (has_prev) {
            if (EXTEND) {
              tma_load_1d(dst + TmaStage<EXTEND>::B_OFF, rprev + c0, cols * 8, &s_full[st]);
              tma_load_1d(dst + TmaStage<EXTEND>::X_OFF, P.thr + c0, cols * 4, &s_full[st]);
            } else {
              tma_load_1d(dst + TmaStage<EXTEND>::X_OFF, nzprev + c0, cols, &s_full[st]);
            }
          }
        };
        if (threadIdx.x == 0)
          for (uint32_t tile = 0; tile < min(n_tiles, (uint32_t)TMA_STAGES); ++tile) issue(tile);
        for (uint32_t tile = 0; tile < n_tiles; ++tile) {
          const uint32_t st = tile % TMA_STAGES;
          mbar_wait(&s_full[st], (phases >> st) & 1u);
          phases ^= 1u << st;
          const uint32_t t = tile * (TMA_TILE / SUPER) + warp;        // this warp's 128-column super tile
          const uint32_t col = warp * SUPER + 4 * lane;               // column inside the staged tile
          const uint32_t k0 = tile * TMA_TILE + col;
          LaneTile lt;
          smem_lane_tile<EXTEND>(P, stage_base + st * TmaStage<EXTEND>::BYTES, has_prev, prev, k0, col, N, thr_cur, lt);
          double acc = __dadd_rn(__dadd_rn(lt.w[0], lt.w[1]), __dadd_rn(lt.w[2], lt.w[3]));
          if (lt.nzmask) { cnt += __popc(lt.nzmask); last = k0 + 31 - __clz(lt.nzmask); }
          acc = warp_sum_f64(acc);
          if (lane == 0 && t < n_super) tile_sum[t] = acc;
          __syncthreads();                                             // every warp is done with stage `st`
          if (threadIdx.x == 0 && tile + TMA_STAGES < n_tiles) issue(tile + TMA_STAGES);
        }
      } else {
        for (uint32_t t = warp; t < n_super; t += DWARPS) {
          const uint32_t k0 = t * SUPER + 4 * lane;
          LaneTile lt;
          load_lane_tile<EXTEND, VEC>(P, rcur, nzcur, rprev, nzprev, has_prev, prev, k0, N, thr_cur, lt);
          double acc = __dadd_rn(__dadd_rn(lt.w[0], lt.w[1]), __dadd_rn(lt.w[2], lt.w[3]));
          if (lt.nzmask) { cnt += __popc(lt.nzmask); last = k0 + 31 - __clz(lt.nzmask); }
          acc = warp_sum_f64(acc);
          if (lane == 0) tile_sum[t] = acc;
        }
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
        const double incl = warp_incl_scan_f64(loc, lane);
        double run = __shfl_up_sync(B2W_FULL, incl, 1);   // exclusive prefix of this lane's segment
        if (lane == 0) run = 0.0;
        const double total = __shfl_sync(B2W_FULL, incl, 31);
        for (uint32_t t = b; t < e; ++t) { run = __dadd_rn(run, tile_sum[t]); tile_sum[t] = run; }
        __syncwarp();
        // first super tile whose inclusive prefix possibly reaches u:  I_t (1 + EPS) >= u * total
        const double uT = u * total;
        uint32_t found = 0xFFFFFFFFu;
        for (uint32_t t0 = 0; t0 < n_super && found == 0xFFFFFFFFu; t0 += 32) {
          const uint32_t t = t0 + lane;
          bool poss = false;
          if (t < n_super) { const double I = tile_sum[t]; poss = fma(I, EPS, I) >= uT; }
          const uint32_t bal = __ballot_sync(B2W_FULL, poss);
          if (bal) found = t0 + __ffs(bal) - 1;
        }
        uint32_t result;
        if (P.flags & B2W_FLAG_FORCE_EXACT_REPLAY) {
          result = 0xFFFFFFFFu;
        } else if (found == 0xFFFFFFFFu) {
          result = 0xFFFFFFFEu;                                       // every prefix certainly < u: overflow
        } else {
          // ---- pass 2: the located super tile, column order = (lane, r)
          const double excl = found ? tile_sum[found - 1] : 0.0;
          const uint32_t k0 = found * SUPER + 4 * lane;
          LaneTile lt;
          load_lane_tile<EXTEND, VEC>(P, rcur, nzcur, rprev, nzprev, has_prev, prev, k0, N, thr_cur, lt);
          const double p0 = lt.w[0], p1 = __dadd_rn(p0, lt.w[1]), p2 = __dadd_rn(p1, lt.w[2]), p3 = __dadd_rn(p2, lt.w[3]);
          const double sc = warp_incl_scan_f64(p3, lane);
          const double base = __dadd_rn(excl, __dsub_rn(sc, p3));     // prefix before this lane's columns
          const double A[4] = {__dadd_rn(base, p0), __dadd_rn(base, p1), __dadd_rn(base, p2), __dadd_rn(base, p3)};
          uint32_t poss = 0, sure = 0;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if ((lt.nzmask >> r) & 1u) {
              const double E = A[r] * EPS;
              if (A[r] + E >= uT) poss |= 1u << r;
              if (A[r] - E >= uT) sure |= 1u << r;
            }
          }
          const uint32_t bal = __ballot_sync(B2W_FULL, poss != 0);
          result = 0xFFFFFFFFu;                                       // default: ambiguous (tile boundary) -> replay
          if (bal) {
            const int fl = __ffs(bal) - 1;
            const uint32_t pm = __shfl_sync(B2W_FULL, poss, fl), sm = __shfl_sync(B2W_FULL, sure, fl);
            const int r = __ffs(pm) - 1;
            if ((sm >> r) & 1u) result = found * SUPER + 4 * fl + r;
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
          for (uint32_t k0 = 0; k0 < N; k0 += 4) {
            LaneTile lt;
            load_lane_tile<EXTEND, false>(P, rcur, nzcur, rprev, nzprev, has_prev, prev, k0, N, thr_cur, lt);
            for (int r = 0; r < 4; ++r) if ((lt.nzmask >> r) & 1u) S = __dadd_rn(S, lt.w[r]);
          }
          double cdf = 0.0;
          uint32_t pick = 0xFFFFFFFEu;
          for (uint32_t k0 = 0; k0 < N && pick == 0xFFFFFFFEu; k0 += 4) {
            LaneTile lt;
            load_lane_tile<EXTEND, false>(P, rcur, nzcur, rprev, nzprev, has_prev, prev, k0, N, thr_cur, lt);
            for (int r = 0; r < 4; ++r) {
              if (!((lt.nzmask >> r) & 1u)) continue;
              cdf = __dadd_rn(cdf, __ddiv_rn(lt.w[r], S));
              if (!(cdf < u)) { pick = k0 + r; break; }
            }
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

template <bool EXTEND, int LD>
int launch_dense(const b2w_graph* g, const WalkParams& P, const uint32_t n_super, cudaStream_t s) {
  const uint32_t ts_cap = (n_super + 15) & ~15u;
  size_t smem = (size_t)ts_cap * sizeof(double) + ((size_t)P.L + 2) * sizeof(uint32_t) +
                (LD == 2 ? (size_t)TMA_STAGES * TmaStage<EXTEND>::BYTES : 0);
  if (smem > 220 * 1024) { b2w_set_error("dense walk: row too wide / walk too long for shared memory (%zu bytes)", smem); return B2W_ERR_UNSUPPORTED; }
  int per_sm = 0;
  B2W_CUDA(cudaFuncSetAttribute(walk_dense_kernel<EXTEND, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2W_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_dense_kernel<EXTEND, LD>, DTHREADS, smem));
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = (uint64_t)per_sm * g->num_sms;
  if (grid > P.n_rows) grid = P.n_rows ? P.n_rows : 1;
  walk_dense_kernel<EXTEND, LD><<<(unsigned)grid, DTHREADS, smem, s>>>(P, n_super, ts_cap);
  return b2w_cuda_fail(cudaGetLastError(), "walk_dense_kernel launch");
}

}  // namespace

int b2w_launch_dense(const b2w_graph* g, int extend, const WalkParams& P_in, cudaStream_t s) {
  WalkParams P = P_in;
  P.counter = reinterpret_cast<unsigned long long*>(P_in.work);
  B2W_CUDA(cudaMemsetAsync(P.counter, 0, 8, s));
  const uint32_t n_super = (g->n + SUPER - 1) / SUPER;
  // vector loads need every row base 16/32-byte aligned: N % 4 == 0 and aligned array bases;
  // TMA staging needs every bulk copy 16-byte aligned and sized: N % 16 == 0
  const bool vec = (g->n % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.dense) & 31) == 0) &&
                   ((reinterpret_cast<uintptr_t>(P.nonzero) & 3) == 0) &&
                   (!extend || (reinterpret_cast<uintptr_t>(P.thr) & 15) == 0);
  const bool tma = vec && (g->n % 16 == 0) && ((reinterpret_cast<uintptr_t>(P.nonzero) & 15) == 0) &&
                   !(P.flags & B2W_FLAG_NO_TMA);
  const int ld = tma ? 2 : (vec ? 1 : 0);
  if (extend) {
    if (ld == 2) return launch_dense<true, 2>(g, P, n_super, s);
    if (ld == 1) return launch_dense<true, 1>(g, P, n_super, s);
    return launch_dense<true, 0>(g, P, n_super, s);
  }
  if (ld == 2) return launch_dense<false, 2>(g, P, n_super, s);
  if (ld == 1) return launch_dense<false, 1>(g, P, n_super, s);
  return launch_dense<false, 0>(g, P, n_super, s);
}
