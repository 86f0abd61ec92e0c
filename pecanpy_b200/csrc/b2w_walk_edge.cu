// b2w_walk_edge.cu -- SparseOTF node2vec on UNWEIGHTED graphs through the per-edge index: one lane per walker.
//
// On an unweighted graph every biased weight of a step is one of  w_in = 1, w_out = f32(1/q), w_ret = f32(1/p)
// (rw/sparse_rw.py:86-87), so the transition distribution of a step taken from `cur` after arriving over the edge
// e = (prev -> cur) is fully described by  deg(cur),  the position kp of prev in row(cur)  and the sorted positions
// p_0 < ... < p_{m-1} of the common neighbours -- which is what the edge index holds for e (b2w_edge_index.cu).
// In integer units of the common weight grid g (host-verified: no f32 partial sum can round, b2w_walk_uw.cu) the
// exact un-normalised prefix is
//     W(k) = (k + 1) a_out + #{p_i <= k} (a_in - a_out) + [kp <= k] (a_ret - a_out),        W_d = W(d - 1),
// piecewise linear in k with m + 1 jumps.  The reference's cdf_k = (W(k) / W_d)(1 + t), |t| <= e_k = 1.02 (k + 3) 2^-24
// (the filter of b2w_walk_uw.cu), so with A = u W_d the reference's searchsorted(cdf, u) is the first k with
// W(k) >= ceil(A (1 - e)), proven whenever that k also has W(k) >= ceil(A (1 + e + 2 e^2)).  The first k at or above
// an integer threshold is found in closed form: a bisection over the m jump positions (W at a jump is monotone) and
// one rounded-up division inside the linear segment.  No row is read at all: a step costs one 16-byte record, the
// list words it bisects (m = 0 for most edges of a sparse graph) and one L2-resident indptr entry.  Steps the filter
// cannot prove (~1 %: hub rows) replay the reference's f32 recurrence exactly from the same list (b2w_replay.cuh).
// The step after the reference's unchecked choice == deg read (pecanpy.py:559; ~1e-7 of the steps) did not arrive
// over a stored edge and is evaluated from the rows by the walker's whole warp, in the reference's own order
// (b2w_offedge.cuh; one lane alone took ~1 ms on two hub rows and set the tail of every launch).
//
// One lane owns one walker (the PreComp mapping): 32 independent chains of dependent loads per warp and no
// cross-lane traffic; every lane stages its row in shared memory and writes complete 32-byte sectors (b2w_rowout.cuh).
//
// Reference: pecanpy.py:164-210 (_random_walks), :522-561 (SparseOTF.get_move_forward),
//            rw/sparse_rw.py:51-91 (get_normalized_probs), :142-230 (isnotin).
#include <cmath>

#include "b2w_offedge.cuh"
#include "b2w_probs.cuh"
#include "b2w_replay.cuh"
#include "b2w_rowout.cuh"

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t KPF_POS_MASK = 0x3FFFFFFFu;
constexpr uint32_t KPF_NOTFOUND = 0x40000000u;
constexpr uint32_t KPF_HAS_TRI = 0x80000000u;
constexpr int EW_THREADS = 256;
constexpr uint32_t CKP = 128;                                         // checkpoint spacing of the exact cdf (positions)

struct EdgeConsts {
  const uint4* __restrict__ rec;
  const uint32_t* __restrict__ tri;
  const float* __restrict__ ckpt;    // exact reference cdf every CKP positions, per (node, slot of prev), or null
  const uint32_t* __restrict__ ckb;  // per node: offset of its checkpoint block
  int a_in, a_out, a_ret;            // 1 / g, w_out / g, w_ret / g  (exact integers, (max_degree + 1) max(a) < 2^24)
  double inv_in, inv_out;            // 1.0 / a_in, 1.0 / a_out
  float g;                           // the common grid of the three weights (a power of two)
};

// The distribution of one step, in integer units of g.
struct StepDist {
  const uint32_t* __restrict__ lst;  // m sorted jump positions (common neighbours), none equal to kp
  uint32_t m, kp, d;
  int a_o, da, dr;
  double inv;                        // 1.0 / a_o

  // W at position k, given the number of list entries <= k
  __device__ __forceinline__ int W(uint32_t k, uint32_t c_incl) const {
    return (int)(k + 1) * a_o + (int)c_incl * da + (kp <= k ? dr : 0);
  }
  // ceil(need / a_o)  (any sign; |need| < 2^26)
  __device__ __forceinline__ int ceil_div(const int need) const {
    int qv = __double2int_rz((double)need * inv);                     // within +-1 of the quotient
    if (qv * a_o < need) ++qv;
    if ((qv - 1) * a_o >= need) --qv;
    return qv;
  }
  // smallest k in [lo, hi] with W(k) >= T, or -1, when no list entry lies inside [lo, hi] and B0 = c da for the c
  // entries before lo.  The jump of kp splits the range in at most two linear pieces; both candidates are formed
  // without branching (lanes of a warp disagree on where kp falls).
  __device__ __forceinline__ int range(const int lo, const int hi, const int B0, const int T) const {
    const bool kp_in = hi >= 0 && kp <= (uint32_t)hi;                 // (NONE never is)
    const int hiA = kp_in ? (int)kp - 1 : hi;                         // an empty range (hi < lo) fails both tests below
    const int kA = max(ceil_div(T - B0) - 1, lo);                     // piece before kp: (k + 1) a_o + B0 >= T
    const int kB = max(ceil_div(T - B0 - dr) - 1, max(lo, (int)kp));  // piece from kp on
    return kA <= hiA ? kA : ((kp_in && kB <= hi) ? kB : -1);
  }
  // first k with W(k) >= T (d - 1 when T exceeds the total); Wk receives W(k)
  __device__ __forceinline__ uint32_t first_at_least(const int T, int& Wk) const {
    uint32_t lo = 0, hi = m;                                          // first jump i with W(p_i) >= T
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      const uint32_t pm = __ldg(lst + mid);
      if (W(pm, mid + 1) >= T) hi = mid; else lo = mid + 1;
    }
    const uint32_t i = lo;
    const int klo = i ? (int)__ldg(lst + i - 1) + 1 : 0;
    const uint32_t pi = i < m ? __ldg(lst + i) : d;
    const int r = range(klo, (int)pi - 1, (int)i * da, T);
    const uint32_t k = r >= 0 ? (uint32_t)r : (i < m ? pi : d - 1);
    const uint32_t c = (r >= 0 || i >= m) ? i : i + 1;
    Wk = W(k, c);
    return k;
  }
};

// Exact replay of the reference's f32 cumsum from the jump list (the list form of replay_exact in b2w_walk_uw.cu),
// starting after element k0 - 1 with the running sum cdf0 (k0 = 0, cdf0 = 0: from the beginning; otherwise a
// checkpoint of the exact cdf, which must be < u).
__device__ __noinline__ uint32_t replay_list(const uint32_t* __restrict__ lst, const uint32_t m, const uint32_t d,
                                             const uint32_t kp, const float fa, const float fo, const float fp,
                                             const double u, const uint32_t k0, const float cdf0) {
  const float ub = upper_float(u);                                    // cdf < u  <=>  cdf < ub
  float cdf = cdf0;
  uint32_t k = k0, choice = d, ti = 0;
  if (k0) {                                                           // first jump at or after k0
    uint32_t lo = 0, hi = m;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(lst + mid) < k0) lo = mid + 1; else hi = mid; }
    ti = lo;
  }
  bool kp_left = kp != NONE && kp >= k0;
  uint32_t nextp = ti < m ? __ldg(lst + ti) : NONE;
  for (;;) {
    uint32_t pos;
    bool is_kp;
    if (kp_left && kp < nextp) { pos = kp; is_kp = true; }
    else if (ti < m) { pos = nextp; is_kp = false; }
    else break;
    if (advance_run(cdf, k, pos - k, fo, ub, choice)) return choice;
    cdf = __fadd_rn(cdf, is_kp ? fp : fa);                            // the special element at `pos`
    if (cdf >= ub) return pos;
    k = pos + 1;
    if (is_kp) kp_left = false;
    else { ++ti; nextp = ti < m ? __ldg(lst + ti) : NONE; }
  }
  if (advance_run(cdf, k, d - k, fo, ub, choice)) return choice;
  return d;                                                           // cdf[-1] < u: the reference's overflow
}

// One-lane form of the same step (graphs without long rows, where it is short): prev is not joined to cur by the edge the
// walker holds, so the step is evaluated from the rows in the reference's own order -- sequential merge, sequential
// f32 sum and cumsum (rw/sparse_rw.py:51-91, pecanpy.py:556-557) -- with the three unweighted biases.
__device__ __noinline__ uint32_t offedge_step(const uint32_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                              const float w_out, const float w_ret, const uint32_t cur,
                                              const uint32_t prev, const double u) {
  const uint32_t cs = indptr[cur], d = indptr[cur + 1] - cs;
  const uint32_t ps = indptr[prev], pd = indptr[prev + 1] - ps;
  float sum = 0.f, cdf = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t i2 = 0;
    for (uint32_t k = 0; k < d; ++k) {
      const uint32_t x = indices[cs + k];
      float w = w_ret;
      if (x != prev) {
        while (i2 < pd && indices[ps + i2] < x) ++i2;
        w = (i2 < pd && indices[ps + i2] == x) ? 1.0f : w_out;
      }
      if (pass == 0) sum = __fadd_rn(sum, w);
      else {
        cdf = __fadd_rn(cdf, __fdiv_rn(w, sum));
        if (!((double)cdf < u)) return k;
      }
    }
  }
  return d;
}

__device__ __forceinline__ uint32_t edge_step(const EdgeConsts& C, const uint32_t flags, const uint32_t cur,
                                              const uint32_t d, const bool has_prev, const uint32_t kpf,
                                              const uint32_t toff, const double u, uint32_t& st_replays) {
  StepDist D;
  D.d = d;
  D.m = 0; D.lst = C.tri; D.kp = NONE;
  if (has_prev) {
    if (!(kpf & KPF_NOTFOUND)) D.kp = kpf & KPF_POS_MASK;
    if (kpf & KPF_HAS_TRI) { D.m = __ldg(C.tri + toff); D.lst = C.tri + toff + 1; }
    D.a_o = C.a_out; D.inv = C.inv_out;
  } else {
    D.a_o = C.a_in; D.inv = C.inv_in;                                 // first step: every weight is 1
  }
  D.da = C.a_in - D.a_o;
  D.dr = C.a_ret - D.a_o;
  const int Wd = (int)d * D.a_o + (int)D.m * D.da + (D.kp != NONE ? D.dr : 0);
  const double EC = 1.02 * 5.9604644775390625e-08;                    // 1.02 * 2^-24
  const double A = u * (double)Wd;
  uint32_t k_first = 0;                                               // no element before it can be the answer
  if (d <= 160000u) {
    // An upper bound k_hi of the answer from the most conservative "sure" threshold and W(k) >= (k + 1) a_o - slack
    // (slack = the downward jumps), then both thresholds with e taken at k_hi (e >= e_k for every k <= k_hi).
    const double e_row = EC * (double)(d + 2);
    const double t_hi = A * (1.0 + e_row + 2.0 * e_row * e_row + 2.9e-14);
    const int slack = (D.da < 0 ? (int)D.m * -D.da : 0) + (D.dr < 0 && D.kp != NONE ? -D.dr : 0);
    const double k_up = (t_hi + (double)slack) * D.inv + 1.0;          // > ceil((t_hi + slack) / a_o) - 1
    const uint32_t k_hi = k_up < (double)(d - 1) ? (uint32_t)k_up : d - 1;
    const double e = EC * (double)(k_hi + 3);
    const int T_poss = (int)ceil(A * (1.0 - e - 2.9e-14));
    const double t_sure = ceil(A * (1.0 + e + 2.0 * e * e + 2.9e-14));
    int Wk;
    const uint32_t k = D.first_at_least(T_poss, Wk);                  // T_poss <= W_d: k exists, k <= k_hi
    if ((double)Wk >= t_sure && !(flags & B2W_FLAG_FORCE_EXACT_REPLAY)) return k;
    k_first = k;                                                      // cdf_j < u is proven for every j < k
  }
  // ---- exact replay.  The reference's probabilities: three exact f32 quotients by S = W_d g (rw/sparse_rw.py:89)
  const float S = __fmul_rn(__int2float_rn(Wd), C.g);                 // exact: W_d < 2^24, g a power of two
  const float fa = __fdiv_rn(__fmul_rn(__int2float_rn(C.a_in), C.g), S);
  const float fo = has_prev ? __fdiv_rn(__fmul_rn(__int2float_rn(C.a_out), C.g), S) : fa;
  const float fp = __fdiv_rn(__fmul_rn(__int2float_rn(C.a_ret), C.g), S);
  uint32_t k0 = 0;
  float cdf0 = 0.f;
  if (C.ckpt != nullptr && D.kp != NONE && k_first >= CKP) {
    // start from the checkpoint of the exact cdf at or before k_first (rows of >= CKP slots, return edge present)
    const uint32_t nck = d / CKP;
    uint32_t j = k_first / CKP;
    if (j > nck) j = nck;
    cdf0 = __ldg(C.ckpt + ((size_t)__ldg(C.ckb + cur) + (size_t)D.kp * nck + (j - 1)));
    k0 = j * CKP;
  }
  ++st_replays;
  return replay_list(D.lst, D.m, d, D.kp, fa, fo, fp, u, k0, cdf0);
}

// ---- checkpoints of the exact cdf: per stored edge (prev -> cur) with deg(cur) >= CKP whose reverse edge exists, the
// reference's f32 cdf after elements CKP - 1, 2 CKP - 1, ... of row(cur), laid out [node][slot of prev][j].  One lane
// per edge runs the same binade-jumping emulation as the replay, so the build is O(#jumps + #binades + deg / CKP).
__global__ void __launch_bounds__(EW_THREADS) edge_ckpt_kernel(const uint64_t nnz, const EdgeConsts C, float* __restrict__ out_all) {
  for (uint64_t e = blockIdx.x * (uint64_t)EW_THREADS + threadIdx.x; e < nnz; e += (uint64_t)gridDim.x * EW_THREADS) {
    const uint4 r = __ldg(C.rec + e);
    const uint32_t d = r.w;
    if (d < CKP || (r.y & KPF_NOTFOUND)) continue;
    const uint32_t kp = r.y & KPF_POS_MASK;
    uint32_t m = 0;
    const uint32_t* lst = C.tri;
    if (r.y & KPF_HAS_TRI) { m = __ldg(C.tri + r.z); lst = C.tri + r.z + 1; }
    const int Wd = (int)d * C.a_out + (int)m * (C.a_in - C.a_out) + (C.a_ret - C.a_out);
    const float S = __fmul_rn(__int2float_rn(Wd), C.g);
    const float fa = __fdiv_rn(__fmul_rn(__int2float_rn(C.a_in), C.g), S);
    const float fo = __fdiv_rn(__fmul_rn(__int2float_rn(C.a_out), C.g), S);
    const float fp = __fdiv_rn(__fmul_rn(__int2float_rn(C.a_ret), C.g), S);
    const uint32_t nck = d / CKP, last = nck * CKP;
    float* const out = out_all + ((size_t)__ldg(C.ckb + r.x) + (size_t)kp * nck);
    const float INF = __uint_as_float(0x7F800000u);
    float cdf = 0.f;
    uint32_t k = 0, j = 0, ti = 0, dummy = 0;
    bool kp_left = true;
    uint32_t nextp = m ? __ldg(lst) : NONE;
    while (k < last) {
      const uint32_t B = (j + 1) * CKP;                               // cdf after element B - 1 is checkpoint j
      const uint32_t ev = min(kp_left ? kp : NONE, nextp);
      const uint32_t stop = min(ev, B);
      if (stop > k) advance_run(cdf, k, stop - k, fo, INF, dummy);    // never stops early: k == stop afterwards
      if (k == B) { out[j++] = cdf; continue; }
      const bool is_kp = kp_left && ev == kp;                         // the special element at k == ev < B
      cdf = __fadd_rn(cdf, is_kp ? fp : fa);
      ++k;
      if (is_kp) kp_left = false;
      else { ++ti; nextp = ti < m ? __ldg(lst + ti) : NONE; }
    }
    if (j < nck) out[j] = cdf;                                        // k == last reached through a special element
  }
}

// per node: floats of its checkpoint block = deg * (deg / CKP) (0 below CKP slots), to be prefix-summed
__global__ void __launch_bounds__(EW_THREADS) edge_ckpt_count_kernel(const uint32_t n, const uint32_t* __restrict__ indptr,
                                                                     uint32_t* __restrict__ ckb, unsigned int* __restrict__ too_big) {
  for (uint64_t v = blockIdx.x * (uint64_t)EW_THREADS + threadIdx.x; v <= n; v += (uint64_t)gridDim.x * EW_THREADS) {
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

// The lanes of a warp stay converged at loop level (a finished or dead-ended walker idles instead of leaving), so that
// the rare step without an edge -- after the reference's unchecked choice == deg read -- can be evaluated by the
// whole warp (b2w_offedge.cuh) instead of stalling it behind one lane.
// COOP = false (graphs whose longest row is short): plain per-lane loops, the off-edge step by its own lane -- keeping
// the warp converged costs ~12 % there (ER config #2: 92 vs 81 G steps/s) and buys nothing.
// MIRROR = true (multi-GPU jobs, b2w_walk_mirrored; always with COOP): every row is also stored into the same rows of
// the other GPUs' matrices over NVLink while the walk goes on -- the all-gather fused into the walk.  The rows leave
// through WarpRowTile (b2w_rowout.cuh): coalesced warp stores of up to 31 words per row every 24 steps.
template <int MINB, bool COOP, bool MIRROR>
__global__ void __launch_bounds__(EW_THREADS, MINB) walk_uw_edge_kernel(const WalkParams P, const EdgeConsts C) {
  static_assert(COOP || !MIRROR, "mirrored stores need the converged loops");
  __shared__ uint32_t s_stage[(MIRROR ? 33 : 8) * EW_THREADS];
  const uint32_t L = P.L;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t st_steps = 0, st_replays = 0, st_overflow = 0;
#define B2W_PUSH(jj, vv) do { if (MIRROR) tilew.put(jj, vv); else row.push(jj, vv); } while (0)
#define B2W_FINISH(cc) do { if (!MIRROR) row.finish(cc); } while (0)
  if (!COOP) {
    WarpRowTile tilew;                                                // (unused here: MIRROR implies COOP)
    const float w_out = __fmul_rn(__int2float_rn(C.a_out), C.g), w_ret = __fmul_rn(__int2float_rn(C.a_ret), C.g);
    for (uint64_t i = blockIdx.x * (uint64_t)EW_THREADS + threadIdx.x; i < P.n_rows; i += (uint64_t)gridDim.x * EW_THREADS) {
      RowWriter<EW_THREADS> row;
      row.begin(P.out + i * P.ld_out, s_stage);
      uint32_t cur = __ldg(P.start + i), prev = 0;
      uint32_t cs = __ldg(P.indptr + cur);
      uint32_t d = __ldg(P.indptr + cur + 1) - cs;
      uint32_t kpf = 0, toff = 0, eff = L + 1;
      bool edge_ok = true;                                            // the walker arrived over a stored edge
      B2W_PUSH(0, cur);
      uint32_t j = 1;
      for (; j <= L; ++j) {
        if (d == 0) { eff = j; break; }                               // pecanpy.py:194-196, 204-206
        const double u = step_uniform(P, i, j);
        uint32_t choice;
        if (edge_ok) choice = edge_step(C, P.flags, cur, d, j > 1, kpf, toff, u, st_replays);
        else choice = offedge_step(P.indptr, P.indices, w_out, w_ret, cur, prev, u);   // after an unchecked choice == deg read
        if (choice == d) ++st_overflow;
        edge_ok = choice < d;
        const uint4 r = __ldg(C.rec + (cs + choice));                 // [cs + d] is the next row's first edge: pecanpy.py:559
        prev = cur;
        cur = r.x; kpf = r.y; toff = r.z; d = r.w;
        B2W_PUSH(j, cur);
        cs = __ldg(P.indptr + cur);
      }
      st_steps += eff - 1;
      for (uint32_t z = j; z <= L; ++z) B2W_PUSH(z, 0u);              // zero tail (np.zeros, pecanpy.py:182)
      B2W_PUSH(L + 1, eff);
      B2W_FINISH(L + 2);
    }
  }
  for (uint64_t i = blockIdx.x * (uint64_t)EW_THREADS + threadIdx.x; COOP && __any_sync(B2W_FULL, i < P.n_rows);
       i += (uint64_t)gridDim.x * EW_THREADS) {
    const bool alive = i < P.n_rows;
    RowWriter<EW_THREADS> row;
    WarpRowTile tilew;
    if (MIRROR) tilew.begin(s_stage + (threadIdx.x >> 5) * (33 * 32), i - lane, (uint32_t)min((uint64_t)32, P.n_rows - (i - lane)));
    uint32_t since = 1;                                               // words put since the last flush
    uint32_t cur = 0, prev = 0, cs = 0, d = 0;
    if (alive) {
      if (!MIRROR) row.begin(P.out + i * P.ld_out, s_stage);
      cur = __ldg(P.start + i);
      cs = __ldg(P.indptr + cur);
      d = __ldg(P.indptr + cur + 1) - cs;
      B2W_PUSH(0, cur);
    }
    uint32_t kpf = 0, toff = 0, eff = L + 1;
    bool walking = alive;
    bool edge_ok = true;                                              // the walker arrived over a stored edge
    for (uint32_t j = 1; j <= L; ++j) {
      if (walking && d == 0) { eff = j; walking = false; }            // pecanpy.py:194-196, 204-206
      uint32_t choice = 0;
      double u = 0.0;
      if (walking) {
        u = step_uniform(P, i, j);
        if (edge_ok) choice = edge_step(C, P.flags, cur, d, j > 1, kpf, toff, u, st_replays);
      }
      __syncwarp();
      uint32_t off = __ballot_sync(B2W_FULL, walking && !edge_ok);    // after an unchecked choice == deg read (rare)
      while (off) {
        const int src = __ffs(off) - 1;
        off &= off - 1;
        const uint32_t c = offedge_uw_warp(P.indptr, P.indices, C.a_in, C.a_out, C.a_ret, C.g,
                                           __shfl_sync(B2W_FULL, cur, src), __shfl_sync(B2W_FULL, prev, src),
                                           __shfl_sync(B2W_FULL, u, src));
        if (lane == (uint32_t)src) choice = c;
      }
      if (walking) {
        if (choice == d) ++st_overflow;
        edge_ok = choice < d;
        const uint4 r = __ldg(C.rec + (cs + choice));                 // [cs + d] is the next row's first edge: pecanpy.py:559
        prev = cur;
        cur = r.x; kpf = r.y; toff = r.z; d = r.w;
        cs = __ldg(P.indptr + cur);
      }
      if (alive) B2W_PUSH(j, walking ? cur : 0u);                     // zero tail after a dead end (np.zeros, pecanpy.py:182)
      if (MIRROR && ++since == MIRROR_PERIOD) { tilew.flush(P, j + 1, false); since = 0; }
    }
    if (alive) {
      st_steps += eff - 1;
      B2W_PUSH(L + 1, eff);
      B2W_FINISH(L + 2);
    }
    if (MIRROR) tilew.flush(P, L + 2, true);
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS)) {
    for (int o = 16; o; o >>= 1) {
      st_steps += __shfl_xor_sync(B2W_FULL, st_steps, o);
      st_replays += __shfl_xor_sync(B2W_FULL, st_replays, o);
      st_overflow += __shfl_xor_sync(B2W_FULL, st_overflow, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (st_steps) atomicAdd((unsigned long long*)&P.stats->steps, (unsigned long long)st_steps);
      if (st_replays) atomicAdd((unsigned long long*)&P.stats->exact_replays, (unsigned long long)st_replays);
      if (st_overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)st_overflow);
    }
  }
}
#undef B2W_PUSH
#undef B2W_FINISH

// ---------------------------------------------------------------------------------------------------
// PreComp through the per-edge index.  The reference's step (pecanpy.py:427-438) bisects prev in row(cur) to find
// the alias table of the (prev, cur) pair: that position is a property of the edge the walker arrived over and sits
// in its record (kpf: the insertion point, exactly what np.searchsorted returns, found or not).  A step is then
//     record (L2: 16 B x nnz)  ->  alias_indptr[cur], indptr[cur] (L2)  ->  one {q, j} entry (HBM)  ->  next record
// instead of log2(deg) dependent probes + two table sectors.  Only the step after the first step's unchecked
// choice == deg read (pecanpy.py:424, 559) has no edge to lean on and bisects like the reference.
template <int MINB>
__global__ void __launch_bounds__(EW_THREADS, MINB) walk_precomp_edge_kernel(const WalkParams P, const uint4* __restrict__ rec) {
  __shared__ uint32_t s_stage[8 * EW_THREADS];
  const uint32_t L = P.L;
  unsigned long long st_steps = 0, st_overflow = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)EW_THREADS + threadIdx.x; i < P.n_rows; i += (uint64_t)gridDim.x * EW_THREADS) {
    RowWriter<EW_THREADS> row;
    row.begin(P.out + i * P.ld_out, s_stage);
    const uint64_t grow = P.row0 + i;
    uint32_t cur = __ldg(P.start + i), prev = 0;
    uint32_t cs = __ldg(P.indptr + cur);
    uint32_t d = __ldg(P.indptr + cur + 1) - cs;
    uint32_t lo = 0, eff = L + 1;
    bool edge_ok = true;
    row.push(0, cur);
    uint32_t j = 1;
    for (; j <= L; ++j) {
      if (d == 0) { eff = j; break; }                                 // pecanpy.py:194-196, 204-206
      StepRng rng;
      rng.begin(P.key0, P.key1, grow, j);
      uint32_t choice;
      if (j == 1) {
        // first step: cumsum / searchsorted on the NON-extended first-order probabilities (pecanpy.py:412-424)
        choice = otf_choice_seq<false>(P, cur, false, 0, rng.uniform());
        if (choice == d) { ++st_overflow; edge_ok = false; }
      } else {
        if (!edge_ok) {                                               // np.searchsorted(indices[start:end], prev) (:429)
          uint32_t a = 0, b = d;
          while (a < b) { const uint32_t mid = (a + b) >> 1; if (__ldg(P.indices + cs + mid) < prev) a = mid + 1; else b = mid; }
          lo = a;
          edge_ok = true;
        }
        const uint64_t off = __ldg(P.alias_indptr + cur) + (uint64_t)d * lo;   // (:433-434)
        const uint32_t kk = rng.randint(d);                           // alias_draw (:668-677)
        const double u = rng.uniform();
        if (P.alias_qj) {
          const uint2 e = __ldg(P.alias_qj + off + kk);
          choice = (u < (double)__uint_as_float(e.x)) ? kk : e.y;
        } else {
          choice = (u < (double)__ldg(P.alias_q + off + kk)) ? kk : __ldg(P.alias_j + off + kk);
        }
      }
      const uint4 r = __ldg(rec + (cs + choice));
      prev = cur;
      cur = r.x; lo = r.y & KPF_POS_MASK; d = r.w;
      row.push(j, cur);
      cs = __ldg(P.indptr + cur);
    }
    st_steps += eff - 1;
    for (uint32_t z = j; z <= L; ++z) row.push(z, 0u);                // zero tail (np.zeros, pecanpy.py:182)
    row.push(L + 1, eff);
    row.finish(L + 2);
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS)) {
    for (int o = 16; o; o >>= 1) {
      st_steps += __shfl_xor_sync(B2W_FULL, st_steps, o);
      st_overflow += __shfl_xor_sync(B2W_FULL, st_overflow, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&P.stats->steps, st_steps);
      if (st_overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, st_overflow);
    }
  }
}

}  // namespace

int b2w_launch_precomp_edge(const b2w_graph* g, const WalkParams& P, cudaStream_t s) {
  if (!(g->flags & B2W_GRAPH_HAS_EDGE_INDEX)) { b2w_set_error("walk_precomp_edge_kernel: no edge index attached"); return B2W_ERR_INVALID; }
  const uint64_t want = (P.n_rows + EW_THREADS - 1) / EW_THREADS;
  const uint64_t cap = (uint64_t)g->num_sms * 64;
  unsigned blocks = (unsigned)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  const uint4* rec = reinterpret_cast<const uint4*>(g->edge_rec);
  const int mb = (int)((P.flags >> 16) & 0xF);                        // tuning: resident CTAs per SM (0 = default)
  // The alias tables (hundreds of MB, one random 64-byte line per step) stream through L2 and push out the edge
  // records (16 B x nnz), which every step reads too.  B2W_FLAG_L2_PERSIST pins the records with an access-policy
  // window (persisting hits, everything else streaming).
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(EW_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr; cfg.numAttrs = 0;
  if (P.flags & B2W_FLAG_L2_PERSIST) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, g->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, g->device);
    size_t bytes = ((size_t)g->nnz + 1) * 16;
    if (max_persist > 0 && max_window > 0) {
      if (bytes > (size_t)max_window) bytes = (size_t)max_window;
      const size_t carve = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(g->edge_rec);
      attr[0].val.accessPolicyWindow.num_bytes = bytes;
      attr[0].val.accessPolicyWindow.hitRatio = (float)((double)carve / (double)bytes);
      attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cfg.numAttrs = 1;
    }
    cudaGetLastError();
  }
  // measured on BASELINE config #4 (G steps/s): 5 CTAs/SM (48 registers) 31.9, 6 CTAs/SM (40 registers) 29.7
  if (mb == 8) cudaLaunchKernelEx(&cfg, walk_precomp_edge_kernel<8>, P, rec);
  else if (mb == 4) cudaLaunchKernelEx(&cfg, walk_precomp_edge_kernel<4>, P, rec);
  else if (mb == 6) cudaLaunchKernelEx(&cfg, walk_precomp_edge_kernel<6>, P, rec);
  else cudaLaunchKernelEx(&cfg, walk_precomp_edge_kernel<5>, P, rec);
  return b2w_cuda_fail(cudaGetLastError(), "walk_precomp_edge_kernel launch");
}

int b2w_launch_uw_edge(const b2w_graph* g, const WalkParams& P, cudaStream_t s) {
  int gexp = 0;
  if (!b2w_uw_grid(g, P.p, P.q, &gexp) || !(g->flags & B2W_GRAPH_HAS_EDGE_INDEX)) {
    b2w_set_error("walk_uw_edge_kernel: graph / p / q not eligible or no edge index attached");
    return B2W_ERR_INVALID;
  }
  EdgeConsts C;
  C.rec = reinterpret_cast<const uint4*>(g->edge_rec);
  C.tri = g->edge_tri;
  const bool ck = g->edge_ckpt && g->edge_ck_p == P.p && g->edge_ck_q == P.q && !(P.flags & B2W_FLAG_NO_CKPT);
  C.ckpt = ck ? g->edge_ckpt : nullptr;
  C.ckb = ck ? g->edge_ckb : nullptr;
  C.g = ldexpf(1.0f, gexp);
  C.a_in = (int)ldexp(1.0, -gexp);                                    // exact integers by construction of g
  C.a_out = (int)ldexp((double)(float)(1.0 / P.q), -gexp);
  C.a_ret = (int)ldexp((double)(float)(1.0 / P.p), -gexp);
  C.inv_in = 1.0 / (double)C.a_in;
  C.inv_out = 1.0 / (double)C.a_out;
  const uint64_t want = (P.n_rows + EW_THREADS - 1) / EW_THREADS;
  const uint64_t cap = (uint64_t)g->num_sms * 64;                     // grid-stride beyond a few waves
  unsigned blocks = (unsigned)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  const int mb = (int)((P.flags >> 16) & 0xF);                        // tuning: resident CTAs per SM (0 = default)
  // long rows: keep the warps converged and evaluate the off-edge steps cooperatively (flag bits 24 / 25 force on / off)
  bool coop = g->max_degree > 256u;
  if (P.flags & B2W_FLAG_OFFEDGE_WARP) coop = true;
  if (P.flags & B2W_FLAG_OFFEDGE_LANE) coop = false;
  const bool mirror = P.n_mirrors > 0;
  if (mirror) coop = true;
#define B2W_EDGE_LAUNCH(MB)                                                                     \
  do {                                                                                          \
    if (mirror) {                                                                               \
      walk_uw_edge_kernel<MB, true, true><<<blocks, EW_THREADS, 0, s>>>(P, C);                  \
    } else {                                                                                    \
      if (coop) walk_uw_edge_kernel<MB, true, false><<<blocks, EW_THREADS, 0, s>>>(P, C);       \
      else walk_uw_edge_kernel<MB, false, false><<<blocks, EW_THREADS, 0, s>>>(P, C);           \
    }                                                                                           \
  } while (0)
  if (mb == 6) B2W_EDGE_LAUNCH(6);
  else if (mb == 4) B2W_EDGE_LAUNCH(4);
  else B2W_EDGE_LAUNCH(5);
#undef B2W_EDGE_LAUNCH
  return b2w_cuda_fail(cudaGetLastError(), "walk_uw_edge_kernel launch");
}

// ---------------------------------------------------------------- checkpoints of the exact cdf (unweighted SparseOTF)
#include "b2w_scan.cuh"

extern "C" size_t b2w_edge_ckpt_work_bytes(const b2w_graph* g) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) return 0;
  return b2w_scan::work_bytes((uint64_t)g->n + 1) + 256;
}

extern "C" int b2w_edge_ckpt_prepare(const b2w_graph* g, uint32_t* d_ckb, void* d_work, size_t work_bytes,
                                     uint64_t* h_ckpt_floats, void* stream) {
  if (!g || !(g->flags & B2W_GRAPH_HAS_EDGE_INDEX)) { b2w_set_error("b2w_edge_ckpt_prepare: attach the edge index first"); return B2W_ERR_INVALID; }
  if (!d_ckb || !h_ckpt_floats || !d_work || work_bytes < b2w_edge_ckpt_work_bytes(g)) { b2w_set_error("b2w_edge_ckpt_prepare: bad argument / scratch too small"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_work) + 256);
  unsigned int* flag = reinterpret_cast<unsigned int*>(d_work);
  B2W_CUDA(cudaMemsetAsync(flag, 0, 4, s));
  edge_ckpt_count_kernel<<<(unsigned)g->num_sms * 4, EW_THREADS, 0, s>>>(g->n, g->indptr, d_ckb, flag);
  B2W_CUDA(cudaGetLastError());
  unsigned long long total = 0;
  B2W_CUDA(b2w_scan::exclusive_scan((uint64_t)g->n + 1, d_ckb, 1, sums, &total, s));
  unsigned int h_flag = 0;
  B2W_CUDA(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, s));
  B2W_CUDA(cudaStreamSynchronize(s));
  *h_ckpt_floats = total;
  if (h_flag || total >= 0xFFFFFFFFull) { b2w_set_error("b2w_edge_ckpt_prepare: %llu checkpoints do not fit 32-bit offsets", total); return B2W_ERR_UNSUPPORTED; }
  return B2W_OK;
}

extern "C" int b2w_edge_ckpt_finish(b2w_graph* g, double p, double q, const uint32_t* d_ckb, float* d_ckpt,
                                    uint64_t ckpt_floats, void* stream) {
  if (!g || !(g->flags & B2W_GRAPH_HAS_EDGE_INDEX)) { b2w_set_error("b2w_edge_ckpt_finish: attach the edge index first"); return B2W_ERR_INVALID; }
  if (!d_ckb || (ckpt_floats && !d_ckpt)) { b2w_set_error("b2w_edge_ckpt_finish: null array"); return B2W_ERR_INVALID; }
  int gexp = 0;
  if (!b2w_uw_grid(g, p, q, &gexp)) { b2w_set_error("b2w_edge_ckpt_finish: graph / p / q not eligible for the unweighted kernels"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  EdgeConsts C{};
  C.rec = reinterpret_cast<const uint4*>(g->edge_rec);
  C.tri = g->edge_tri;
  C.ckb = d_ckb;
  C.g = ldexpf(1.0f, gexp);
  C.a_in = (int)ldexp(1.0, -gexp);
  C.a_out = (int)ldexp((double)(float)(1.0 / q), -gexp);
  C.a_ret = (int)ldexp((double)(float)(1.0 / p), -gexp);
  if (ckpt_floats) {
    edge_ckpt_kernel<<<(unsigned)g->num_sms * 16, EW_THREADS, 0, s>>>(g->nnz, C, d_ckpt);
    B2W_CUDA(cudaGetLastError());
  }
  B2W_CUDA(cudaStreamSynchronize(s));
  g->edge_ckpt = ckpt_floats ? d_ckpt : nullptr; g->edge_ckb = d_ckb; g->edge_ck_p = p; g->edge_ck_q = q;
  return B2W_OK;
}

extern "C" int b2w_graph_clear_edge_ckpt(b2w_graph* g) {
  if (!g) { b2w_set_error("clear_edge_ckpt: null graph"); return B2W_ERR_INVALID; }
  g->edge_ckpt = nullptr; g->edge_ckb = nullptr;
  return B2W_OK;
}
