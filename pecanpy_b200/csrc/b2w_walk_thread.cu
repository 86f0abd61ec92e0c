// b2w_walk_thread.cu -- lane-per-walker kernels.
//
// PreComp (alias draw), the two first-order modes, and a lane-per-walker SparseOTF variant.
// One thread owns one walker for its whole life: per step the work is a handful of dependent
// gathers (binary search of prev in cur's row, alias_q/alias_j, indices), so the natural
// mapping is one lane per walker with 32 independent gather chains in flight per warp.
// Reference: pecanpy.py:164-210 (_random_walks), :409-438 (PreComp.move_forward),
// :304-307 (FirstOrderUnweighted), :327-332 (PreCompFirstOrder), :668-677 (alias_draw).
#include "b2w_probs.cuh"
#include "b2w_rowout.cuh"

namespace {

// alias_draw (pecanpy.py:668-677): kk = randint(k); rand() < q[kk] ? kk : j[kk]
__device__ __forceinline__ uint32_t alias_draw(const uint32_t* __restrict__ j, const float* __restrict__ q,
                                               const uint2* __restrict__ qj, uint64_t off, uint32_t k, StepRng& rng) {
  uint32_t kk = rng.randint(k);
  double u = rng.uniform();
  if (qj) {                                                          // packed table: one 8-byte entry per draw
    const uint2 e = __ldg(qj + off + kk);
    return (u < (double)__uint_as_float(e.x)) ? kk : e.y;
  }
  float qv = q[off + kk];
  return (u < (double)qv) ? kk : j[off + kk];
}

// MINB = minimum resident CTAs per SM (register cap): the PreComp step is a chain of ~6 dependent gathers, so
// what it wants is warps in flight, not registers.
template <int MODE, bool EXTEND, int MINB>
__global__ void __launch_bounds__(256, MINB) walk_thread_kernel(const WalkParams P) {
  __shared__ uint32_t s_stage[8 * 256];                                // rows leave as complete 32-byte sectors (b2w_rowout.cuh)
  const uint32_t L = P.L;
  uint64_t steps = 0, overflow = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < P.n_rows;
       i += (uint64_t)gridDim.x * blockDim.x) {
    RowWriter<256> out;
    out.begin(P.out + i * P.ld_out, s_stage);
    const uint64_t row = P.row0 + i;
    uint32_t cur = P.start[i];
    uint32_t prev = 0;
    uint32_t eff = L + 1;
    out.push(0, cur);
    uint32_t j = 1;
    for (; j <= L; ++j) {
      uint32_t cs = P.indptr[cur];
      uint32_t deg = P.indptr[cur + 1] - cs;
      if (deg == 0) { eff = j; break; }                                // pecanpy.py:194-196,204-206
      uint32_t choice;
      if (MODE == B2W_MODE_SPARSE_OTF) {
        double u = step_uniform(P, i, j);
        choice = otf_choice_seq<EXTEND>(P, cur, j > 1, prev, u);
        if (choice == deg) ++overflow;
      } else {
        StepRng rng;
        rng.begin(P.key0, P.key1, row, j);
        if (MODE == B2W_MODE_PRECOMP) {
          if (j == 1) {
            // first step: cumsum/searchsorted on the NON-extended first-order probs (:412-424)
            choice = otf_choice_seq<false>(P, cur, false, 0, rng.uniform());
            if (choice == deg) ++overflow;
          } else {
            // np.searchsorted(indices[start:end], prev) (:429); table at alias_indptr + deg*idx (:433-434)
            uint32_t lo = 0, hi = deg;
            while (lo < hi) {
              uint32_t mid = (lo + hi) >> 1;
              if (P.indices[cs + mid] < prev) lo = mid + 1; else hi = mid;
            }
            uint64_t off = P.alias_indptr[cur] + (uint64_t)deg * lo;
            choice = alias_draw(P.alias_j, P.alias_q, P.alias_qj, off, deg, rng);
          }
        } else if (MODE == B2W_MODE_FIRST_ORDER_UNWEIGHTED) {
          choice = rng.randint(deg);                                   // randint(start, end) (:306-307)
        } else {                                                       // PRECOMP_FIRST_ORDER (:329-332)
          choice = alias_draw(P.alias_j, P.alias_q, nullptr, cs, deg, rng);
        }
      }
      uint32_t nxt = P.indices[cs + choice];
      out.push(j, nxt);
      prev = cur;
      cur = nxt;
      ++steps;
    }
    for (uint32_t z = j; z <= L; ++z) out.push(z, 0u);                 // zero tail (np.zeros, :182)
    out.push(L + 1, eff);
    out.finish(L + 2);
  }
  if (P.stats && !(P.flags & B2W_FLAG_NO_FILTER_STATS)) {
    // one atomic per warp
    for (int o = 16; o; o >>= 1) {
      steps += __shfl_xor_sync(B2W_FULL, steps, o);
      overflow += __shfl_xor_sync(B2W_FULL, overflow, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&P.stats->steps, (unsigned long long)steps);
      if (overflow) atomicAdd((unsigned long long*)&P.stats->overflow_choices, (unsigned long long)overflow);
    }
  }
}

template <int MODE, bool EXTEND, int MINB = 1>
int launch(const b2w_graph* g, const WalkParams& P, cudaStream_t s) {
  int threads = 256;
  uint64_t want = (P.n_rows + threads - 1) / threads;
  uint64_t cap = (uint64_t)g->num_sms * 8 * 4;   // grid-stride beyond a few waves
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  walk_thread_kernel<MODE, EXTEND, MINB><<<blocks, threads, 0, s>>>(P);
  return b2w_cuda_fail(cudaGetLastError(), "walk_thread_kernel launch");
}

}  // namespace

int b2w_launch_thread_walk(const b2w_graph* g, int mode, int extend, const WalkParams& P, cudaStream_t s) {
  switch (mode) {
    case B2W_MODE_SPARSE_OTF:
      return extend ? launch<B2W_MODE_SPARSE_OTF, true>(g, P, s) : launch<B2W_MODE_SPARSE_OTF, false>(g, P, s);
    case B2W_MODE_PRECOMP: {
      // measured on BASELINE config #4 (G steps/s): 64 regs / 4 CTAs 19.0, 48 / 5 20.2, 40 / 6 17.8, 32 / 8 17.6
      const int mb = (int)((P.flags >> 16) & 0xF);                     // tuning: resident CTAs per SM (0 = default)
      if (mb == 4) return launch<B2W_MODE_PRECOMP, false, 1>(g, P, s);
      if (mb == 6) return launch<B2W_MODE_PRECOMP, false, 6>(g, P, s);
      if (mb == 8) return launch<B2W_MODE_PRECOMP, false, 8>(g, P, s);
      return launch<B2W_MODE_PRECOMP, false, 5>(g, P, s);
    }
    case B2W_MODE_FIRST_ORDER_UNWEIGHTED: return launch<B2W_MODE_FIRST_ORDER_UNWEIGHTED, false>(g, P, s);
    case B2W_MODE_PRECOMP_FIRST_ORDER: return launch<B2W_MODE_PRECOMP_FIRST_ORDER, false>(g, P, s);
  }
  b2w_set_error("b2w_launch_thread_walk: unsupported mode %d", mode);
  return B2W_ERR_UNSUPPORTED;
}
