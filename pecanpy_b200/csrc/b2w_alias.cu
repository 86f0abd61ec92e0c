// b2w_alias.cu -- PreComp alias-table builder.
//
// One lane per (node, neighbour-slot) table.  Vose's construction in the reference
// (pecanpy.py:617-665) is order dependent -- two LIFO stacks, pop-pair loop with f32/f64
// mixed rounding -- so each table is built sequentially by one lane, exactly in the
// reference's order; the Sigma(deg) tables are independent and run in parallel.  Consecutive
// lanes take consecutive slots of the same node, so the lanes of a warp share `deg` (uniform
// trip counts) and broadcast-read the same row of `cur`.
// Reference: pecanpy.py:442-507 (preprocess_transition_probs), :336-361 (first order),
//            rw/sparse_rw.py:51-130 (probabilities), pecanpy.py:617-665 (alias_setup).
#include <cstdlib>

#include "b2w_probs.cuh"

namespace {

// ST = element stride of q[] and j[]: 1 for the reference's two arrays, 2 for the packed {q, j} table
template <int ST>
__device__ __forceinline__ void alias_setup_inplace(uint32_t k, uint32_t* __restrict__ j, float* __restrict__ q,
                                                    uint32_t* __restrict__ stk) {
  // on entry q[kk] holds the normalised probability probs[kk]
  uint32_t sp = 0, lp = 0;   // smaller grows up from stk[0], larger grows down from stk[k-1]
  for (uint32_t kk = 0; kk < k; ++kk) {
    float v = (float)__dmul_rn((double)k, (double)q[kk * ST]);        // q[kk] = k * probs[kk] (:642)
    q[kk * ST] = v;
    j[kk * ST] = 0;
    if (v < 1.0f) stk[sp++] = kk; else stk[k - 1 - lp++] = kk;
  }
  while (sp > 0 && lp > 0) {                                          // (:650-663)
    uint32_t small = stk[--sp];
    uint32_t large = stk[k - 1 - (--lp)];
    j[small * ST] = large;
    float v = (float)__dsub_rn((double)__fadd_rn(q[large * ST], q[small * ST]), 1.0);
    q[large * ST] = v;
    if (v < 1.0f) stk[sp++] = large; else stk[k - 1 - lp++] = large;
  }
}

// PACKED: alias_q points at the packed table (uint2 {q bits, j} per entry, b2w_alias_build_packed), alias_j is unused
template <bool EXTEND, bool FIRST_ORDER, bool PACKED>
__global__ void __launch_bounds__(128) alias_build_kernel(const WalkParams P, const uint64_t* __restrict__ aip,
                                                          uint32_t* __restrict__ alias_j,
                                                          float* __restrict__ alias_q, uint32_t* __restrict__ work,
                                                          uint32_t work_stride, uint64_t n_tables) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t* stk = work + tid * (uint64_t)work_stride;
  for (uint64_t t = tid; t < n_tables; t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t idx, deg, prev = 0;
    uint64_t off;
    if (FIRST_ORDER) {
      idx = (uint32_t)t;
      uint32_t cs = P.indptr[idx];
      deg = P.indptr[idx + 1] - cs;
      off = cs;
      if (deg == 0) continue;
    } else {
      // owner node of CSR slot t: largest idx with indptr[idx] <= t
      uint32_t lo = 0, hi = P.n;
      while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if ((uint64_t)P.indptr[mid] <= t) lo = mid; else hi = mid - 1;
      }
      idx = lo;
      uint32_t cs = P.indptr[idx];
      deg = P.indptr[idx + 1] - cs;
      uint32_t nb = (uint32_t)(t - cs);
      prev = P.indices[t];
      off = aip[idx] + (uint64_t)deg * nb;                            // (:501)
    }
    constexpr int ST = PACKED ? 2 : 1;
    float* q = alias_q + off * ST;
    uint32_t* j = PACKED ? reinterpret_cast<uint32_t*>(q) + 1 : alias_j + off;
    BiasStream<EXTEND> bs(P, idx, !FIRST_ORDER, prev);
    float sum = 0.f;
    for (uint32_t k = 0; k < deg; ++k) {
      float w = bs.weight(k);
      q[k * ST] = w;
      sum = __fadd_rn(sum, w);                                        // sequential f32 sum
    }
    for (uint32_t k = 0; k < deg; ++k) q[k * ST] = __fdiv_rn(q[k * ST], sum);   // rw/sparse_rw.py:89
    alias_setup_inplace<ST>(deg, j, q, stk);
  }
}

// Shared-memory variant (max degree <= ALIAS_SMEM_DEG): Vose's loop is a chain of dependent loads and stores
// on q[], j[] and the two stacks; in global scratch every link costs an L2 round trip.  Here each warp
// builds its 32 tables in shared memory, laid out [k][lane] with a 33-word pitch so that the lane-private,
// data-dependent accesses of the construction AND the transposed (table-contiguous) write-out are both
// bank-conflict free; the finished tables leave with coalesced 128-byte stores.
constexpr int ALIAS_SMEM_DEG = 64;
constexpr int ALIAS_WARPS = 4;
constexpr int ALIAS_PITCH = 33;

template <bool EXTEND, bool FIRST_ORDER, bool PACKED>
__global__ void __launch_bounds__(ALIAS_WARPS * 32) alias_build_smem_kernel(const WalkParams P,
                                                                             const uint64_t* __restrict__ aip,
                                                                             uint32_t* __restrict__ alias_j,
                                                                             float* __restrict__ alias_q,
                                                                             uint64_t n_tables) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PER = ALIAS_SMEM_DEG * ALIAS_PITCH;
  float* sq = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 3 * PER;
  uint32_t* sj = reinterpret_cast<uint32_t*>(sq + PER);
  uint32_t* sk = sj + PER;
  const uint64_t n_batches = (n_tables + 31) / 32;
  for (uint64_t b = blockIdx.x * (uint64_t)ALIAS_WARPS + warp; b < n_batches; b += (uint64_t)gridDim.x * ALIAS_WARPS) {
    const uint64_t t = b * 32 + lane;
    uint32_t idx = 0, deg = 0, prev = 0;
    uint64_t off = 0;
    if (t < n_tables) {
      if (FIRST_ORDER) {
        idx = (uint32_t)t;
        const uint32_t cs = P.indptr[idx];
        deg = P.indptr[idx + 1] - cs;
        off = cs;
      } else {
        uint32_t lo = 0, hi = P.n;
        while (lo < hi) {
          const uint32_t mid = (lo + hi + 1) >> 1;
          if ((uint64_t)P.indptr[mid] <= t) lo = mid; else hi = mid - 1;
        }
        idx = lo;
        const uint32_t cs = P.indptr[idx];
        deg = P.indptr[idx + 1] - cs;
        prev = P.indices[t];
        off = aip[idx] + (uint64_t)deg * (uint32_t)(t - cs);
      }
    }
    // ---- probabilities (sequential f32 sum, rw/sparse_rw.py:89) into sq[k][lane]
    if (deg) {
      BiasStream<EXTEND> bs(P, idx, !FIRST_ORDER, prev);
      float sum = 0.f;
      for (uint32_t k = 0; k < deg; ++k) {
        const float w = bs.weight(k);
        sq[k * ALIAS_PITCH + lane] = w;
        sum = __fadd_rn(sum, w);
      }
      // ---- alias_setup (pecanpy.py:632-665) entirely in shared memory
      uint32_t sp = 0, lp = 0;
      for (uint32_t kk = 0; kk < deg; ++kk) {
        const float v = (float)__dmul_rn((double)deg, (double)__fdiv_rn(sq[kk * ALIAS_PITCH + lane], sum));
        sq[kk * ALIAS_PITCH + lane] = v;
        sj[kk * ALIAS_PITCH + lane] = 0;
        if (v < 1.0f) sk[(sp++) * ALIAS_PITCH + lane] = kk; else sk[(deg - 1 - lp++) * ALIAS_PITCH + lane] = kk;
      }
      while (sp > 0 && lp > 0) {
        const uint32_t small = sk[(--sp) * ALIAS_PITCH + lane];
        const uint32_t large = sk[(deg - 1 - (--lp)) * ALIAS_PITCH + lane];
        sj[small * ALIAS_PITCH + lane] = large;
        const float v = (float)__dsub_rn((double)__fadd_rn(sq[large * ALIAS_PITCH + lane], sq[small * ALIAS_PITCH + lane]), 1.0);
        sq[large * ALIAS_PITCH + lane] = v;
        if (v < 1.0f) sk[(sp++) * ALIAS_PITCH + lane] = large; else sk[(deg - 1 - lp++) * ALIAS_PITCH + lane] = large;
      }
    }
    __syncwarp();
    // ---- coalesced write-out: table s of this batch is contiguous in global memory
    for (int s2 = 0; s2 < 32; ++s2) {
      const uint32_t dg = __shfl_sync(B2W_FULL, deg, s2);
      const uint64_t of = __shfl_sync(B2W_FULL, off, s2);
      for (uint32_t k = lane; k < dg; k += 32) {
        if (PACKED) {
          reinterpret_cast<uint2*>(alias_q)[of + k] = make_uint2(__float_as_uint(sq[k * ALIAS_PITCH + s2]), sj[k * ALIAS_PITCH + s2]);
        } else {
          alias_q[of + k] = sq[k * ALIAS_PITCH + s2];
          alias_j[of + k] = sj[k * ALIAS_PITCH + s2];
        }
      }
    }
    __syncwarp();
  }
}

uint64_t alias_threads(const b2w_graph* g) {
  // resident lanes, bounded so that the stack scratch stays below 256 MiB
  uint64_t lanes = (uint64_t)g->num_sms * 2048;
  uint64_t per = (uint64_t)(g->max_degree ? g->max_degree : 1) * 4;
  uint64_t budget = (256ull << 20) / per;
  if (lanes > budget) lanes = budget;
  uint64_t floor_ = (uint64_t)g->num_sms * 128;
  if (lanes < floor_) lanes = floor_;
  lanes = (lanes / 128) * 128;
  return lanes;
}

}  // namespace

extern "C" size_t b2w_alias_build_work_bytes(const b2w_graph* g) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) return 0;
  return (size_t)(alias_threads(g) * (uint64_t)(g->max_degree ? g->max_degree : 1) * 4);
}

static WalkParams alias_params(const b2w_graph* g, double p, double q, const float* d_thr);
static bool g_alias_force_global = false;   // test hook (B2W_ALIAS_FORCE_GLOBAL=1): exercise the global-scratch builder

static int alias_build_common(const b2w_graph* g, double p, double q, int extend, const float* d_thr,
                              const uint64_t* aip, uint32_t* aj, float* aq, void* d_work, size_t work_bytes,
                              void* stream, bool first_order, bool packed = false) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("alias build: CSR graph handle required"); return B2W_ERR_INVALID; }
  if ((!packed && !aj) || !aq || (!first_order && !aip)) { b2w_set_error("alias build: null output/offset pointer"); return B2W_ERR_INVALID; }
  if (extend && !d_thr) { b2w_set_error("alias build: extend requires noise thresholds"); return B2W_ERR_INVALID; }
  if (!(p > 0.0) || !(q > 0.0)) { b2w_set_error("alias build: p and q must be > 0"); return B2W_ERR_INVALID; }
  size_t need = b2w_alias_build_work_bytes(g);
  if (work_bytes < need || (!d_work && need)) {
    b2w_set_error("alias build: scratch too small (%zu < %zu bytes)", work_bytes, need);
    return B2W_ERR_INVALID;
  }
  B2W_CUDA(cudaSetDevice(g->device));
  { const char* e = getenv("B2W_ALIAS_FORCE_GLOBAL"); g_alias_force_global = e && e[0] == '1'; }
  WalkParams P = alias_params(g, p, q, d_thr);
  uint64_t lanes = alias_threads(g);
  uint64_t n_tables = first_order ? g->n : g->nnz;
  if (n_tables == 0) return B2W_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (g->max_degree <= (uint32_t)ALIAS_SMEM_DEG && !(g_alias_force_global)) {
    const size_t smem = (size_t)ALIAS_WARPS * 3 * ALIAS_SMEM_DEG * ALIAS_PITCH * sizeof(float);
    uint64_t nb = (n_tables + 32 * ALIAS_WARPS - 1) / (32 * ALIAS_WARPS);
    uint64_t cap = (uint64_t)g->num_sms * 16;
    unsigned blocks = (unsigned)(nb < cap ? nb : cap);
#define B2W_ALIAS_SMEM(E, F, K)                                                                                     \
    do {                                                                                                            \
      B2W_CUDA(cudaFuncSetAttribute(alias_build_smem_kernel<E, F, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      alias_build_smem_kernel<E, F, K><<<blocks, ALIAS_WARPS * 32, smem, s>>>(P, aip, aj, aq, n_tables);          \
    } while (0)
    if (first_order) B2W_ALIAS_SMEM(false, true, false);
    else if (extend) { if (packed) B2W_ALIAS_SMEM(true, false, true); else B2W_ALIAS_SMEM(true, false, false); }
    else { if (packed) B2W_ALIAS_SMEM(false, false, true); else B2W_ALIAS_SMEM(false, false, false); }
#undef B2W_ALIAS_SMEM
    return b2w_cuda_fail(cudaGetLastError(), "alias_build_smem_kernel launch");
  }
  uint64_t blocks = (n_tables + 127) / 128;
  if (blocks > lanes / 128) blocks = lanes / 128;
  uint32_t stride = g->max_degree ? g->max_degree : 1;
  if (first_order)
    alias_build_kernel<false, true, false><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  else if (extend && packed)
    alias_build_kernel<true, false, true><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  else if (extend)
    alias_build_kernel<true, false, false><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  else if (packed)
    alias_build_kernel<false, false, true><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  else
    alias_build_kernel<false, false, false><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  return b2w_cuda_fail(cudaGetLastError(), "alias_build_kernel launch");
}

void b2w_fill_bias_params(WalkParams& P, double p, double q);

static WalkParams alias_params(const b2w_graph* g, double p, double q, const float* d_thr) {
  WalkParams P{};
  P.n = g->n; P.indptr = g->indptr; P.indices = g->indices; P.data = g->data; P.thr = d_thr;
  b2w_fill_bias_params(P, p, q);
  return P;
}

extern "C" int b2w_alias_build(const b2w_graph* g, double p, double q, int extend, const float* d_thr,
                               const uint64_t* d_alias_indptr, uint32_t* d_alias_j, float* d_alias_q,
                               void* d_work, size_t work_bytes, void* stream) {
  return alias_build_common(g, p, q, extend, d_thr, d_alias_indptr, d_alias_j, d_alias_q, d_work, work_bytes,
                            stream, false);
}

extern "C" int b2w_alias_build_packed(const b2w_graph* g, double p, double q, int extend, const float* d_thr,
                                      const uint64_t* d_alias_indptr, uint64_t* d_alias_qj, void* d_work,
                                      size_t work_bytes, void* stream) {
  if (d_alias_qj && (reinterpret_cast<uintptr_t>(d_alias_qj) & 7) != 0) { b2w_set_error("alias build: packed table must be 8-byte aligned"); return B2W_ERR_INVALID; }
  return alias_build_common(g, p, q, extend, d_thr, d_alias_indptr, nullptr, reinterpret_cast<float*>(d_alias_qj), d_work,
                            work_bytes, stream, false, true);
}

extern "C" int b2w_alias_build_first_order(const b2w_graph* g, uint32_t* d_alias_j, float* d_alias_q,
                                           void* d_work, size_t work_bytes, void* stream) {
  return alias_build_common(g, 1.0, 1.0, 0, nullptr, nullptr, d_alias_j, d_alias_q, d_work, work_bytes, stream, true);
}
