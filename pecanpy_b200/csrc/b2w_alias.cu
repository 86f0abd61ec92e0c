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
#include "b2w_probs.cuh"

namespace {

__device__ __forceinline__ void alias_setup_inplace(uint32_t k, uint32_t* __restrict__ j, float* __restrict__ q,
                                                    uint32_t* __restrict__ stk) {
  // on entry q[kk] holds the normalised probability probs[kk]
  uint32_t sp = 0, lp = 0;   // smaller grows up from stk[0], larger grows down from stk[k-1]
  for (uint32_t kk = 0; kk < k; ++kk) {
    float v = (float)__dmul_rn((double)k, (double)q[kk]);             // q[kk] = k * probs[kk] (:642)
    q[kk] = v;
    j[kk] = 0;
    if (v < 1.0f) stk[sp++] = kk; else stk[k - 1 - lp++] = kk;
  }
  while (sp > 0 && lp > 0) {                                          // (:650-663)
    uint32_t small = stk[--sp];
    uint32_t large = stk[k - 1 - (--lp)];
    j[small] = large;
    float v = (float)__dsub_rn((double)__fadd_rn(q[large], q[small]), 1.0);
    q[large] = v;
    if (v < 1.0f) stk[sp++] = large; else stk[k - 1 - lp++] = large;
  }
}

template <bool EXTEND, bool FIRST_ORDER>
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
    float* q = alias_q + off;
    uint32_t* j = alias_j + off;
    BiasStream<EXTEND> bs(P, idx, !FIRST_ORDER, prev);
    float sum = 0.f;
    for (uint32_t k = 0; k < deg; ++k) {
      float w = bs.weight(k);
      q[k] = w;
      sum = __fadd_rn(sum, w);                                        // sequential f32 sum
    }
    for (uint32_t k = 0; k < deg; ++k) q[k] = __fdiv_rn(q[k], sum);   // rw/sparse_rw.py:89
    alias_setup_inplace(deg, j, q, stk);
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

static int alias_build_common(const b2w_graph* g, double p, double q, int extend, const float* d_thr,
                              const uint64_t* aip, uint32_t* aj, float* aq, void* d_work, size_t work_bytes,
                              void* stream, bool first_order) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("alias build: CSR graph handle required"); return B2W_ERR_INVALID; }
  if (!aj || !aq || (!first_order && !aip)) { b2w_set_error("alias build: null output/offset pointer"); return B2W_ERR_INVALID; }
  if (extend && !d_thr) { b2w_set_error("alias build: extend requires noise thresholds"); return B2W_ERR_INVALID; }
  if (!(p > 0.0) || !(q > 0.0)) { b2w_set_error("alias build: p and q must be > 0"); return B2W_ERR_INVALID; }
  size_t need = b2w_alias_build_work_bytes(g);
  if (work_bytes < need || (!d_work && need)) {
    b2w_set_error("alias build: scratch too small (%zu < %zu bytes)", work_bytes, need);
    return B2W_ERR_INVALID;
  }
  B2W_CUDA(cudaSetDevice(g->device));
  WalkParams P = alias_params(g, p, q, d_thr);
  uint64_t lanes = alias_threads(g);
  uint64_t n_tables = first_order ? g->n : g->nnz;
  if (n_tables == 0) return B2W_OK;
  uint64_t blocks = (n_tables + 127) / 128;
  if (blocks > lanes / 128) blocks = lanes / 128;
  uint32_t stride = g->max_degree ? g->max_degree : 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (first_order)
    alias_build_kernel<false, true><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  else if (extend)
    alias_build_kernel<true, false><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
  else
    alias_build_kernel<false, false><<<(unsigned)blocks, 128, 0, s>>>(P, aip, aj, aq, (uint32_t*)d_work, stride, n_tables);
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

extern "C" int b2w_alias_build_first_order(const b2w_graph* g, uint32_t* d_alias_j, float* d_alias_q,
                                           void* d_work, size_t work_bytes, void* stream) {
  return alias_build_common(g, 1.0, 1.0, 0, nullptr, nullptr, d_alias_j, d_alias_q, d_work, work_bytes, stream, true);
}
