// b2w_api.cu -- the C ABI of libb2w.so (see include/b2w.h): handles, validation, dispatch.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "b2w_common.cuh"

// ---------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

void b2w_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int b2w_cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return B2W_OK;
  b2w_set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return B2W_ERR_CUDA;
}

extern "C" int b2w_version(void) { return B2W_VERSION; }
extern "C" const char* b2w_last_error(void) { return g_err; }
extern "C" int b2w_device_count(int* out) {
  if (!out) { b2w_set_error("b2w_device_count: null out"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaGetDeviceCount(out));
  return B2W_OK;
}

// ---------------------------------------------------------------- validation kernels
struct CheckResult {
  unsigned int bad_indptr, bad_order, bad_index, bad_weight, not_unweighted, max_degree;
};

__global__ void csr_check_kernel(uint32_t n, uint64_t nnz, const uint32_t* __restrict__ indptr,
                                 const uint32_t* __restrict__ indices, const float* __restrict__ data,
                                 CheckResult* res) {
  unsigned int bad_indptr = 0, bad_order = 0, bad_index = 0, bad_weight = 0, notuw = 0, maxdeg = 0;
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t s = indptr[r], e = indptr[r + 1];
    if (e < s || e > nnz) { bad_indptr = 1; continue; }
    if (r == 0 && s != 0) bad_indptr = 1;
    if (r == n - 1 && e != nnz) bad_indptr = 1;
    maxdeg = max(maxdeg, e - s);
    uint32_t last = 0;
    for (uint32_t k = s; k < e; ++k) {
      uint32_t x = indices[k];
      float w = data[k];
      if (x >= n) bad_index = 1;
      if (k > s && x <= last) bad_order = 1;                           // rows sorted ascending, no duplicates
      if (!(w >= 0.f) || !(w < INFINITY)) bad_weight = 1;              // NaN, negative, inf
      if (w != 1.0f) notuw = 1;
      last = x;
    }
  }
  if (bad_indptr) atomicOr(&res->bad_indptr, 1u);
  if (bad_order) atomicOr(&res->bad_order, 1u);
  if (bad_index) atomicOr(&res->bad_index, 1u);
  if (bad_weight) atomicOr(&res->bad_weight, 1u);
  if (notuw) atomicOr(&res->not_unweighted, 1u);
  atomicMax(&res->max_degree, maxdeg);
}

__global__ void dense_check_kernel(uint64_t total, const double* __restrict__ data, const uint8_t* __restrict__ nz,
                                   CheckResult* res) {
  unsigned int bad_weight = 0, bad_mask = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    double w = data[i];
    if (!(w >= 0.0) || !(w < (double)INFINITY)) bad_weight = 1;
    if ((w != 0.0) != (nz[i] != 0)) bad_mask = 1;                      // graph.py:580
  }
  if (bad_weight) atomicOr(&res->bad_weight, 1u);
  if (bad_mask) atomicOr(&res->bad_order, 1u);
}

// Staging state of b2w_walk_host, cached in the handle so that repeated calls do not pay
// cudaMalloc/cudaFree/stream creation again (a PreComp pass is ~2 ms of kernel time).
struct b2w_host_pipe {
  static constexpr int NS = 2;
  std::mutex mu;
  cudaStream_t st[NS] = {nullptr, nullptr};
  uint32_t* d_start[NS] = {nullptr, nullptr};
  uint32_t* d_out[NS] = {nullptr, nullptr};
  void* d_work[NS] = {nullptr, nullptr};
  b2w_walk_stats* d_stats = nullptr;
  cudaEvent_t ev = nullptr;
  size_t start_cap = 0, out_cap = 0, work_cap = 0;
  void release() {
    if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (int k = 0; k < NS; ++k) {
      if (d_start[k]) cudaFree(d_start[k]);
      if (d_out[k]) cudaFree(d_out[k]);
      if (d_work[k]) cudaFree(d_work[k]);
      if (st[k]) cudaStreamDestroy(st[k]);
      d_start[k] = d_out[k] = nullptr; d_work[k] = nullptr; st[k] = nullptr;
    }
    if (d_stats) cudaFree(d_stats);
    d_stats = nullptr;
    start_cap = out_cap = work_cap = 0;
  }
  cudaError_t ensure(size_t start_bytes, size_t out_bytes, size_t work_bytes) {
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < NS && e == cudaSuccess; ++k)
      if (!st[k]) e = cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking);
    if (e == cudaSuccess && !d_stats) e = cudaMalloc(&d_stats, sizeof(b2w_walk_stats));
    if (e == cudaSuccess && start_bytes > start_cap) {
      for (int k = 0; k < NS && e == cudaSuccess; ++k) { if (d_start[k]) cudaFree(d_start[k]); d_start[k] = nullptr; e = cudaMalloc(&d_start[k], start_bytes); }
      start_cap = e == cudaSuccess ? start_bytes : 0;
    }
    if (e == cudaSuccess && out_bytes > out_cap) {
      for (int k = 0; k < NS && e == cudaSuccess; ++k) { if (d_out[k]) cudaFree(d_out[k]); d_out[k] = nullptr; e = cudaMalloc(&d_out[k], out_bytes); }
      out_cap = e == cudaSuccess ? out_bytes : 0;
    }
    if (e == cudaSuccess && work_bytes > work_cap) {
      for (int k = 0; k < NS && e == cudaSuccess; ++k) { if (d_work[k]) cudaFree(d_work[k]); d_work[k] = nullptr; e = cudaMalloc(&d_work[k], work_bytes); }
      work_cap = e == cudaSuccess ? work_bytes : 0;
    }
    return e;
  }
};

void b2w_host_pipe_destroy(b2w_host_pipe* p) {
  if (!p) return;
  p->release();
  delete p;
}

static int new_handle(int device, b2w_graph** out, b2w_graph** gp) {
  if (!out) { b2w_set_error("graph create: null out"); return B2W_ERR_INVALID; }
  *out = nullptr;
  B2W_CUDA(cudaSetDevice(device));
  b2w_graph* g = new (std::nothrow) b2w_graph();
  if (!g) { b2w_set_error("graph create: out of host memory"); return B2W_ERR_NOMEM; }
  memset(g, 0, sizeof *g);
  g->device = device;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete g; return b2w_cuda_fail(e, "cudaGetDeviceProperties"); }
  g->num_sms = prop.multiProcessorCount;
  {
    // The walk kernels gather 16-byte records / 4-byte entries at random: ask L2 to fetch single 32-byte sectors
    // from DRAM instead of the default 64 bytes (ncu, round 2: 2.4 DRAM sectors per gathered sector).  A hint,
    // per device; B2W_L2_FETCH=0 leaves the device setting alone, any other value is passed on.
    const char* e = getenv("B2W_L2_FETCH");
    const long v = e ? strtol(e, nullptr, 10) : 32;
    if (v > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v);
    cudaGetLastError();
  }
  g->pipe = new (std::nothrow) b2w_host_pipe();
  if (!g->pipe) { delete g; b2w_set_error("graph create: out of host memory"); return B2W_ERR_NOMEM; }
  *gp = g;
  return B2W_OK;
}

static int run_check(CheckResult* h, void (*launch)(CheckResult*, void*), void* ctx) {
  CheckResult* d = nullptr;
  B2W_CUDA(cudaMalloc(&d, sizeof(CheckResult)));
  cudaError_t e = cudaMemset(d, 0, sizeof(CheckResult));
  if (e == cudaSuccess) { launch(d, ctx); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(CheckResult), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return b2w_cuda_fail(e, "graph validation");
}

extern "C" int b2w_graph_csr_create(int device, uint32_t n, uint64_t nnz, const uint32_t* d_indptr,
                                    const uint32_t* d_indices, const float* d_data, b2w_graph** out) {
  if (!d_indptr || (nnz && (!d_indices || !d_data))) { b2w_set_error("csr create: null array"); return B2W_ERR_INVALID; }
  if (n == 0) { b2w_set_error("csr create: empty graph"); return B2W_ERR_INVALID; }
  if (nnz >= 0xFFFFFFFFull) { b2w_set_error("csr create: nnz must fit uint32 (reference indptr is uint32, graph.py:325)"); return B2W_ERR_INVALID; }
  b2w_graph* g = nullptr;
  int rc = new_handle(device, out, &g);
  if (rc) return rc;
  g->n = n; g->nnz = nnz; g->indptr = d_indptr; g->indices = d_indices; g->data = d_data;
  g->flags = B2W_GRAPH_CSR;
  struct Ctx { uint32_t n; uint64_t nnz; const uint32_t* ip; const uint32_t* ix; const float* dt; int sms; } ctx{n, nnz, d_indptr, d_indices, d_data, g->num_sms};
  CheckResult h{};
  rc = run_check(&h, [](CheckResult* d, void* c) {
    Ctx* x = (Ctx*)c;
    csr_check_kernel<<<x->sms * 8, 256>>>(x->n, x->nnz, x->ip, x->ix, x->dt, d);
  }, &ctx);
  if (rc) { b2w_host_pipe_destroy(g->pipe); delete g; return rc; }
  if (h.bad_indptr || h.bad_order || h.bad_index || h.bad_weight) {
    b2w_set_error("csr create: invalid graph (%s%s%s%s)", h.bad_indptr ? "indptr not monotone / inconsistent with nnz; " : "",
                  h.bad_order ? "a row is not sorted ascending and duplicate-free; " : "",
                  h.bad_index ? "column index out of range; " : "", h.bad_weight ? "negative, NaN or infinite weight" : "");
    b2w_host_pipe_destroy(g->pipe); delete g;
    return B2W_ERR_GRAPH;
  }
  g->max_degree = h.max_degree;
  if (!h.not_unweighted) g->flags |= B2W_GRAPH_UNWEIGHTED;
  *out = g;
  return B2W_OK;
}

extern "C" int b2w_graph_dense_create(int device, uint32_t n, const double* d_data, const uint8_t* d_nonzero,
                                      b2w_graph** out) {
  if (!d_data || !d_nonzero || n == 0) { b2w_set_error("dense create: null array / empty graph"); return B2W_ERR_INVALID; }
  b2w_graph* g = nullptr;
  int rc = new_handle(device, out, &g);
  if (rc) return rc;
  g->n = n; g->nnz = 0; g->dense = d_data; g->nonzero = d_nonzero; g->flags = B2W_GRAPH_DENSE; g->max_degree = n;
  struct Ctx { uint64_t total; const double* d; const uint8_t* z; int sms; } ctx{(uint64_t)n * n, d_data, d_nonzero, g->num_sms};
  CheckResult h{};
  rc = run_check(&h, [](CheckResult* d, void* c) {
    Ctx* x = (Ctx*)c;
    dense_check_kernel<<<x->sms * 16, 256>>>(x->total, x->d, x->z, d);
  }, &ctx);
  if (rc) { b2w_host_pipe_destroy(g->pipe); delete g; return rc; }
  if (h.bad_weight || h.bad_order) {
    b2w_set_error("dense create: invalid graph (%s%s)", h.bad_weight ? "negative, NaN or infinite weight; " : "",
                  h.bad_order ? "nonzero mask != (data != 0)" : "");
    b2w_host_pipe_destroy(g->pipe); delete g;
    return B2W_ERR_GRAPH;
  }
  *out = g;
  return B2W_OK;
}

extern "C" int b2w_graph_info_get(const b2w_graph* g, b2w_graph_info* out) {
  if (!g || !out) { b2w_set_error("graph info: null argument"); return B2W_ERR_INVALID; }
  out->num_nodes = g->n; out->nnz = g->nnz; out->max_degree = g->max_degree; out->flags = g->flags;
  return B2W_OK;
}

extern "C" void b2w_graph_destroy(b2w_graph* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  b2w_host_pipe_destroy(g->pipe);
  delete g;
}

extern "C" int b2w_graph_set_alias(b2w_graph* g, const uint64_t* aip, const uint32_t* aj, const float* aq) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("set_alias: CSR graph handle required"); return B2W_ERR_INVALID; }
  g->alias_indptr = aip; g->alias_j = aj; g->alias_q = aq; g->alias_qj = nullptr;
  if (aj && aq) g->flags |= B2W_GRAPH_HAS_ALIAS; else g->flags &= ~B2W_GRAPH_HAS_ALIAS;
  return B2W_OK;
}

extern "C" int b2w_graph_set_alias_packed(b2w_graph* g, const uint64_t* aip, const uint64_t* aqj) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("set_alias_packed: CSR graph handle required"); return B2W_ERR_INVALID; }
  if (aqj && !aip) { b2w_set_error("set_alias_packed: the packed layout is for PreComp tables (needs alias_indptr)"); return B2W_ERR_INVALID; }
  g->alias_indptr = aip; g->alias_j = nullptr; g->alias_q = nullptr; g->alias_qj = reinterpret_cast<const uint2*>(aqj);
  if (aqj) g->flags |= B2W_GRAPH_HAS_ALIAS; else g->flags &= ~B2W_GRAPH_HAS_ALIAS;
  return B2W_OK;
}

// ---------------------------------------------------------------- walk dispatch
static bool is_pow2(double v, float* inv) {
  int e;
  double m = frexp(v, &e);
  if (m == 0.5 && e > -60 && e < 60) { *inv = (float)(1.0 / v); return true; }
  *inv = 0.f;
  return false;
}

void b2w_fill_bias_params(WalkParams& P, double p, double q) {
  P.p = p; P.q = q;
  P.invq = 1.0 / q;
  P.supp = P.invq < 1.0 ? P.invq : 1.0;
  P.p_pow2 = is_pow2(p, &P.invp_f);
  P.q_pow2 = is_pow2(q, &P.invq_f);
}

extern "C" size_t b2w_walk_work_bytes(const b2w_graph* g, int mode) {
  if (!g) return 0;
  if (mode == B2W_MODE_SPARSE_OTF && (g->flags & B2W_GRAPH_CSR)) {
    size_t a = b2w_sparse_warp_work_bytes(g), b = b2w_uw_work_bytes(g);
    return a > b ? a : b;
  }
  if (mode == B2W_MODE_DENSE_OTF) return 256;
  return 0;
}

static int walk_impl(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                     const uint32_t* d_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length,
                     uint64_t seed, int rng_mode, const double* d_feed, uint32_t* d_out, uint64_t ld_out,
                     void* d_work, size_t work_bytes, b2w_walk_stats* d_stats, uint32_t flags, void* stream,
                     int n_mirrors, uint32_t* const* d_mirrors);

extern "C" int b2w_walk(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                        const uint32_t* d_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length,
                        uint64_t seed, int rng_mode, const double* d_feed, uint32_t* d_out, uint64_t ld_out,
                        void* d_work, size_t work_bytes, b2w_walk_stats* d_stats, uint32_t flags, void* stream) {
  return walk_impl(g, mode, p, q, extend, d_thr, d_start, row0, n_rows, walk_length, seed, rng_mode, d_feed, d_out, ld_out,
                   d_work, work_bytes, d_stats, flags, stream, 0, nullptr);
}

extern "C" int b2w_walk_mirrored(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                                 const uint32_t* d_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length,
                                 uint64_t seed, uint32_t* d_out, uint64_t ld_out, void* d_work, size_t work_bytes,
                                 b2w_walk_stats* d_stats, uint32_t flags, void* stream, int n_mirrors,
                                 uint32_t* const* d_mirrors) {
  if (n_mirrors < 0 || n_mirrors > 7 || (n_mirrors && !d_mirrors)) { b2w_set_error("b2w_walk_mirrored: 0..7 mirrors"); return B2W_ERR_INVALID; }
  for (int k = 0; k < n_mirrors; ++k)
    if (!d_mirrors[k] || ((reinterpret_cast<uintptr_t>(d_mirrors[k]) ^ reinterpret_cast<uintptr_t>(d_out)) & 31)) {
      b2w_set_error("b2w_walk_mirrored: mirror %d is null or not congruent to d_out modulo 32 bytes", k);
      return B2W_ERR_INVALID;
    }
  return walk_impl(g, mode, p, q, extend, d_thr, d_start, row0, n_rows, walk_length, seed, B2W_RNG_PHILOX, nullptr, d_out,
                   ld_out, d_work, work_bytes, d_stats, flags, stream, n_mirrors, d_mirrors);
}

static int walk_impl(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                     const uint32_t* d_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length,
                     uint64_t seed, int rng_mode, const double* d_feed, uint32_t* d_out, uint64_t ld_out,
                     void* d_work, size_t work_bytes, b2w_walk_stats* d_stats, uint32_t flags, void* stream,
                     int n_mirrors, uint32_t* const* d_mirrors) {
  if (!g) { b2w_set_error("b2w_walk: null graph"); return B2W_ERR_INVALID; }
  if (n_rows == 0) return B2W_OK;
  if (!d_start || !d_out) { b2w_set_error("b2w_walk: null start/out"); return B2W_ERR_INVALID; }
  if (walk_length < 1 || walk_length > (1u << 24)) { b2w_set_error("b2w_walk: walk_length out of range"); return B2W_ERR_INVALID; }
  if (ld_out < (uint64_t)walk_length + 2) { b2w_set_error("b2w_walk: ld_out < walk_length + 2"); return B2W_ERR_INVALID; }
  if (!(p > 0.0) || !(q > 0.0) || !std::isfinite(p) || !std::isfinite(q)) { b2w_set_error("b2w_walk: p and q must be finite and > 0"); return B2W_ERR_INVALID; }
  if (rng_mode != B2W_RNG_PHILOX && rng_mode != B2W_RNG_FEED) { b2w_set_error("b2w_walk: bad rng_mode"); return B2W_ERR_INVALID; }
  if (rng_mode == B2W_RNG_FEED && !d_feed) { b2w_set_error("b2w_walk: B2W_RNG_FEED needs d_feed"); return B2W_ERR_INVALID; }
  const bool dense = mode == B2W_MODE_DENSE_OTF;
  if (dense != ((g->flags & B2W_GRAPH_DENSE) != 0)) { b2w_set_error("b2w_walk: mode %d does not match the graph layout", mode); return B2W_ERR_INVALID; }
  if (mode < 0 || mode > B2W_MODE_PRECOMP_FIRST_ORDER) { b2w_set_error("b2w_walk: unknown mode %d", mode); return B2W_ERR_INVALID; }
  if (rng_mode == B2W_RNG_FEED && mode != B2W_MODE_SPARSE_OTF && mode != B2W_MODE_DENSE_OTF) {
    b2w_set_error("b2w_walk: B2W_RNG_FEED is defined for the OTF modes only (draw counts are data dependent otherwise)");
    return B2W_ERR_UNSUPPORTED;
  }
  if ((mode == B2W_MODE_PRECOMP || mode == B2W_MODE_PRECOMP_FIRST_ORDER) &&
      (!(g->flags & B2W_GRAPH_HAS_ALIAS) || (mode == B2W_MODE_PRECOMP && !g->alias_indptr))) {
    b2w_set_error("b2w_walk: alias tables not attached (b2w_graph_set_alias)");
    return B2W_ERR_INVALID;
  }
  const bool ext = extend != 0 && (mode == B2W_MODE_SPARSE_OTF || mode == B2W_MODE_DENSE_OTF);
  if (ext && !d_thr) { b2w_set_error("b2w_walk: extend requires noise thresholds"); return B2W_ERR_INVALID; }
  if ((mode == B2W_MODE_FIRST_ORDER_UNWEIGHTED || mode == B2W_MODE_PRECOMP_FIRST_ORDER) && (p != 1.0 || q != 1.0)) {
    b2w_set_error("b2w_walk: first-order modes require p == q == 1 (cli.py:179-197)");
    return B2W_ERR_INVALID;
  }
  const bool warp_kernel = mode == B2W_MODE_SPARSE_OTF && !(flags & B2W_FLAG_THREAD_PER_WALKER);
  size_t need = (warp_kernel || dense) ? b2w_walk_work_bytes(g, mode) : 0;
  if (need && (!d_work || work_bytes < need)) {
    b2w_set_error("b2w_walk: scratch too small (%zu < %zu bytes)", work_bytes, need);
    return B2W_ERR_INVALID;
  }
  B2W_CUDA(cudaSetDevice(g->device));
  WalkParams P{};
  P.n = g->n; P.indptr = g->indptr; P.indices = g->indices; P.data = g->data;
  P.dense = g->dense; P.nonzero = g->nonzero; P.thr = d_thr;
  P.alias_indptr = g->alias_indptr; P.alias_j = g->alias_j; P.alias_q = g->alias_q; P.alias_qj = g->alias_qj;
  P.start = d_start; P.feed = d_feed; P.out = d_out; P.ld_out = ld_out;
  P.row0 = row0; P.n_rows = n_rows; P.L = walk_length;
  P.key0 = (uint32_t)seed; P.key1 = (uint32_t)(seed >> 32);
  P.rng_mode = rng_mode; P.extend = ext ? 1 : 0;
  b2w_fill_bias_params(P, p, q);
  P.flags = flags; P.work = (float*)d_work; P.stats = d_stats;
  P.n_mirrors = n_mirrors;
  for (int k = 0; k < n_mirrors; ++k) P.mirror_delta[k] = (long long)(d_mirrors[k] - d_out);
  cudaStream_t s = (cudaStream_t)stream;
  if (n_mirrors) {
    // only the kernel that writes its rows through WarpRowTile mirrors them (unweighted SparseOTF with the edge index)
    const bool ok = warp_kernel && !ext && !(flags & B2W_FLAG_NO_UNWEIGHTED_KERNEL) && b2w_uw_eligible(g, p, q) &&
                    (g->flags & B2W_GRAPH_HAS_EDGE_INDEX) && !(flags & (B2W_FLAG_NO_EDGE_INDEX | B2W_FLAG_COOP)) && !((flags >> 8) & 0xFF);
    if (!ok) { b2w_set_error("b2w_walk_mirrored: this mode / graph is not served by a mirroring kernel (use b2w_walk + an all-gather)"); return B2W_ERR_UNSUPPORTED; }
  }
  if (dense) return b2w_launch_dense(g, ext, P, s);
  if (warp_kernel) {
    // unweighted graphs with exactly representable biases: the membership-bitmap kernel
    if (!ext && !(flags & B2W_FLAG_NO_UNWEIGHTED_KERNEL) && b2w_uw_eligible(g, p, q)) {
      // with the per-edge index: closed-form steps, one lane per walker (b2w_walk_edge.cu)
      if ((g->flags & B2W_GRAPH_HAS_EDGE_INDEX) && !(flags & (B2W_FLAG_NO_EDGE_INDEX | B2W_FLAG_COOP)) && !((flags >> 8) & 0xFF))
        return b2w_launch_uw_edge(g, P, s);
      return b2w_launch_uw(g, P, s);
    }
    // weighted graphs / node2vec+ / any p, q: the weighted per-edge index built for exactly these parameters
    if (!(flags & (B2W_FLAG_NO_EDGE_INDEX | B2W_FLAG_NO_UNWEIGHTED_KERNEL | B2W_FLAG_COOP)) && !((flags >> 8) & 0xFF) &&
        b2w_windex_matches(g, p, q, ext, d_thr))
      return b2w_launch_wedge(g, P, s);
    return b2w_launch_sparse_warp(g, P, s);
  }
  if (mode == B2W_MODE_PRECOMP_FIRST_ORDER && g->alias_qj) {
    b2w_set_error("b2w_walk: PreCompFirstOrder needs the two-array tables (b2w_graph_set_alias)");
    return B2W_ERR_INVALID;
  }
  // PreComp through the per-edge index: no search of prev in row(cur), one record per step (b2w_walk_edge.cu)
  if (mode == B2W_MODE_PRECOMP && (g->flags & B2W_GRAPH_HAS_EDGE_INDEX) && !(flags & B2W_FLAG_NO_EDGE_INDEX))
    return b2w_launch_precomp_edge(g, P, s);
  return b2w_launch_thread_walk(g, mode, ext, P, s);
}

extern "C" const char* b2w_walk_kernel_name(const b2w_graph* g, int mode, double p, double q, int extend, uint32_t flags) {
  if (!g) return "";
  switch (mode) {
    case B2W_MODE_DENSE_OTF: return "walk_dense_kernel";
    case B2W_MODE_PRECOMP:
      if ((g->flags & B2W_GRAPH_HAS_EDGE_INDEX) && !(flags & B2W_FLAG_NO_EDGE_INDEX)) return "walk_precomp_edge_kernel";
      return "walk_thread_kernel<PRECOMP>";
    case B2W_MODE_FIRST_ORDER_UNWEIGHTED: return "walk_thread_kernel<FIRST_ORDER_UNWEIGHTED>";
    case B2W_MODE_PRECOMP_FIRST_ORDER: return "walk_thread_kernel<PRECOMP_FIRST_ORDER>";
    case B2W_MODE_SPARSE_OTF:
      if (flags & B2W_FLAG_THREAD_PER_WALKER) return "walk_thread_kernel<SPARSE_OTF>";
      if (!extend && !(flags & B2W_FLAG_NO_UNWEIGHTED_KERNEL) && b2w_uw_eligible(g, p, q)) {
        if ((g->flags & B2W_GRAPH_HAS_EDGE_INDEX) && !(flags & (B2W_FLAG_NO_EDGE_INDEX | B2W_FLAG_COOP)) && !((flags >> 8) & 0xFF))
          return "walk_uw_edge_kernel";
        return "walk_uw_kernel";
      }
      if (!(flags & (B2W_FLAG_NO_EDGE_INDEX | B2W_FLAG_NO_UNWEIGHTED_KERNEL | B2W_FLAG_COOP)) && !((flags >> 8) & 0xFF) &&
          (g->flags & B2W_GRAPH_HAS_WINDEX) && g->w_p == p && g->w_q == q && g->w_extend == (extend ? 1 : 0))
        return "walk_wedge_kernel";
      return "walk_sparse_warp_kernel";
  }
  return "";
}

// ---------------------------------------------------------------- host-buffer wrapper
extern "C" int b2w_walk_host(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                             const uint32_t* h_start, uint64_t row0, uint64_t n_rows, uint32_t L, uint64_t seed,
                             uint32_t* h_out, uint64_t batch_rows, b2w_walk_stats* h_stats, uint32_t flags) {
  if (!g) { b2w_set_error("b2w_walk_host: null graph"); return B2W_ERR_INVALID; }
  if (n_rows == 0) return B2W_OK;
  if (!h_start || !h_out) { b2w_set_error("b2w_walk_host: null start/out"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  if (batch_rows == 0) {
    // default: ~8 batches so that H2D / kernel / D2H of neighbouring batches overlap even for small jobs (with two
    // batches the last D2H is fully exposed), between 64K and 512K rows (config #3: 20 batches of 512K rows)
    batch_rows = ((n_rows + 7) / 8 + 1023) & ~(uint64_t)1023;
    if (batch_rows < (1u << 16)) batch_rows = 1u << 16;
    if (batch_rows > (1u << 19)) batch_rows = 1u << 19;
  }
  if (batch_rows > n_rows) batch_rows = n_rows;
  const uint64_t ld = (uint64_t)L + 2;
  const size_t wb = b2w_walk_work_bytes(g, mode);
  b2w_host_pipe* pipe = g->pipe;
  std::lock_guard<std::mutex> lock(pipe->mu);                          // one host-buffer call per handle at a time
  constexpr int NS = b2w_host_pipe::NS;
  int rc = B2W_OK;
  cudaError_t e = pipe->ensure(batch_rows * sizeof(uint32_t), batch_rows * ld * sizeof(uint32_t), wb);
  if (e == cudaSuccess) e = cudaMemsetAsync(pipe->d_stats, 0, sizeof(b2w_walk_stats), pipe->st[0]);
  if (e == cudaSuccess) e = cudaStreamSynchronize(pipe->st[0]);
  // The internal streams are non-blocking: order them after whatever the caller queued on the default stream of
  // this device (tables / thresholds built just before, e.g. by b2w_alias_build(..., stream = NULL)).  Producers on
  // OTHER streams must be synchronised by the caller (b2w.h).
  if (e == cudaSuccess) {
    if (!pipe->ev) e = cudaEventCreateWithFlags(&pipe->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(pipe->ev, cudaStreamLegacy);
    for (int k = 0; k < b2w_host_pipe::NS && e == cudaSuccess; ++k) e = cudaStreamWaitEvent(pipe->st[k], pipe->ev, 0);
  }
  if (e != cudaSuccess) { pipe->release(); return b2w_cuda_fail(e, "b2w_walk_host setup"); }
  uint64_t done = 0;
  for (int b = 0; rc == B2W_OK && done < n_rows; ++b) {
    const int k = b % NS;
    const uint64_t rows = (n_rows - done < batch_rows) ? (n_rows - done) : batch_rows;
    // the stream serialises reuse of this slot's buffers with its previous batch
    e = cudaMemcpyAsync(pipe->d_start[k], h_start + done, rows * sizeof(uint32_t), cudaMemcpyHostToDevice, pipe->st[k]);
    if (e != cudaSuccess) { rc = b2w_cuda_fail(e, "H2D start"); break; }
    rc = b2w_walk(g, mode, p, q, extend, d_thr, pipe->d_start[k], row0 + done, rows, L, seed, B2W_RNG_PHILOX, nullptr,
                  pipe->d_out[k], ld, pipe->d_work[k], wb, pipe->d_stats, flags, pipe->st[k]);
    if (rc) break;
    e = cudaMemcpyAsync(h_out + done * ld, pipe->d_out[k], rows * ld * sizeof(uint32_t), cudaMemcpyDeviceToHost, pipe->st[k]);
    if (e != cudaSuccess) { rc = b2w_cuda_fail(e, "D2H walks"); break; }
    done += rows;
  }
  for (int k = 0; k < NS; ++k) {
    e = cudaStreamSynchronize(pipe->st[k]);
    if (e != cudaSuccess && rc == B2W_OK) rc = b2w_cuda_fail(e, "b2w_walk_host sync");
  }
  if (rc == B2W_OK && h_stats) {
    e = cudaMemcpy(h_stats, pipe->d_stats, sizeof(b2w_walk_stats), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = b2w_cuda_fail(e, "stats D2H");
  }
  return rc;
}

// ---------------------------------------------------------------- several GPUs from one process
// Walkers never interact (pecanpy.py:189-206 writes only row i) and the graph is read-only, so rows shard freely:
// replica k of the graph (one handle per device) walks the contiguous block [k R, (k + 1) R) of the start array,
// R = ceil(n_rows / n_graphs), on its own host thread through b2w_walk_host, straight into its slice of the ONE host
// matrix.  Philox is keyed by the global row, so the matrix is identical for any number of devices.
extern "C" int b2w_walk_multi(int n_graphs, b2w_graph* const* graphs, int mode, double p, double q, int extend,
                              const float* const* d_thr, const uint32_t* h_start, uint64_t n_rows, uint32_t L,
                              uint64_t seed, uint32_t* h_out, uint64_t batch_rows, b2w_walk_stats* h_stats,
                              uint32_t flags) {
  if (n_graphs < 1 || !graphs) { b2w_set_error("b2w_walk_multi: no graph replicas"); return B2W_ERR_INVALID; }
  for (int k = 0; k < n_graphs; ++k) {
    if (!graphs[k]) { b2w_set_error("b2w_walk_multi: null replica %d", k); return B2W_ERR_INVALID; }
    for (int j = 0; j < k; ++j)
      if (graphs[j] == graphs[k]) { b2w_set_error("b2w_walk_multi: replica %d listed twice (one handle per device)", k); return B2W_ERR_INVALID; }
  }
  if (n_rows == 0) return B2W_OK;
  if (!h_start || !h_out) { b2w_set_error("b2w_walk_multi: null start/out"); return B2W_ERR_INVALID; }
  const uint64_t R = (n_rows + (uint64_t)n_graphs - 1) / (uint64_t)n_graphs;
  const uint64_t ld = (uint64_t)L + 2;
  std::vector<int> rc((size_t)n_graphs, B2W_OK);
  std::vector<std::string> msg((size_t)n_graphs);
  std::vector<b2w_walk_stats> st((size_t)n_graphs);
  std::vector<std::thread> th;
  try {
    for (int k = 0; k < n_graphs; ++k) {
      const uint64_t lo = (uint64_t)k * R < n_rows ? (uint64_t)k * R : n_rows;
      const uint64_t hi = lo + R < n_rows ? lo + R : n_rows;
      th.emplace_back([&, k, lo, hi]() {
        memset(&st[(size_t)k], 0, sizeof(b2w_walk_stats));
        if (hi <= lo) return;
        rc[(size_t)k] = b2w_walk_host(graphs[k], mode, p, q, extend, d_thr ? d_thr[k] : nullptr, h_start + lo, lo, hi - lo, L,
                                      seed, h_out + lo * ld, batch_rows, &st[(size_t)k], flags);
        if (rc[(size_t)k]) msg[(size_t)k] = b2w_last_error();       // the message is thread-local: carry it over
      });
    }
  } catch (...) {
    for (auto& t : th) t.join();
    b2w_set_error("b2w_walk_multi: cannot start a host thread per device");
    return B2W_ERR_NOMEM;
  }
  for (auto& t : th) t.join();
  b2w_walk_stats tot{};
  for (int k = 0; k < n_graphs; ++k) {
    if (rc[(size_t)k]) { b2w_set_error("b2w_walk_multi: replica %d: %s", k, msg[(size_t)k].c_str()); return rc[(size_t)k]; }
    tot.steps += st[(size_t)k].steps; tot.exact_replays += st[(size_t)k].exact_replays;
    tot.seq_sums += st[(size_t)k].seq_sums; tot.overflow_choices += st[(size_t)k].overflow_choices;
  }
  if (h_stats) *h_stats = tot;
  return B2W_OK;
}

// ---------------------------------------------------------------- helpers
__global__ void count_steps_kernel(const uint32_t* __restrict__ out, uint64_t n_rows, uint32_t L, uint64_t ld,
                                   unsigned long long* steps) {
  unsigned long long acc = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_rows; i += (uint64_t)gridDim.x * blockDim.x)
    acc += out[i * ld + L + 1] - 1u;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(B2W_FULL, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(steps, acc);
}

extern "C" int b2w_count_steps(const uint32_t* d_out, uint64_t n_rows, uint32_t L, uint64_t ld_out, uint64_t* d_steps,
                               void* stream) {
  if (!d_out || !d_steps) { b2w_set_error("b2w_count_steps: null pointer"); return B2W_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  B2W_CUDA(cudaMemsetAsync(d_steps, 0, 8, s));
  if (n_rows == 0) return B2W_OK;
  uint64_t blocks = (n_rows + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  count_steps_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_out, n_rows, L, ld_out, (unsigned long long*)d_steps);
  return b2w_cuda_fail(cudaGetLastError(), "count_steps_kernel launch");
}

__global__ void philox_selftest_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  uint32_t o[4];
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], o);
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3];
}

extern "C" int b2w_philox_selftest(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t* d = nullptr;
  B2W_CUDA(cudaMalloc(&d, 10 * sizeof(uint32_t)));
  cudaError_t e = cudaMemcpy(d, ctr, 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d + 4, key, 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { philox_selftest_kernel<<<1, 1>>>(d, d + 4, d + 6); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(out, d + 6, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return b2w_cuda_fail(e, "philox selftest");
}
