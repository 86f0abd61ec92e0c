// b2w_common.cuh -- shared device/host definitions of the B200 walk engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2w.h"

#define B2W_FULL 0xFFFFFFFFu

struct b2w_host_pipe;

struct b2w_graph {
  int device;
  uint32_t n;
  uint64_t nnz;
  uint32_t max_degree;
  uint32_t flags;
  int num_sms;
  // CSR (borrowed)
  const uint32_t* indptr;
  const uint32_t* indices;
  const float* data;
  // dense (borrowed)
  const double* dense;
  const uint8_t* nonzero;
  // alias tables (borrowed)
  const uint64_t* alias_indptr;
  const uint32_t* alias_j;
  const float* alias_q;
  const uint2* alias_qj;   // packed {q bits, j} table (b2w_alias_build_packed); alias_j / alias_q are then null
  // per-edge index (borrowed; b2w_edge_index.cu)
  const void* edge_rec;
  const uint32_t* edge_tri;
  uint64_t edge_tri_words;
  // exact-cdf checkpoints of the unweighted SparseOTF replay (borrowed; valid for the p, q they were built with)
  const float* edge_ckpt;
  const uint32_t* edge_ckb;
  double edge_ck_p, edge_ck_q;
  // weighted per-edge index (borrowed; b2w_wedge.cu), valid for the bias parameters it was built with
  const void* w_rec;
  const void* w_exc;
  const float* w_bw;
  const double* w_bq;
  const float* w_ckpt;
  const uint32_t* w_ckb;
  const float* w_thr;
  double w_p, w_q;
  int w_extend;
  // staging buffers / streams of b2w_walk_host (lazily allocated, guarded by their own mutex)
  b2w_host_pipe* pipe;
};

// Parameters shared by every walk kernel (passed by value).
struct WalkParams {
  uint32_t n;
  const uint32_t* __restrict__ indptr;
  const uint32_t* __restrict__ indices;
  const float* __restrict__ data;
  const double* __restrict__ dense;
  const uint8_t* __restrict__ nonzero;
  const float* __restrict__ thr;
  const uint64_t* __restrict__ alias_indptr;
  const uint32_t* __restrict__ alias_j;
  const float* __restrict__ alias_q;
  const uint2* __restrict__ alias_qj;
  const uint32_t* __restrict__ start;
  const double* __restrict__ feed;
  uint32_t* __restrict__ out;
  uint64_t ld_out;
  uint64_t row0;
  uint64_t n_rows;
  uint32_t L;
  uint32_t key0, key1;
  int rng_mode;
  int extend;
  double p, q;
  double invq;   // 1/q            (rw/sparse_rw.py:119)
  double supp;   // min(1, 1/q)    (rw/sparse_rw.py:124)
  float invp_f, invq_f;  // exact f32 reciprocals when p / q are powers of two
  int p_pow2, q_pow2;
  uint32_t flags;
  float* work;            // per-warp scratch rows for degrees beyond the smem window
  uint32_t work_stride;   // floats per warp
  unsigned long long* counter;  // dynamic row queue
  b2w_walk_stats* stats;
  // mirrors of the output rows in the other GPUs' matrices (b2w_walk_mirrored): out + mirror_delta[q], in words
  long long mirror_delta[7];
  int n_mirrors;
};

// ---------------------------------------------------------------- Philox4x32-10
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// numba/cpython/randomimpl.py:134-147: two 32-bit words -> 53-bit uniform in [0,1)
__host__ __device__ __forceinline__ double words_to_uniform(uint32_t w0, uint32_t w1) {
  uint32_t a = w0 >> 5, b = w1 >> 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

// Word stream of one (row, step): Philox blocks b = 0, 1, ... concatenated (Appendix B).
struct StepRng {
  uint32_t key0, key1, row_lo, row_hi, step, block, widx;
  uint32_t buf[4];
  __device__ __forceinline__ void begin(uint32_t k0, uint32_t k1, uint64_t row, uint32_t st) {
    key0 = k0; key1 = k1; row_lo = (uint32_t)row; row_hi = (uint32_t)(row >> 32);
    step = st; block = 0; widx = 4;
  }
  __device__ __forceinline__ uint32_t word() {
    if (widx == 4) {
      philox4x32_10(row_lo, row_hi, step, block, key0, key1, buf);
      block++; widx = 0;
    }
    uint32_t w = (widx == 0) ? buf[0] : (widx == 1) ? buf[1] : (widx == 2) ? buf[2] : buf[3];
    widx++;
    return w;
  }
  __device__ __forceinline__ double uniform() {
    uint32_t a = word();
    uint32_t b = word();
    return words_to_uniform(a, b);
  }
  // numba/cpython/randomimpl.py:454-533 ('np' state): masked rejection, no draw when n == 1
  __device__ __forceinline__ uint32_t randint(uint32_t n) {
    if (n == 1) return 0;
    uint32_t mask = 0xFFFFFFFFu >> __clz(n - 1);
    uint32_t v;
    do { v = word() & mask; } while (v >= n);
    return v;
  }
};

__device__ __forceinline__ double step_uniform(const WalkParams& P, uint64_t row_rel, uint32_t step) {
  if (P.rng_mode == B2W_RNG_FEED) return P.feed[row_rel * (uint64_t)P.L + (step - 1)];
  uint32_t o[4];
  uint64_t row = P.row0 + row_rel;
  philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), step, 0u, P.key0, P.key1, o);
  return words_to_uniform(o[0], o[1]);
}

// f32( f64(w) / f64(d) ) -- the reference's `w /= q` with a float64 (or int64) scalar
// (SURVEY.md Appendix A.4).  When d is a power of two the f32 product with the exact
// reciprocal is bit-identical and avoids the f64 divide.
__device__ __forceinline__ float div_by(float w, double d, float inv_f, int pow2) {
  return pow2 ? w * inv_f : (float)((double)w / d);
}

// host-side helpers shared by the .cu files
void b2w_set_error(const char* fmt, ...);
int b2w_cuda_fail(cudaError_t e, const char* what);
#define B2W_CUDA(call)                                              \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return b2w_cuda_fail(e__, #call);       \
  } while (0)

// kernel launchers (defined in the per-mode .cu files)
int b2w_launch_sparse_warp(const b2w_graph* g, const WalkParams& P, cudaStream_t s);
int b2w_launch_thread_walk(const b2w_graph* g, int mode, int extend, const WalkParams& P, cudaStream_t s);
int b2w_launch_dense(const b2w_graph* g, int extend, const WalkParams& P, cudaStream_t s);
size_t b2w_sparse_warp_work_bytes(const b2w_graph* g);
bool b2w_uw_eligible(const b2w_graph* g, double p, double q);
size_t b2w_uw_work_bytes(const b2w_graph* g);
int b2w_launch_uw(const b2w_graph* g, const WalkParams& P, cudaStream_t s);
int b2w_launch_uw_edge(const b2w_graph* g, const WalkParams& P, cudaStream_t s);
int b2w_launch_precomp_edge(const b2w_graph* g, const WalkParams& P, cudaStream_t s);
int b2w_launch_wedge(const b2w_graph* g, const WalkParams& P, cudaStream_t s);
bool b2w_windex_matches(const b2w_graph* g, double p, double q, int extend, const float* d_thr);
bool b2w_uw_grid(const b2w_graph* g, double p, double q, int* grid_exp);
uint32_t b2w_sparse_warp_total_warps(const b2w_graph* g);
