// b2w_thresholds.cu -- node2vec+ noise thresholds on the device, bit-identical to the reference's NumPy.
//
// Reference: SparseRWGraph.get_noise_thresholds (rw/sparse_rw.py:22-35) and
// DenseRWGraph.get_noise_thresholds (rw/dense_rw.py:11-19):
//     thr[i] = max(mean(w_i) + gamma * std(w_i), 0)          w_i = the stored weights of row i
// a Python loop over the nodes there (tens of seconds at 10^6 nodes).  NumPy's mean/std are pairwise
// summations whose association order depends on the row length; b2w_pairwise.cuh restates them, and the
// kernels only decide who runs that (sequential, order-defined) arithmetic and how the operands arrive:
//   CSR    one lane per row.  f32 throughout; rows of neighbouring lanes are adjacent in `data`, so the
//          per-lane streams share sectors through L1.
//   dense  one warp per row.  The compressed row `data[i, nonzero[i]]` (f64) is never materialised: all 32
//          lanes execute the same (redundant, warp-uniform) reduction over a shared-memory ring that the
//          warp refills cooperatively -- coalesced loads of 128 mask bytes + their f64 weights, ballot
//          + popcount to keep the column order.  Three passes per row (count, sum, squares).
#include <cmath>

#include "b2w_common.cuh"
#include "b2w_pairwise.cuh"

namespace {

// NumPy's maximum(x, 0): NaN propagates (fmaxf would drop it).
template <typename T>
__device__ __forceinline__ T clip0(T x) { return (x >= (T)0 || x != x) ? x : (T)0; }

// ---------------------------------------------------------------- CSR
struct CsrRaw {
  const float* p;
  __device__ __forceinline__ float next() { return __ldg(p++); }
};
struct CsrSq {
  const float* p;
  float mean;
  __device__ __forceinline__ float next() {
    const float x = __fsub_rn(__ldg(p++), mean);
    return __fmul_rn(x, x);
  }
};
struct CsrMake {
  const float* row;
  __device__ __forceinline__ CsrRaw raw() const { return CsrRaw{row}; }
  __device__ __forceinline__ CsrSq centered_sq(float mean) const { return CsrSq{row, mean}; }
};

__global__ void __launch_bounds__(128) thresholds_csr_kernel(uint32_t n, const uint32_t* __restrict__ indptr,
                                                             const float* __restrict__ data, float gamma_f,
                                                             float* __restrict__ thr) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = __ldg(indptr + i), e = __ldg(indptr + i + 1);
    float mean, sd;
    b2w_mean_std<float>(CsrMake{data + s}, e - s, mean, sd);
    // np.float32 + python_float * np.float32: the Python scalar is weak (NEP 50), everything stays float32
    thr[i] = clip0(__fadd_rn(mean, __fmul_rn(gamma_f, sd)));
  }
}

// ---------------------------------------------------------------- dense
constexpr int DN_WARPS = 8;
constexpr uint32_t DN_RING = 512;                                     // f64 entries per warp
constexpr uint32_t DN_CHUNK = 128;                                    // columns per refill iteration

struct DenseStream {
  const double* row;
  const uint8_t* nz;
  double* ring;
  uint32_t N, col, head, tail;
  int lane;
  bool sq;
  double mean;
  __device__ __forceinline__ void refill() {
    __syncwarp();                                                     // every lane has consumed the old contents
    while (col < N && tail - head <= DN_RING - DN_CHUNK) {
      bool v[4];
      double w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t c = col + 32u * k + lane;
        v[k] = c < N && nz[c] != 0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = v[k] ? row[col + 32u * k + lane] : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t b = __ballot_sync(B2W_FULL, v[k]);
        if (v[k]) {
          double x = w[k];
          if (sq) { const double t = __dsub_rn(x, mean); x = __dmul_rn(t, t); }
          ring[(tail + __popc(b & ((1u << lane) - 1u))) & (DN_RING - 1)] = x;
        }
        tail += __popc(b);
      }
      col += DN_CHUNK;
    }
    __syncwarp();
  }
  __device__ __forceinline__ double next() {
    if (head == tail) refill();                                       // warp-uniform
    return ring[(head++) & (DN_RING - 1)];
  }
};
struct DenseMake {
  const double* row;
  const uint8_t* nz;
  double* ring;
  uint32_t N;
  int lane;
  __device__ __forceinline__ DenseStream raw() const { return DenseStream{row, nz, ring, N, 0u, 0u, 0u, lane, false, 0.0}; }
  __device__ __forceinline__ DenseStream centered_sq(double mean) const {
    return DenseStream{row, nz, ring, N, 0u, 0u, 0u, lane, true, mean};
  }
};

__global__ void __launch_bounds__(DN_WARPS * 32) thresholds_dense_kernel(uint32_t N, const double* __restrict__ data,
                                                                         const uint8_t* __restrict__ nonzero, double gamma,
                                                                         float* __restrict__ thr) {
  __shared__ double s_ring[DN_WARPS][DN_RING];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (uint64_t i = (uint64_t)blockIdx.x * DN_WARPS + wib; i < N; i += (uint64_t)gridDim.x * DN_WARPS) {
    const double* row = data + i * N;
    const uint8_t* nz = nonzero + i * N;
    uint32_t cnt = 0;
    for (uint32_t c = lane; c < N; c += 32) cnt += nz[c] != 0;
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(B2W_FULL, cnt, o);
    double mean, sd;
    b2w_mean_std<double>(DenseMake{row, nz, s_ring[wib], N, lane}, cnt, mean, sd);
    __syncwarp();                                                     // the ring is reused by the next row
    // float64 scalars; the assignment into the float32 array rounds once (rw/dense_rw.py:16)
    if (lane == 0) thr[i] = clip0((float)__dadd_rn(mean, __dmul_rn(gamma, sd)));
  }
}

}  // namespace

extern "C" int b2w_noise_thresholds(const b2w_graph* g, double gamma, float* d_thr, void* stream) {
  if (!g || !d_thr) { b2w_set_error("b2w_noise_thresholds: null argument"); return B2W_ERR_INVALID; }
  if (!std::isfinite(gamma)) { b2w_set_error("b2w_noise_thresholds: gamma must be finite"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (g->flags & B2W_GRAPH_DENSE) {
    uint64_t blocks = ((uint64_t)g->n + DN_WARPS - 1) / DN_WARPS;
    const uint64_t cap = (uint64_t)g->num_sms * 8;
    if (blocks > cap) blocks = cap;
    thresholds_dense_kernel<<<(unsigned)blocks, DN_WARPS * 32, 0, s>>>(g->n, g->dense, g->nonzero, gamma, d_thr);
  } else {
    uint64_t blocks = ((uint64_t)g->n + 127) / 128;
    const uint64_t cap = (uint64_t)g->num_sms * 64;
    if (blocks > cap) blocks = cap;
    thresholds_csr_kernel<<<(unsigned)blocks, 128, 0, s>>>(g->n, g->indptr, g->data, (float)gamma, d_thr);
  }
  return b2w_cuda_fail(cudaGetLastError(), "noise thresholds kernel launch");
}
