// b2w_csr_build.cu -- edge list -> CSR on the device, with the reference's ingest conventions.
//
// Reference: AdjlstGraph.read / _read_edge_line / add_edge (graph.py:160-305) fill a list of {col: weight} dicts
// line by line -- a later line with the same (row, col) overwrites the earlier weight, an undirected line (a, b, w)
// writes both (a, b) and (b, a) -- and to_csr (graph.py:308-341) emits every row with its columns sorted ascending
// as indptr u32[n+1], indices u32[nnz], data f32[nnz].  The Python loop over 10^7 edges takes minutes (SURVEY.md 8f
// rank 3); here the same result is one stable radix sort:
//   1. expand: entry 2e = (src, dst), entry 2e + 1 = (dst, src) (undirected) keyed (row << 32 | col), payload e.
//      Interleaving keeps entry order == line order, so "later line wins" == "last of a run of equal keys" after a
//      STABLE sort;
//   2. cub::DeviceRadixSort::SortPairs on the 64-bit keys (only the bits that can be set are sorted);
//   3. flag the last entry of every run, exclusive scan -> output slot; scatter col / f32(weight);
//   4. indptr[r] = number of unique keys < (r << 32): one binary search per row over the compacted keys.
// The node numbering (first appearance in the file, graph.py:217-236) and the dropping of non-positive weights stay
// on the host with the text parsing; this file starts from integer endpoints.  CUB is the CUDA toolkit's own
// device-wide sort/scan (library code, as cuBLAS would be for a GEMM) -- ingest is not the walk hot path.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "b2w_common.cuh"

namespace {

__global__ void expand_edges_kernel(uint64_t m, int directed, const uint32_t* __restrict__ src,
                                    const uint32_t* __restrict__ dst, uint64_t* __restrict__ keys,
                                    uint32_t* __restrict__ vals) {
  for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < m; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t a = src[e], b = dst[e];
    if (directed) {
      keys[e] = (a << 32) | b;
      vals[e] = (uint32_t)e;
    } else {
      keys[2 * e] = (a << 32) | b;
      keys[2 * e + 1] = (b << 32) | a;
      vals[2 * e] = (uint32_t)e;
      vals[2 * e + 1] = (uint32_t)e;
    }
  }
}

__global__ void flag_last_kernel(uint64_t M, const uint64_t* __restrict__ keys, uint32_t* __restrict__ flags) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x)
    flags[i] = (i + 1 == M || keys[i] != keys[i + 1]) ? 1u : 0u;
}

__global__ void scatter_unique_kernel(uint64_t M, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                      const uint32_t* __restrict__ flags, const uint32_t* __restrict__ slot,
                                      const double* __restrict__ w, uint64_t* __restrict__ ukeys,
                                      uint32_t* __restrict__ indices, float* __restrict__ data,
                                      unsigned long long* __restrict__ nnz_out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
    if (!flags[i]) continue;
    const uint32_t p = slot[i];
    const uint64_t k = keys[i];
    ukeys[p] = k;
    indices[p] = (uint32_t)k;
    data[p] = w ? (float)w[vals[i]] : 1.0f;                           // float64 -> float32, round to nearest (astype)
    if (i + 1 == M) *nnz_out = (unsigned long long)p + 1ull;
  }
}

__global__ void indptr_kernel(uint32_t n, uint64_t nnz, const uint64_t* __restrict__ ukeys, uint32_t* __restrict__ indptr) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r <= n; r += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t target = r << 32;                                  // first key of row r
    uint64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
    }
    indptr[r] = (uint32_t)lo;
  }
}

struct Layout {
  size_t keys_a, keys_b, vals_a, vals_b, flags, slot, nnz, cub, total, cub_bytes;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int end_bit_for(uint32_t n) {
  int b = 0;
  while (b < 32 && (1ull << b) < (uint64_t)n) ++b;                     // bits needed for a node index
  return 32 + (b ? b : 1);
}

cudaError_t layout_for(uint32_t n, uint64_t M, Layout* L) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)M, 0, end_bit_for(n));
  if (e != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)M);
  if (e != cudaSuccess) return e;
  L->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  size_t off = 0;
  L->keys_a = off; off += align256(M * 8);
  L->keys_b = off; off += align256(M * 8);
  L->vals_a = off; off += align256(M * 4);
  L->vals_b = off; off += align256(M * 4);
  L->flags = off; off += align256(M * 4);
  L->slot = off; off += align256(M * 4);
  L->nnz = off; off += 256;
  L->cub = off; off += align256(L->cub_bytes);
  L->total = off;
  return cudaSuccess;
}

}  // namespace

extern "C" size_t b2w_csr_from_edges_work_bytes(uint32_t num_nodes, uint64_t num_edges, int directed) {
  const uint64_t M = directed ? num_edges : 2 * num_edges;
  if (M == 0) return 256;
  Layout L;
  if (layout_for(num_nodes, M, &L) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return L.total;
}

extern "C" int b2w_csr_from_edges(int device, uint32_t n, uint64_t m, const uint32_t* d_src, const uint32_t* d_dst,
                                  const double* d_weight, int directed, uint32_t* d_indptr, uint32_t* d_indices,
                                  float* d_data, uint64_t* h_nnz, void* d_work, size_t work_bytes, void* stream) {
  if (!d_indptr || !h_nnz || n == 0) { b2w_set_error("b2w_csr_from_edges: null output / empty node set"); return B2W_ERR_INVALID; }
  const uint64_t M = directed ? m : 2 * m;
  if (M >= 0xFFFFFFFFull) { b2w_set_error("b2w_csr_from_edges: too many edges for uint32 indptr (graph.py:325)"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  *h_nnz = 0;
  if (M == 0) {
    B2W_CUDA(cudaMemsetAsync(d_indptr, 0, ((size_t)n + 1) * sizeof(uint32_t), s));
    B2W_CUDA(cudaStreamSynchronize(s));
    return B2W_OK;
  }
  if (!d_src || !d_dst || !d_indices || !d_data || !d_work) { b2w_set_error("b2w_csr_from_edges: null array"); return B2W_ERR_INVALID; }
  Layout L;
  B2W_CUDA(layout_for(n, M, &L));
  if (work_bytes < L.total) { b2w_set_error("b2w_csr_from_edges: scratch too small (%zu < %zu bytes)", work_bytes, L.total); return B2W_ERR_INVALID; }
  char* base = (char*)d_work;
  uint64_t* keys_a = (uint64_t*)(base + L.keys_a);
  uint64_t* keys_b = (uint64_t*)(base + L.keys_b);
  uint32_t* vals_a = (uint32_t*)(base + L.vals_a);
  uint32_t* vals_b = (uint32_t*)(base + L.vals_b);
  uint32_t* flags = (uint32_t*)(base + L.flags);
  uint32_t* slot = (uint32_t*)(base + L.slot);
  unsigned long long* d_nnz = (unsigned long long*)(base + L.nnz);
  void* cub_tmp = base + L.cub;
  size_t cub_bytes = L.cub_bytes;
  const unsigned blocks = 148 * 8, threads = 256;
  expand_edges_kernel<<<blocks, threads, 0, s>>>(m, directed, d_src, d_dst, keys_a, vals_a);
  B2W_CUDA(cudaGetLastError());
  B2W_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const uint64_t*)keys_a, keys_b, (const uint32_t*)vals_a, vals_b,
                                           (int64_t)M, 0, end_bit_for(n), s));
  flag_last_kernel<<<blocks, threads, 0, s>>>(M, keys_b, flags);
  B2W_CUDA(cudaGetLastError());
  cub_bytes = L.cub_bytes;
  B2W_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, (const uint32_t*)flags, slot, (int64_t)M, s));
  // keys_a is free again: it receives the compacted (unique) keys
  scatter_unique_kernel<<<blocks, threads, 0, s>>>(M, keys_b, vals_b, flags, slot, d_weight, keys_a, d_indices, d_data, d_nnz);
  B2W_CUDA(cudaGetLastError());
  unsigned long long nnz = 0;
  B2W_CUDA(cudaMemcpyAsync(&nnz, d_nnz, sizeof nnz, cudaMemcpyDeviceToHost, s));
  B2W_CUDA(cudaStreamSynchronize(s));
  indptr_kernel<<<blocks, threads, 0, s>>>(n, nnz, keys_a, d_indptr);
  B2W_CUDA(cudaGetLastError());
  B2W_CUDA(cudaStreamSynchronize(s));
  *h_nnz = nnz;
  return B2W_OK;
}
