// b2w_shared.cu -- walk matrices shared between the ranks of one node (CUDA IPC), and their all-gather by the copy
// engines (no kernel, no SMs).
//
// Multi-process jobs (one rank per GPU) end with every rank holding the whole walk matrix.  The walk
// kernels fill the chip, so a collective that runs its own kernels beside them (NCCL) takes SMs away and gains nothing
// when overlapped with the walk.  Here every rank allocates its matrix through
// b2w_shared_alloc (cudaMalloc + an IPC handle), maps the peers' matrices with b2w_shared_open (cudaIpcOpenMemHandle
// from its OWN device: peer access over NVLink is enabled lazily, no context is created on the peer GPU), and after
// each walked batch calls b2w_push_rows: one cudaMemcpyAsync per peer, device to device, on a local side stream --
// DMA over NVLink while the next batch is being walked.  A barrier between the ranks ends the pass.
// The same mapped matrices are what b2w_walk_mirrored (b2w_api.cu, b2w_rowout.cuh) stores into from inside the walk
// kernel -- the variant that removes the gather altogether.
#include <cstring>

#include "b2w_common.cuh"

extern "C" int b2w_shared_alloc(int device, size_t bytes, void** d_ptr, unsigned char handle[64]) {
  if (!d_ptr || !handle || bytes == 0) { b2w_set_error("b2w_shared_alloc: bad argument"); return B2W_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  *d_ptr = nullptr;
  B2W_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  B2W_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return b2w_cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle, &h, 64);
  *d_ptr = p;
  return B2W_OK;
}

extern "C" int b2w_shared_free(int device, void* d_ptr) {
  if (!d_ptr) return B2W_OK;
  B2W_CUDA(cudaSetDevice(device));
  B2W_CUDA(cudaFree(d_ptr));
  return B2W_OK;
}

extern "C" int b2w_shared_open(int device, const unsigned char handle[64], void** d_ptr) {
  if (!d_ptr || !handle) { b2w_set_error("b2w_shared_open: bad argument"); return B2W_ERR_INVALID; }
  *d_ptr = nullptr;
  B2W_CUDA(cudaSetDevice(device));                                    // the LOCAL device: the mapping lives in its context
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  B2W_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = p;
  return B2W_OK;
}

extern "C" int b2w_shared_close(int device, void* d_ptr) {
  if (!d_ptr) return B2W_OK;
  B2W_CUDA(cudaSetDevice(device));
  B2W_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return B2W_OK;
}

extern "C" int b2w_push_rows(int device, void* const* d_peers, int n_peers, int self, uint64_t row_lo, uint64_t rows,
                             uint64_t row_bytes, void* stream) {
  if (!d_peers || n_peers < 1 || self < 0 || self >= n_peers || !d_peers[self]) { b2w_set_error("b2w_push_rows: bad argument"); return B2W_ERR_INVALID; }
  if (rows == 0) return B2W_OK;
  B2W_CUDA(cudaSetDevice(device));
  const size_t off = (size_t)(row_lo * row_bytes), n = (size_t)(rows * row_bytes);
  const char* src = static_cast<const char*>(d_peers[self]) + off;
  for (int k = 1; k < n_peers; ++k) {
    const int p = (self + k) % n_peers;                               // every rank starts with a different peer
    if (!d_peers[p]) { b2w_set_error("b2w_push_rows: peer %d not mapped", p); return B2W_ERR_INVALID; }
    B2W_CUDA(cudaMemcpyAsync(static_cast<char*>(d_peers[p]) + off, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return B2W_OK;
}

// The same with one stream per peer (streams[p], p != self), so that the copies to different peers run on different
// copy engines at the same time: one stream drives ~455 GB/s out of a GPU with 7 peers, NVLink takes more.
extern "C" int b2w_push_rows_streams(int device, void* const* d_peers, int n_peers, int self, uint64_t row_lo,
                                     uint64_t rows, uint64_t row_bytes, void* const* streams) {
  if (!d_peers || !streams || n_peers < 1 || self < 0 || self >= n_peers || !d_peers[self]) { b2w_set_error("b2w_push_rows_streams: bad argument"); return B2W_ERR_INVALID; }
  if (rows == 0) return B2W_OK;
  B2W_CUDA(cudaSetDevice(device));
  const size_t off = (size_t)(row_lo * row_bytes), n = (size_t)(rows * row_bytes);
  const char* src = static_cast<const char*>(d_peers[self]) + off;
  for (int k = 1; k < n_peers; ++k) {
    const int p = (self + k) % n_peers;                               // every rank starts with a different peer
    if (!d_peers[p]) { b2w_set_error("b2w_push_rows_streams: peer %d not mapped", p); return B2W_ERR_INVALID; }
    B2W_CUDA(cudaMemcpyAsync(static_cast<char*>(d_peers[p]) + off, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)streams[p]));
  }
  return B2W_OK;
}

extern "C" int b2w_allgather_rows(int device, void* const* d_peers, int n_peers, int self, uint64_t rows_per_rank,
                                  uint32_t row_len, void* stream) {
  if (self < 0) { b2w_set_error("b2w_allgather_rows: bad argument"); return B2W_ERR_INVALID; }
  return b2w_push_rows(device, d_peers, n_peers, self, (uint64_t)self * rows_per_rank, rows_per_rank, 4ull * row_len, stream);
}
