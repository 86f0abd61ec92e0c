// b2w_edge_index.cu -- the per-edge index of a CSR graph: everything a 2nd-order step along a stored edge needs,
// in one 16-byte record per edge plus a list of common-neighbour positions.
//
// A 2nd-order node2vec step is taken from `cur` after arriving over the stored edge e = (prev -> cur).  What the
// reference recomputes for that step every time it is taken (rw/sparse_rw.py:51-91, :142-230; pecanpy.py:427-431)
// depends on the EDGE only, not on the walker:
//     * where `prev` sits in row(cur)                          (the return bias `w[prev] /= p`, sparse_rw.py:87;
//                                                               PreComp's np.searchsorted(row(cur), prev), pecanpy.py:429)
//     * which positions of row(cur) hold common neighbours of cur and prev  (isnotin, sparse_rw.py:201-230)
//     * deg(cur) and the node id of cur.
// The index stores exactly that, once, at graph-preparation time (one pass of sorted-row intersections over the
// edges, the same work as ONE walk step per edge):
//     rec[e] = { nxt  = indices[e]                                              (node the edge leads to)
//                kpf  = lower_bound(row(nxt), src(e))  | NOTFOUND << 30 | HAS_TRI << 31
//                tri  = offset of the edge's list in `tri` (valid when HAS_TRI)
//                deg  = deg(nxt) }
//     tri[off] = m, tri[off + 1 .. off + m] = ascending positions k in row(nxt) with row(nxt)[k] in N(src(e)),
//                row(nxt)[k] != src(e)                                          (sparse_rw.py:84: prev is never "common")
// rec has nnz + 1 entries: entry [nnz] serves the reference's unchecked read indices[indptr[cur] + choice] with
// choice == deg on the last row (pecanpy.py:559).  With it an unweighted SparseOTF step is O(1 + log m) arithmetic
// on one record (b2w_walk_edge.cu) and a PreComp step needs no search in row(cur).
//
// Two phases because the list size is data dependent: b2w_edge_index_prepare writes the records, counts and
// prefix-sums the list lengths and reports the number of list words; the caller allocates them;
// b2w_edge_index_finish fills the lists and attaches the index to the handle.
#include "b2w_membership.cuh"
#include "b2w_scan.cuh"

namespace {

constexpr uint32_t KPF_POS_MASK = 0x3FFFFFFFu;
constexpr uint32_t KPF_NOTFOUND = 0x40000000u;
constexpr uint32_t KPF_HAS_TRI = 0x80000000u;
constexpr int EI_THREADS = 256;

// src[e] = owner row of CSR slot e (one warp per row, coalesced)
__global__ void __launch_bounds__(EI_THREADS) edge_src_kernel(const uint32_t n, const uint32_t* __restrict__ indptr,
                                                              uint32_t* __restrict__ src) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)EI_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * EI_THREADS) >> 5;
  for (uint64_t r = warp; r < n; r += nwarps) {
    const uint32_t s = __ldg(indptr + r), e = __ldg(indptr + r + 1);
    for (uint32_t k = s + lane; k < e; k += 32) src[k] = (uint32_t)r;
  }
}

// One warp per block of 32 consecutive edges; the edges of the block are intersected one after the other by all 32
// lanes (the cheaper side is searched in the other, as in membership_bitmap), lane t keeps the result of edge t and
// the records leave with one coalesced 512-byte store.  FILL = false: records + list lengths; FILL = true: lists.
template <bool FILL>
__global__ void __launch_bounds__(EI_THREADS) edge_index_kernel(const uint32_t n, const uint64_t nnz,
                                                                const uint32_t* __restrict__ indptr,
                                                                const uint32_t* __restrict__ indices,
                                                                const uint32_t* __restrict__ src, uint4* __restrict__ rec,
                                                                uint32_t* __restrict__ tri) {
  const Tile<32> T;
  const uint32_t lane = T.lane;
  const uint64_t warp = (blockIdx.x * (uint64_t)EI_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * EI_THREADS) >> 5;
  const uint64_t nblk = (nnz + 1 + 31) >> 5;                          // the pad record [nnz] included
  for (uint64_t blk = warp; blk < nblk; blk += nwarps) {
    const uint64_t e = (blk << 5) + lane;
    uint32_t a = 0, b = 0, as = 0, ad = 0, bs = 0, bd = 0, kpf = 0, off = 0;
    const bool real = e < nnz;
    if (e <= nnz) {
      b = __ldg(indices + e);                                         // [nnz]: the caller's pad element
      if (b < n) { bs = __ldg(indptr + b); bd = __ldg(indptr + b + 1) - bs; }
      if (real) {
        a = __ldg(src + e);
        as = __ldg(indptr + a);
        ad = __ldg(indptr + a + 1) - as;
      }
      if (FILL) { const uint4 r = rec[e]; kpf = r.y; off = r.z; }
    }
    uint32_t work = __ballot_sync(B2W_FULL, FILL ? (real && (kpf & KPF_HAS_TRI)) : (real && bd > 0));
    uint32_t my_cnt = 0, my_kpf = KPF_NOTFOUND;
    while (work) {
      const int t = __ffs(work) - 1;
      work &= work - 1;
      const uint32_t ta = __shfl_sync(B2W_FULL, a, t);
      const uint32_t tas = __shfl_sync(B2W_FULL, as, t), tad = __shfl_sync(B2W_FULL, ad, t);
      const uint32_t tbs = __shfl_sync(B2W_FULL, bs, t), tbd = __shfl_sync(B2W_FULL, bd, t);
      const uint32_t toff = __shfl_sync(B2W_FULL, off, t);
      const uint32_t* const arow = indices + tas;                     // row(prev): tad >= 1 (it holds the edge)
      const uint32_t* const brow = indices + tbs;                     // row(cur):  tbd >= 1
      const uint32_t ka = 31 - __clz(tad), kb = 31 - __clz(tbd);
      uint32_t m = 0;
      if (!FILL) {
        // where prev sits in row(cur) -- or would be inserted (pecanpy.py:429 uses the insertion point as is)
        bool found;
        const uint32_t pos = lower_bound_eq<true>(brow, tbd, ta, kb, found);
        if (lane == (uint32_t)t) my_kpf = pos | (found ? 0u : KPF_NOTFOUND);
      }
      const uint32_t fwd_cost = ((tbd + 31) >> 5) * (ka + 3);
      const uint32_t rev_cost = ((tad + 31) >> 5) * (kb + 3);
      uint32_t* const lst = FILL ? tri + toff + 1 : nullptr;
      if (fwd_cost <= rev_cost) {
        // every neighbour of cur looked up in row(prev): positions come out in order
        for (uint32_t c0 = 0; c0 < tbd; c0 += 32) {
          const uint32_t k = c0 + lane;
          const bool valid = k < tbd;
          const uint32_t x = valid ? __ldg(brow + k) : B2W_NONE;
          bool found;
          lower_bound_eq<true>(arow, tad, x, ka, found);
          const bool hit = valid && found && x != ta;
          const uint32_t bal = __ballot_sync(B2W_FULL, hit);
          if (FILL && hit) lst[m + __popc(bal & ((1u << lane) - 1u))] = k;
          m += __popc(bal);
        }
      } else {
        // every neighbour of prev looked up in row(cur): ascending keys give ascending positions
        for (uint32_t c0 = 0; c0 < tad; c0 += 32) {
          const uint32_t i = c0 + lane;
          const bool valid = i < tad;
          const uint32_t y = valid ? __ldg(arow + i) : B2W_NONE;
          bool found;
          const uint32_t pos = lower_bound_eq<true>(brow, tbd, y, kb, found);
          const bool hit = valid && found && y != ta;
          const uint32_t bal = __ballot_sync(B2W_FULL, hit);
          if (FILL && hit) lst[m + __popc(bal & ((1u << lane) - 1u))] = pos;
          m += __popc(bal);
        }
      }
      if (FILL) { if (lane == 0) tri[toff] = m; }
      else if (lane == (uint32_t)t) my_cnt = m;
    }
    if (!FILL && e <= nnz) {
      if (my_cnt) my_kpf |= KPF_HAS_TRI;
      rec[e] = make_uint4(b, my_kpf, my_cnt ? my_cnt + 1 : 0u, bd);    // .z: list words, turned into the offset by the scan
    }
  }
}

size_t src_bytes(const b2w_graph* g) { return (((size_t)g->nnz + 1) * sizeof(uint32_t) + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t b2w_edge_index_work_bytes(const b2w_graph* g) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) return 0;
  return src_bytes(g) + b2w_scan::work_bytes(g->nnz + 1) + 256;
}

static int check_args(const b2w_graph* g, const void* d_rec, const char* what) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("%s: CSR graph handle required", what); return B2W_ERR_INVALID; }
  if (!d_rec) { b2w_set_error("%s: null record array", what); return B2W_ERR_INVALID; }
  if ((reinterpret_cast<uintptr_t>(d_rec) & 15) != 0) { b2w_set_error("%s: record array must be 16-byte aligned", what); return B2W_ERR_INVALID; }
  if (g->max_degree > KPF_POS_MASK - 1) { b2w_set_error("%s: max degree beyond 2^30", what); return B2W_ERR_UNSUPPORTED; }
  return B2W_OK;
}

extern "C" int b2w_edge_index_prepare(const b2w_graph* g, void* d_rec, void* d_work, size_t work_bytes,
                                      uint64_t* h_tri_words, void* stream) {
  int rc = check_args(g, d_rec, "b2w_edge_index_prepare");
  if (rc) return rc;
  if (!h_tri_words) { b2w_set_error("b2w_edge_index_prepare: null output"); return B2W_ERR_INVALID; }
  const size_t need = b2w_edge_index_work_bytes(g);
  if (!d_work || work_bytes < need) { b2w_set_error("b2w_edge_index_prepare: scratch too small (%zu < %zu bytes)", work_bytes, need); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* src = reinterpret_cast<uint32_t*>(d_work);
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_work) + src_bytes(g));
  const uint64_t count = g->nnz + 1;
  const unsigned grid = (unsigned)g->num_sms * 8;
  edge_src_kernel<<<grid, EI_THREADS, 0, s>>>(g->n, g->indptr, src);
  edge_index_kernel<false><<<grid, EI_THREADS, 0, s>>>(g->n, g->nnz, g->indptr, g->indices, src, reinterpret_cast<uint4*>(d_rec), nullptr);
  B2W_CUDA(cudaGetLastError());
  unsigned long long h_total = 0;
  B2W_CUDA(b2w_scan::exclusive_scan(count, reinterpret_cast<uint32_t*>(d_rec) + 2, 4, sums, &h_total, s));   // field .z of uint4
  *h_tri_words = h_total;
  if (h_total >= 0xFFFFFFFFull) {
    b2w_set_error("b2w_edge_index_prepare: %llu list words do not fit 32-bit offsets (use the on-the-fly kernels)", h_total);
    return B2W_ERR_UNSUPPORTED;
  }
  return b2w_cuda_fail(cudaGetLastError(), "edge index prepare");
}

extern "C" int b2w_edge_index_finish(b2w_graph* g, void* d_rec, uint32_t* d_tri, uint64_t tri_words, void* d_work,
                                     size_t work_bytes, void* stream) {
  int rc = check_args(g, d_rec, "b2w_edge_index_finish");
  if (rc) return rc;
  if (tri_words && !d_tri) { b2w_set_error("b2w_edge_index_finish: null list array"); return B2W_ERR_INVALID; }
  if (!d_work || work_bytes < b2w_edge_index_work_bytes(g)) { b2w_set_error("b2w_edge_index_finish: scratch too small"); return B2W_ERR_INVALID; }
  B2W_CUDA(cudaSetDevice(g->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (tri_words) {
    const unsigned grid = (unsigned)g->num_sms * 8;
    edge_index_kernel<true><<<grid, EI_THREADS, 0, s>>>(g->n, g->nnz, g->indptr, g->indices, reinterpret_cast<const uint32_t*>(d_work),
                                                        reinterpret_cast<uint4*>(d_rec), d_tri);
    B2W_CUDA(cudaGetLastError());
  }
  B2W_CUDA(cudaStreamSynchronize(s));                                 // the index is complete before any walk may use it
  g->edge_rec = d_rec; g->edge_tri = d_tri; g->edge_tri_words = tri_words;
  g->flags |= B2W_GRAPH_HAS_EDGE_INDEX;
  return B2W_OK;
}

extern "C" int b2w_graph_set_edge_index(b2w_graph* g, const void* d_rec, const uint32_t* d_tri, uint64_t tri_words) {
  if (!g || !(g->flags & B2W_GRAPH_CSR)) { b2w_set_error("set_edge_index: CSR graph handle required"); return B2W_ERR_INVALID; }
  if (d_rec && (reinterpret_cast<uintptr_t>(d_rec) & 15) != 0) { b2w_set_error("set_edge_index: record array must be 16-byte aligned"); return B2W_ERR_INVALID; }
  g->edge_rec = d_rec; g->edge_tri = d_tri; g->edge_tri_words = d_rec ? tri_words : 0;
  if (d_rec) g->flags |= B2W_GRAPH_HAS_EDGE_INDEX; else g->flags &= ~B2W_GRAPH_HAS_EDGE_INDEX;
  return B2W_OK;
}
