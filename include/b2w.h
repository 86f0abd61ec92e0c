/*
 * b2w.h -- C ABI of libb2w.so, the B200 (sm_100a) biased random-walk engine.
 *
 * This is the drop-in boundary for ONE hot path of krishnanlab/PecanPy: node2vec /
 * node2vec+ walk generation behind pecanpy.pecanpy.{PreComp,SparseOTF,DenseOTF}
 * .simulate_walks.  PecanPy has no FFI of its own (it is Python + Numba); the seam the
 * entry points below replace is the call of the njit kernel `Base._random_walks`
 * (reference src/pecanpy/pecanpy.py:149-157) and, for PreComp, the table builder
 * `PreComp.preprocess_transition_probs` (pecanpy.py:442-507).  INTEGRATION.md shows the
 * ctypes binding a PecanPy maintainer would add.
 *
 * Conventions
 *   - Plain C: pointers, sizes, scalars.  No torch / C++ types cross this boundary.
 *   - `d_*` pointers are DEVICE pointers owned by the caller (e.g. torch tensors); the
 *     library borrows them for the duration of the call, or of the graph handle for
 *     pointers passed to a *_create function.  `h_*` pointers are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All device
 *     entry points are asynchronous with respect to the host unless stated otherwise.
 *   - Every function returns B2W_OK (0) or a negative b2w_status; b2w_last_error()
 *     returns a thread-local message for the last failure on the calling thread.
 *   - A graph handle is read-only after creation: concurrent b2w_walk calls on
 *     different streams are legal.
 */
#ifndef B2W_H_
#define B2W_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2W_VERSION 100 /* 0.1.0 */

typedef enum {
  B2W_OK = 0,
  B2W_ERR_INVALID = -1,   /* bad argument */
  B2W_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  B2W_ERR_GRAPH = -3,     /* graph failed validation (unsorted / duplicate / negative / NaN) */
  B2W_ERR_UNSUPPORTED = -4,
  B2W_ERR_NOMEM = -5
} b2w_status;

/* Walk strategy: replaces the get_move_forward()/get_has_nbrs() closure pair that the
 * reference passes into _random_walks (pecanpy.py:143-157; abstract in graph.py:99-105). */
typedef enum {
  B2W_MODE_SPARSE_OTF = 0,             /* pecanpy.py:510-561  SparseOTF            */
  B2W_MODE_PRECOMP = 1,                /* pecanpy.py:364-507  PreComp              */
  B2W_MODE_DENSE_OTF = 2,              /* pecanpy.py:564-614  DenseOTF             */
  B2W_MODE_FIRST_ORDER_UNWEIGHTED = 3, /* pecanpy.py:293-309  FirstOrderUnweighted */
  B2W_MODE_PRECOMP_FIRST_ORDER = 4     /* pecanpy.py:312-361  PreCompFirstOrder    */
} b2w_mode;

/* Random-number regime. */
typedef enum {
  /* Philox4x32-10, key = (seed lo, seed hi), counter = (row lo, row hi, step, block):
   * the production regime; output is independent of grid, thread and GPU count. */
  B2W_RNG_PHILOX = 0,
  /* Caller-fed uniforms U[row - row0, step - 1] (double, row-major, leading dim =
   * walk_length): replays the reference's MT19937 stream (regime R1).  OTF modes only. */
  B2W_RNG_FEED = 1
} b2w_rng;

/* b2w_walk flags */
#define B2W_FLAG_FORCE_EXACT_REPLAY 0x1u /* test hook: always take the sequential replay path */
#define B2W_FLAG_NO_FILTER_STATS 0x2u    /* do not update the fallback counters */
#define B2W_FLAG_THREAD_PER_WALKER 0x4u  /* SparseOTF: use the lane-per-walker kernel */
#define B2W_FLAG_NO_UNWEIGHTED_KERNEL 0x8u /* SparseOTF: always use the generic (weight-streaming) kernel */
#define B2W_FLAG_NO_TMA 0x10u /* DenseOTF: per-lane vector loads instead of cp.async.bulk staging */
#define B2W_FLAG_COOP 0x20u /* unweighted SparseOTF, G < 32: warp-cooperative state-machine kernel (long rows by all 32 lanes) */
#define B2W_FLAG_NO_EDGE_INDEX 0x40u /* unweighted SparseOTF / PreComp: ignore an attached edge index (on-the-fly membership kernels) */
#define B2W_FLAG_NO_CKPT 0x80u /* unweighted SparseOTF through the edge index: replay from the start of the row, not from a checkpoint */
#define B2W_FLAG_OFFEDGE_WARP 0x1000000u /* edge-index kernels: converged warps, steps without an edge by the whole warp (default on graphs with rows > 256) */
#define B2W_FLAG_OFFEDGE_LANE 0x2000000u /* edge-index kernels: plain per-lane loops, steps without an edge by their own lane */
#define B2W_FLAG_L2_PERSIST 0x4000000u /* PreComp through the edge index: pin the edge records in L2 (access-policy window) */
#define B2W_FLAG_GROUP(n) (((uint32_t)(n) & 0xFFu) << 8) /* tuning: lanes per walker (8/16/32), 0 = auto */

typedef struct b2w_graph b2w_graph; /* opaque */

typedef struct {
  uint32_t num_nodes;
  uint64_t nnz;
  uint32_t max_degree;
  uint32_t flags; /* B2W_GRAPH_* */
} b2w_graph_info;

#define B2W_GRAPH_CSR 0x1u
#define B2W_GRAPH_DENSE 0x2u
#define B2W_GRAPH_UNWEIGHTED 0x4u /* every stored weight == 1.0f */
#define B2W_GRAPH_HAS_ALIAS 0x8u
#define B2W_GRAPH_HAS_EDGE_INDEX 0x10u
#define B2W_GRAPH_HAS_WINDEX 0x20u /* weighted per-edge index attached (valid for the p, q, extend it was built with) */

typedef struct {
  uint64_t steps;            /* walk steps taken (sum of effective_length - 1)           */
  uint64_t exact_replays;    /* OTF steps whose parallel filter was inconclusive         */
  uint64_t seq_sums;         /* OTF steps whose normaliser needed the sequential f32 sum */
  uint64_t overflow_choices; /* steps that reproduced the reference's choice == degree   */
} b2w_walk_stats;

int b2w_version(void);
const char* b2w_last_error(void);
int b2w_device_count(int* out);

/* ---- graph handles ---------------------------------------------------------------------
 * CSR layout of the reference (typing.py:31, graph.py:323-341): indptr u32[n+1],
 * indices u32[nnz] with every row sorted ascending and duplicate-free, data f32[nnz].
 * `d_indices` MUST have room for nnz+1 elements: element [nnz] is read (never written)
 * when the reference's unchecked `indices[indptr[cur] + choice]` (pecanpy.py:559) is hit with
 * choice == degree on the last non-empty row; set it to 0.
 * Validation (sortedness, uniqueness, weights finite and >= 0) runs on the device and
 * synchronises the stream once.  Replaces: SparseGraph members read as closure constants in
 * pecanpy.py:397-399,535-537. */
int b2w_graph_csr_create(int device, uint32_t num_nodes, uint64_t nnz, const uint32_t* d_indptr,
                         const uint32_t* d_indices, const float* d_data, b2w_graph** out);

/* Dense layout of the reference (graph.py:576-580): data f64[n,n] row-major, nonzero
 * u8/bool[n,n] == (data != 0).  Weights must be finite and >= 0.
 * Replaces: DenseGraph members read in pecanpy.py:589-590. */
int b2w_graph_dense_create(int device, uint32_t num_nodes, const double* d_data,
                           const uint8_t* d_nonzero, b2w_graph** out);

int b2w_graph_info_get(const b2w_graph* g, b2w_graph_info* out);
void b2w_graph_destroy(b2w_graph* g);

/* ---- ingest: `.edg` text -> integer endpoints (HOST code, no device work) ---------------------
 * Replaces the per-line Python of AdjlstGraph.read / _read_edge_line / add_edge / add_node (graph.py:160-180,
 * 217-236, 258-305) with the same conventions: terms = line.strip().split(delimiter), ids stripped, weighted files
 * need exactly three columns, lines with weight <= 0 are ignored and do not register their ids, node index = order
 * of first appearance (id1 before id2).  b2w_edgelist_parse reads the whole file and reports the sizes;
 * b2w_edgelist_fetch copies endpoints / weights (file order) and the ids (NUL separated, node order) into
 * caller buffers of those sizes; b2w_edgelist_free releases the handle.  Files containing non-ASCII bytes return
 * B2W_ERR_UNSUPPORTED (Python's strip() is Unicode aware; fall back to it).  Malformed lines: B2W_ERR_GRAPH. */
typedef struct b2w_edgelist b2w_edgelist; /* opaque */
int b2w_edgelist_parse(const char* path, int weighted, const char* delimiter, b2w_edgelist** out,
                       uint64_t* num_edges, uint32_t* num_nodes, uint64_t* names_bytes, uint64_t* num_dropped);
int b2w_edgelist_fetch(const b2w_edgelist* e, uint32_t* h_src, uint32_t* h_dst, double* h_weight, char* h_names,
                       uint64_t* h_dropped_lines /* room for 20 */, uint32_t* n_dropped_lines);
void b2w_edgelist_free(b2w_edgelist* e);

/* ---- ingest: edge list -> CSR ------------------------------------------------------------
 * Replaces the dict-of-dicts build of AdjlstGraph.read / add_edge (graph.py:160-305) + to_csr (graph.py:308-341)
 * once the text has been parsed into integer endpoints (node numbering by first appearance and the dropping of
 * non-positive weights stay with the host parser): d_src/d_dst u32[m] and d_weight f64[m] (NULL = unweighted,
 * every weight 1) in FILE ORDER.  A later edge with the same (row, col) overwrites an earlier one; directed == 0
 * stores both directions.  Outputs (caller-allocated): d_indptr u32[n+1], d_indices u32 / d_data f32 with room for
 * m (directed) or 2m (undirected) entries -- allocate one more element of d_indices if the arrays go straight
 * into b2w_graph_csr_create.  *h_nnz receives the number of stored entries.  Synchronous (the output size is
 * data dependent).  d_work: at least b2w_csr_from_edges_work_bytes(n, m, directed) bytes of device scratch. */
size_t b2w_csr_from_edges_work_bytes(uint32_t num_nodes, uint64_t num_edges, int directed);
int b2w_csr_from_edges(int device, uint32_t num_nodes, uint64_t num_edges, const uint32_t* d_src,
                       const uint32_t* d_dst, const double* d_weight, int directed, uint32_t* d_indptr,
                       uint32_t* d_indices, float* d_data, uint64_t* h_nnz, void* d_work, size_t work_bytes,
                       void* stream);

/* ---- node2vec+ noise thresholds ----------------------------------------------------------
 * Replaces SparseRWGraph.get_noise_thresholds (rw/sparse_rw.py:22-35) and
 * DenseRWGraph.get_noise_thresholds (rw/dense_rw.py:11-19), Python loops over the nodes there:
 *     d_thr[i] = max(mean(w_i) + gamma * std(w_i), 0),   w_i = stored weights of row i   (f32[n])
 * with NumPy's own arithmetic: pairwise summation, float32 for CSR rows, float64 for dense rows
 * (rounded to float32 on the store), NaN for a row without stored weights.  Bit-identical to NumPy >= 2
 * for a Python-float gamma (the scalar is "weak": the CSR expression stays in float32). */
int b2w_noise_thresholds(const b2w_graph* g, double gamma, float* d_thr, void* stream);

/* ---- PreComp alias tables --------------------------------------------------------------
 * Replaces PreComp.preprocess_transition_probs (pecanpy.py:442-507) + alias_setup
 * (pecanpy.py:617-665).  The caller computes alias_indptr = [0, cumsum(deg^2)] as u64[n+1]
 * (pecanpy.py:474-476) and allocates d_alias_j u32 / d_alias_q f32 with alias_indptr[n] +
 * max_degree elements (the slack is only read, after the reference's failed neighbour search,
 * pecanpy.py:429-434).  `d_thr` (f32[n], get_noise_thresholds, rw/sparse_rw.py:22-35) is
 * required when extend != 0.  `d_work`/`work_bytes`: scratch, at least
 * b2w_alias_build_work_bytes(g) bytes.  Output is bit-identical to the reference's arrays. */
size_t b2w_alias_build_work_bytes(const b2w_graph* g);
int b2w_alias_build(const b2w_graph* g, double p, double q, int extend, const float* d_thr,
                    const uint64_t* d_alias_indptr, uint32_t* d_alias_j, float* d_alias_q,
                    void* d_work, size_t work_bytes, void* stream);

/* The same tables in ONE array of 8-byte entries { float q; uint32 j } (little endian: q in the low word), so that an
 * alias draw (pecanpy.py:668-677: q[kk], then maybe j[kk]) touches one memory sector instead of two.  Entry e is the
 * pair (alias_q[e], alias_j[e]) of the reference's arrays, bit for bit; d_alias_qj needs alias_indptr[n] + max_degree
 * entries, 8-byte aligned.  Attach with b2w_graph_set_alias_packed. */
int b2w_alias_build_packed(const b2w_graph* g, double p, double q, int extend, const float* d_thr,
                           const uint64_t* d_alias_indptr, uint64_t* d_alias_qj, void* d_work, size_t work_bytes,
                           void* stream);
int b2w_graph_set_alias_packed(b2w_graph* g, const uint64_t* d_alias_indptr, const uint64_t* d_alias_qj);

/* PreCompFirstOrder tables (pecanpy.py:336-361): one table per node, laid out like `data`. */
int b2w_alias_build_first_order(const b2w_graph* g, uint32_t* d_alias_j, float* d_alias_q,
                                void* d_work, size_t work_bytes, void* stream);

/* Attach caller-owned tables to the handle (borrowed until detach/destroy).
 * For B2W_MODE_PRECOMP_FIRST_ORDER pass d_alias_indptr = NULL. */
int b2w_graph_set_alias(b2w_graph* g, const uint64_t* d_alias_indptr, const uint32_t* d_alias_j,
                        const float* d_alias_q);

/* ---- per-edge index ----------------------------------------------------------------------
 * What a 2nd-order step along a stored edge e = (prev -> cur) needs and the reference recomputes every time the
 * step is taken: the position of prev in row(cur) (rw/sparse_rw.py:87; np.searchsorted in pecanpy.py:429), the
 * positions in row(cur) of the common neighbours of cur and prev (isnotin, rw/sparse_rw.py:142-230), deg(cur) and
 * cur itself -- a function of the EDGE, so it is computed once per graph (one sorted-row intersection per edge, the
 * work of one walk step per edge) and kept in HBM:
 *     d_rec  b2w_edge_rec[nnz + 1]   16 bytes per stored edge (+ one pad record)
 *     d_tri  uint32[tri_words]       per edge with common neighbours: their count m, then m ascending positions
 * With the index attached, unweighted SparseOTF walks run a lane-per-walker kernel whose step is O(1 + log m)
 * arithmetic on one record (no row is read), and PreComp steps need no search in row(cur).  Walks are bit-identical
 * with and without it.  Two phases because the list size is data dependent:
 *   b2w_edge_index_prepare  writes the records, counts the lists; *h_tri_words = words to allocate for d_tri
 *                           (synchronous; B2W_ERR_UNSUPPORTED when they do not fit 32-bit offsets);
 *   b2w_edge_index_finish   fills the lists and attaches the index (borrowed until detached / destroy; synchronous).
 * d_work: at least b2w_edge_index_work_bytes(g) bytes, the SAME untouched buffer in both calls.  d_rec must be
 * 16-byte aligned.  b2w_graph_set_edge_index(g, NULL, NULL, 0) detaches. */
typedef struct {
  uint32_t nxt; /* indices[e]: the node the edge leads to                                                  */
  uint32_t kpf; /* bits 0..29 lower_bound(row(nxt), src(e)); bit 30: src(e) is NOT in row(nxt); bit 31: list */
  uint32_t tri; /* offset of the edge's list in d_tri (valid when bit 31 of kpf is set)                    */
  uint32_t deg; /* deg(nxt)                                                                                */
} b2w_edge_rec;
size_t b2w_edge_index_work_bytes(const b2w_graph* g);
int b2w_edge_index_prepare(const b2w_graph* g, void* d_rec, void* d_work, size_t work_bytes, uint64_t* h_tri_words,
                           void* stream);
int b2w_edge_index_finish(b2w_graph* g, void* d_rec, uint32_t* d_tri, uint64_t tri_words, void* d_work,
                          size_t work_bytes, void* stream);
int b2w_graph_set_edge_index(b2w_graph* g, const void* d_rec, const uint32_t* d_tri, uint64_t tri_words);

/* Checkpoints for the exact replay of the unweighted SparseOTF kernel (optional, per (p, q)): the ~1 % of the steps
 * whose filter is inconclusive replay the reference's sequential f32 cumsum; on hub rows that is long, and walkers
 * circulating among hubs do it every step.  For every stored edge (prev -> cur) with deg(cur) >= 128 whose reverse
 * edge exists, the exact f32 cdf after elements 127, 255, ... of row(cur) is kept (float[deg(cur)][deg(cur) / 128] per
 * such node), so a replay starts at most 128 positions before the answer.
 *   b2w_edge_ckpt_prepare  d_ckb uint32[n + 1] <- per-node block offsets; *h_ckpt_floats = floats to allocate
 *   b2w_edge_ckpt_finish   fills d_ckpt for (p, q) and attaches it (borrowed; used by walks with the same p, q)
 * Needs the edge index attached; synchronous; d_work >= b2w_edge_ckpt_work_bytes(g). */
size_t b2w_edge_ckpt_work_bytes(const b2w_graph* g);
int b2w_edge_ckpt_prepare(const b2w_graph* g, uint32_t* d_ckb, void* d_work, size_t work_bytes, uint64_t* h_ckpt_floats,
                          void* stream);
int b2w_edge_ckpt_finish(b2w_graph* g, double p, double q, const uint32_t* d_ckb, float* d_ckpt, uint64_t ckpt_floats,
                         void* stream);
int b2w_graph_clear_edge_ckpt(b2w_graph* g);

/* ---- weighted per-edge index (SparseOTF on weighted graphs, node2vec+, any p and q) ------------------------
 * The biased weights of a step taken after arriving over a stored edge (rw/sparse_rw.py:51-130) are a function of the
 * edge and of (p, q, node2vec+ thresholds).  Prepared once per (graph, parameters), they turn the step into
 * O(log deg) arithmetic for one lane: per row the BASE biased weights (slot neither prev nor common) and their f64
 * prefix sums; per edge a 32-byte record (next node, degree, row start, position and weight of the return edge, the
 * reference's exact f32 normaliser of that step), the common neighbours whose weight deviates from the base
 * ("exceptions": position, exact f32 weight, f64 prefix of the deviations) and, for rows of >= 32 slots, checkpoints
 * of the reference's exact f32 cdf every 32 positions (so that the rare exact replay is bounded).  Walks are
 * bit-identical with and without it.  Caller-owned arrays:
 *     d_rec   32 bytes x (nnz + 1), 32-byte aligned        d_bw  float[nnz]        d_bq  double[nnz]
 *     d_ckb   uint32[n + 1] (per-node checkpoint offsets)  d_exc 24 bytes x exc_entries    d_ckpt float[ckpt_floats]
 * b2w_windex_prepare fills d_rec / d_bw / d_bq and reports the two data-dependent sizes (synchronous;
 * B2W_ERR_UNSUPPORTED when they do not fit 32-bit offsets); b2w_windex_finish fills d_exc / d_ckpt, computes the
 * normalisers and attaches the index (borrowed until b2w_graph_clear_windex / destroy; synchronous).  d_work: at least
 * b2w_windex_work_bytes(g) bytes, the SAME untouched buffer in both calls.  b2w_walk uses the index when mode, p, q,
 * extend and the d_thr pointer are the ones it was built with (B2W_FLAG_NO_EDGE_INDEX: never). */
size_t b2w_windex_work_bytes(const b2w_graph* g);
int b2w_windex_prepare(const b2w_graph* g, double p, double q, int extend, const float* d_thr, void* d_rec, float* d_bw,
                       double* d_bq, uint32_t* d_ckb, void* d_work, size_t work_bytes, uint64_t* h_exc_entries,
                       uint64_t* h_ckpt_floats, void* stream);
int b2w_windex_finish(b2w_graph* g, double p, double q, int extend, const float* d_thr, void* d_rec, float* d_bw,
                      double* d_bq, const uint32_t* d_ckb, void* d_exc, uint64_t exc_entries, float* d_ckpt,
                      uint64_t ckpt_floats, void* d_work, size_t work_bytes, void* stream);
int b2w_graph_clear_windex(b2w_graph* g);

/* ---- the walk kernel -------------------------------------------------------------------
 * Replaces Base._random_walks (pecanpy.py:164-210) together with the move_forward closure of
 * `mode`.  Walks rows [row0, row0 + n_rows) of the caller's (host-shuffled, pecanpy.py:135-141)
 * start array: d_start[i] is the start node of global row row0 + i, and row i of d_out
 * (uint32, leading dimension ld_out >= walk_length + 2) receives
 *     [start, step 1, ..., step L, 0-padding ..., effective_length]
 * exactly as the reference's matrix (pecanpy.py:182-206): column L+1 holds the number of valid
 * node entries (L+1, or 1 for an isolated start, or j when the walker is stuck before step j).
 * The kernel writes every element of the row; d_out needs no initialisation.
 * p, q: return / in-out parameters; extend: node2vec+ (requires d_thr, f32[n]).
 * seed: Philox key (B2W_RNG_PHILOX); d_feed: uniforms (B2W_RNG_FEED).
 * d_work/work_bytes: scratch of at least b2w_walk_work_bytes(g, mode) bytes (may be 0/NULL
 * when that returns 0).  d_stats: optional device pointer to a b2w_walk_stats that the kernel
 * ADDS to (zero it first), or NULL. */
size_t b2w_walk_work_bytes(const b2w_graph* g, int mode);
int b2w_walk(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
             const uint32_t* d_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length,
             uint64_t seed, int rng_mode, const double* d_feed, uint32_t* d_out, uint64_t ld_out,
             void* d_work, size_t work_bytes, b2w_walk_stats* d_stats, uint32_t flags, void* stream);

/* b2w_walk (reference: Base._random_walks, pecanpy.py:164-210; the reference has no multi-GPU path) with the
 * all-gather FUSED into the kernel (multi-process jobs on one node): every row the kernel writes to
 * d_out is also stored, sector by sector as it is produced, at the same place of up to 7 other matrices --
 * d_mirrors[k] = the address in peer k's matrix that corresponds to d_out (mapped with b2w_shared_open; congruent to
 * d_out modulo 32 bytes).  The stores travel over NVLink while the walk goes on; when the kernel has finished on every
 * rank (stream sync + a barrier) every matrix holds every row, without a gather phase.  Philox regime only.  Served by
 * the unweighted SparseOTF edge-index kernel; other modes return B2W_ERR_UNSUPPORTED (use b2w_walk + an all-gather). */
int b2w_walk_mirrored(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                      const uint32_t* d_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length, uint64_t seed,
                      uint32_t* d_out, uint64_t ld_out, void* d_work, size_t work_bytes, b2w_walk_stats* d_stats,
                      uint32_t flags, void* stream, int n_mirrors, uint32_t* const* d_mirrors);

/* Name of the kernel b2w_walk would launch for these arguments (static string; for logs/benchmarks). */
const char* b2w_walk_kernel_name(const b2w_graph* g, int mode, double p, double q, int extend, uint32_t flags);

/* Host-buffer convenience wrapper (the end-to-end call): copies h_start to the device in
 * batches, walks, and copies the rows back into h_out [n_rows, walk_length + 2]; H2D, kernel and
 * D2H of consecutive batches overlap on internal streams.  Pinned host buffers are used
 * directly; pageable buffers work but copy slower.  Synchronous.  h_stats may be NULL.
 * Stream contract: the internal streams are ordered after the work already queued on the device's DEFAULT stream
 * (e.g. b2w_alias_build / b2w_noise_thresholds with stream = NULL); inputs produced on any OTHER stream must be
 * complete before the call (synchronise that stream). */
int b2w_walk_host(const b2w_graph* g, int mode, double p, double q, int extend, const float* d_thr,
                  const uint32_t* h_start, uint64_t row0, uint64_t n_rows, uint32_t walk_length,
                  uint64_t seed, uint32_t* h_out, uint64_t batch_rows, b2w_walk_stats* h_stats,
                  uint32_t flags);

/* The same job on SEVERAL GPUs from one process (walkers are independent, the graph is read-only):
 * graphs[k] is a replica of the graph on its own device (thresholds / alias tables / edge index attached per
 * replica; d_thr[k] lives on that device, d_thr may be NULL without node2vec+).  Replica k walks the contiguous block
 * [k R, (k + 1) R), R = ceil(n_rows / n_graphs), of the start array on its own host thread and delivers its rows
 * straight into its slice of the ONE host matrix h_out [n_rows, walk_length + 2] -- the call a single-process host
 * (e.g. PecanPy's simulate_walks) makes to use every GPU of the box.  Philox is keyed by the global row: the matrix
 * is bit-identical for any number of replicas.  Synchronous.  (Multi-process jobs -- one rank per GPU -- call
 * b2w_walk with row0 and all-gather the device blocks with NCCL: pecanpy_b200/dist.py.) */
int b2w_walk_multi(int n_graphs, b2w_graph* const* graphs, int mode, double p, double q, int extend,
                   const float* const* d_thr, const uint32_t* h_start, uint64_t n_rows, uint32_t walk_length,
                   uint64_t seed, uint32_t* h_out, uint64_t batch_rows, b2w_walk_stats* h_stats, uint32_t flags);

/* ---- the start array of Base.simulate_walks (reference pecanpy.py:135-141), host side ---------------------
 * h_start[num_nodes * num_walks] = num_walks copies of 0 .. num_nodes-1, shuffled exactly as
 *     np.random.seed(random_state); np.random.shuffle(start)
 * does (NumPy's legacy generator: Fisher-Yates from the top, partner by masked rejection on 32-bit MT19937 words) --
 * the shuffle fixes the row order of the walk matrix.  mt_key[624] / *mt_pos = the state of NumPy's global generator
 * after the caller has seeded it the reference's way (np.random.get_state()); on return they hold the state after the
 * shuffle (store it back with np.random.set_state).  Several times faster than NumPy at 10^7 walkers (the partners
 * are drawn ahead of the swaps and prefetched).  Plain host code: no device, no handle. */
int b2w_shuffled_start(uint32_t num_nodes, uint32_t num_walks, uint32_t* mt_key, int32_t* mt_pos, uint32_t* h_start);

/* ---- one-node all-gather by the copy engines (multi-process jobs, one rank per GPU) ----------------------
 * Replaces the trailing ncclAllGather of SURVEY.md 8b/8e where it has to OVERLAP with the walk: the walk kernels fill
 * the chip, a collective that runs kernels beside them slows them down; device-to-device DMA does not.
 *   b2w_shared_alloc  cudaMalloc of the rank's full-size walk matrix + its 64-byte IPC handle (exchange the handles
 *                     between the ranks with any host-side channel, e.g. torch.distributed.all_gather_object)
 *   b2w_shared_open   map a PEER's matrix from its handle into this process, from THIS rank's device (peer access
 *                     over NVLink is enabled lazily; nothing runs on the peer GPU)
 *   b2w_push_rows     rows [row_lo, row_lo + rows) of d_peers[self] -> the same rows of every other d_peers[p]:
 *                     one cudaMemcpyAsync per peer on `stream` (order it after the walk of those rows; use a side
 *                     stream so that the next batch's walk overlaps)
 *   b2w_shared_close / b2w_shared_free   unmap a peer's matrix / free the own one (unmap everywhere first).
 * After a rank has synchronised its push stream AND a barrier between the ranks, its matrix holds every row. */
int b2w_shared_alloc(int device, size_t bytes, void** d_ptr, unsigned char handle[64]);
int b2w_shared_free(int device, void* d_ptr);
int b2w_shared_open(int device, const unsigned char handle[64], void** d_ptr);
int b2w_shared_close(int device, void* d_ptr);
int b2w_push_rows(int device, void* const* d_peers, int n_peers, int self, uint64_t row_lo, uint64_t rows,
                  uint64_t row_bytes, void* stream);
/* the same with one stream per peer (streams[p] for p != self): copies to different peers run concurrently */
int b2w_push_rows_streams(int device, void* const* d_peers, int n_peers, int self, uint64_t row_lo, uint64_t rows,
                          uint64_t row_bytes, void* const* streams);
/* SURVEY.md 8b's b2w_allgather_rows (the in-place all-gather that ends a multi-rank job; reference counterpart: none,
 * pecanpy.py:182-206 fills one host matrix), without an NCCL dependency in this library: rank `self` owns rows
 * [self * rows_per_rank, (self + 1) * rows_per_rank) of the [n_peers * rows_per_rank, row_len] u32 matrix and queues
 * their copy into every peer's mapped matrix on `stream` (= b2w_push_rows of that block).  Complete on a rank after
 * every rank has synchronised its stream (a host barrier).  Callers that hold an ncclComm_t can equally run
 * ncclAllGather on the same buffers (pecanpy_b200/dist.py does, by default); b2w_walk_mirrored needs neither. */
int b2w_allgather_rows(int device, void* const* d_peers, int n_peers, int self, uint64_t rows_per_rank,
                       uint32_t row_len, void* stream);

/* Sum of (effective_length - 1) over the rows of a device walk matrix (the metric's unit). */
int b2w_count_steps(const uint32_t* d_out, uint64_t n_rows, uint32_t walk_length, uint64_t ld_out,
                    uint64_t* d_steps /* device u64, overwritten */, void* stream);

/* Philox4x32-10 of one counter/key on the device (known-answer self test). */
int b2w_philox_selftest(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* B2W_H_ */
