/*
 * sglb200.h -- C ABI of libsglb200.so, the B200 (sm_100a) implementation of PKU-DAIR/SGL's SGAP pre-processing
 * hot path: K-hop CSR x dense propagation  [X, A^X, ..., A^^K X]  and the cross-hop message aggregation.
 *
 * Every entry point is extern "C", takes plain pointers and sizes (no torch / C++ types), returns an int status
 * (SGLB200_OK == 0) and never throws; sglb200_last_error() returns a thread-local description of the last failure.
 * Device pointers are ordinary CUDA device addresses (e.g. torch.Tensor.data_ptr()); `stream` is a cudaStream_t
 * passed as void* (NULL = the legacy default stream).  Calls taking device pointers are asynchronous with respect
 * to the host; calls taking host pointers return after the result is in host memory.
 * A graph handle is used by one host thread at a time (the reference path is single threaded).
 *
 * Reference interfaces replaced (file:line in PKU-DAIR/SGL @ 69cb3248):
 *   sgl/operators/csrc/matmul.h:5        void FloatCSRMulDenseOMP(float[],float[],int[],int[],float[],int,int)
 *   sgl/operators/csrc/cudamatmul.c:28   int  FloatCSRMulDense(float[],int,float[],int[],int[],float[],int,int)
 *   sgl/operators/utils.py:10-40         csr_sparse_dense_matmul(adj, feature)         -> sglb200_spmm*
 *   sgl/operators/base_op.py:19-36       GraphOp.propagate(adj, feature)               -> sglb200_propagate*
 *   sgl/operators/utils.py:76-88         adj_to_symmetric_norm                         -> sglb200_normalize_values
 *   sgl/operators/message_op/*.py        MessageOp._combine (sum/mean/max/min/concat/weighted/over-smooth)
 *                                                                                      -> sglb200_aggregate*
 *   sgl/operators/message_op/learnable_weighted_messahe_op.py:59-101                   -> sglb200_lw_*
 */
#ifndef SGLB200_H_
#define SGLB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SGLB200_API __attribute__((visibility("default")))
#else
#define SGLB200_API
#endif

#define SGLB200_VERSION 100 /* major*10000 + minor*100 + patch -> 0.1.0 */

enum sglb200_status {
    SGLB200_OK = 0,
    SGLB200_ERR_INVALID = 1,   /* bad argument (NULL, negative size, unsupported combination) */
    SGLB200_ERR_CUDA = 2,      /* a CUDA runtime call failed; see sglb200_last_error() */
    SGLB200_ERR_NO_DEVICE = 3, /* no CUDA device / wrong architecture: this library has NO CPU fallback */
    SGLB200_ERR_ALLOC = 4
};

/* where a buffer handed to the library lives */
enum sglb200_location { SGLB200_HOST = 0, SGLB200_DEVICE = 1 };

/* accumulation-order contract of one hop
 *   FAST : rows longer than the split threshold are cut into tiles and re-combined in a fixed order
 *          (deterministic run to run; differs from the reference by fp32 re-association on those rows only)
 *   EXACT: every output element is ONE sequential fused-multiply-add chain in CSR order starting from 0
 *          == bit-for-bit the reference's shipped libmatmul.so (matmul.c:23-40 built with -mfma) */
enum sglb200_mode { SGLB200_MODE_FAST = 0, SGLB200_MODE_EXACT = 1 };

/* cross-hop combiners (reference sgl/operators/message_op/) */
enum sglb200_agg {
    SGLB200_AGG_SUM = 0,      /* sum_message_op.py:9-10   left-to-right fp32 sum                         */
    SGLB200_AGG_MEAN = 1,     /* mean_message_op.py:9-10  sum then one true division by the hop count    */
    SGLB200_AGG_MAX = 2,      /* max_message_op.py:11-12                                                 */
    SGLB200_AGG_MIN = 3,      /* min_message_op.py:11-12                                                 */
    SGLB200_AGG_WEIGHTED = 4, /* simple_weighted_message_op.py:40-56 + utils.py:91-102  acc += y_k * w_k */
    SGLB200_AGG_CONCAT = 5,   /* concat_message_op.py:11-12  hstack into [N, n_feats*d]                  */
    SGLB200_AGG_OSD = 6,      /* over_smooth_distance_op.py:11-33  NAFS cosine-softmax hop weights       */
    SGLB200_AGG_LAST = 7      /* last_message_op.py:9-10  the last hop (sglb200_propagate_fused only)    */
};

typedef struct sglb200_graph *sglb200_graph_t;

/* ---- library / device ------------------------------------------------------------------------------------- */
SGLB200_API int sglb200_version(void);
SGLB200_API const char *sglb200_last_error(void);
/* number of visible CUDA devices, or a negative status; fails with SGLB200_ERR_NO_DEVICE when none */
SGLB200_API int sglb200_device_count(void);
/* selects `device`, checks compute capability 10.x; every later call on this thread uses it */
SGLB200_API int sglb200_set_device(int device);

/* ---- graph handle: CSR operator resident in HBM -------------------------------------------------------------
 * A is n_rows x n_cols (n_rows != n_cols is allowed: a row partition of a larger operator).
 * indptr: n_rows+1 entries, int32 when indptr_is64 == 0 else int64 (nnz >= 2^31 needs int64);
 * indices: nnz int32 column ids; vals: nnz float32 (the normalised weights A^_ij, cast once -- the reference
 * re-casts float64 -> float32 every hop, utils.py:32).  `loc` says whether the three arrays are host or device
 * memory; they are copied, the caller keeps ownership.  The handle owns the CSR copy, the tile schedules and a
 * small workspace that grows with the largest feature width seen.
 * tile_items: merge-path items (rows + non-zeros) per warp, 0 = default; split_threshold: rows with at most
 * this many non-zeros are never cut across warps in FAST mode, 0 = default. */
SGLB200_API int sglb200_graph_create(sglb200_graph_t *out, int64_t n_rows, int64_t n_cols, int64_t nnz, const void *indptr,
                         int indptr_is64, const int32_t *indices, const float *vals, int loc, int tile_items,
                         int split_threshold, void *stream);
SGLB200_API int sglb200_graph_destroy(sglb200_graph_t g);
/* replaces the nnz float32 values of an existing handle (same structure, e.g. another r / alpha) */
SGLB200_API int sglb200_graph_set_values(sglb200_graph_t g, const float *vals, int loc, void *stream);
/* info[0]=n_rows [1]=n_cols [2]=nnz [3]=tiles(FAST) [4]=carry runs(FAST) [5]=tiles(EXACT) [6]=tile_items
 * [7]=split_threshold [8]=bytes resident */
SGLB200_API int sglb200_graph_info(sglb200_graph_t g, int64_t info[9]);

/* ---- a4: degree normalisation (utils.py:76-88, ppr_graph_op.py:19) -------------------------------------------
 * Given the structure of A^ = (A+I)^T already in the handle with raw weights w (= (A+I)[j,i] stored at (i,j)) in
 * raw_w (device or host, nnz float64) and the float64 vectors d_left = deg^(r-1), d_right = deg^(-r)
 * (n entries each), writes vals[i,j] = fl32( (1-alpha) * fl64(fl64(w * d_left[i]) * d_right[j]) + alpha*[i==j] )
 * into the handle (alpha == 0 skips the PPR step exactly like LaplacianGraphOp).  All arithmetic is IEEE float64
 * on the device in the reference's product order, so the float32 values equal the reference's bit for bit. */
SGLB200_API int sglb200_normalize_values(sglb200_graph_t g, const double *raw_w, const double *d_left, const double *d_right,
                             double alpha, int apply_ppr, int loc, void *stream);

/* ---- f2: structure of A^ from a COO edge list, on the device (utils.py:76-78,87, laplacian_graph_op.py:19) -----
 * rows / cols: n_edges int64 device arrays (duplicates allowed, any order), weights: n_edges float32 device or NULL (= 1).
 * Builds A~ = A (+ I when add_identity) exactly as scipy does -- duplicates of A summed in float32 in input order, the
 * identity added in float64, exact zeros dropped -- its float64 row sums `deg` (column order), and the CSR of the
 * TRANSPOSE A~^T (= the structure of A^) with sorted columns.  Returns the number of stored entries in *nnz_out and an
 * opaque builder; sglb200_adjacency_export writes indptr [n+1], indices [nnz] (int32), raw_w [nnz] (float64: the merged
 * weight of A~[j,i] at position (i,j); may be NULL) and deg [n] (may be NULL) into caller-owned device buffers.
 * n < 2^31, n_edges + n < 2^32.  Own radix sort / scan kernels, no host pass over the edges. */
typedef struct sglb200_adj_builder sglb200_adj_builder;
SGLB200_API int sglb200_adjacency_build(sglb200_adj_builder **out, int64_t n, int64_t n_edges, const int64_t *rows,
                                        const int64_t *cols, const float *weights, int add_identity, int64_t *nnz_out,
                                        void *stream);
SGLB200_API int sglb200_adjacency_export(sglb200_adj_builder *b, int64_t *indptr, int32_t *indices, double *raw_w, double *deg,
                                         void *stream);
SGLB200_API void sglb200_adjacency_free(sglb200_adj_builder *b);

/* ---- a5/a6/a7: one hop  Y = A X  (+ Y_in when accumulate != 0) ----------------------------------------------
 * X: [n_cols, d] float32 with row stride ldx (elements), Y: [n_rows, d] with row stride ldy; device pointers.
 * accumulate != 0 reproduces the reference kernel's `answer += ...` semantics (matmul.c:36-37): each chain
 * starts from the value already in Y. */
SGLB200_API int sglb200_spmm(sglb200_graph_t g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                 int accumulate, void *stream);

/* chunked hop for pipelining a hop against the halo exchange of a row partition:
 * sglb200_graph_chunks splits the warp schedule of `mode` into n_chunks consecutive tile ranges and reports, per
 * chunk, the tile bounds and the rows whose final value that chunk produces (row_bounds[c] .. row_bounds[c+1]-1);
 * sglb200_spmm_tiles runs one such range (cut rows that finish inside it are folded before it returns to the stream).
 * Cut rows are folded by per-row arrival counters kept in the handle, so the ranges of ONE hop must be issued
 * completely, in ascending order, on one stream; an out-of-order range is rejected (SGLB200_ERR_INVALID), a hop that
 * starts at tile 0 while an earlier hop was left unfinished starts from fresh counters, and a hop issued on a
 * different stream than the previous one first waits (host synchronisation) for that hop's fold. */
SGLB200_API int sglb200_graph_chunks(sglb200_graph_t g, int mode, int n_chunks, int64_t *tile_bounds, int64_t *row_bounds);
SGLB200_API int sglb200_spmm_tiles(sglb200_graph_t g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                       int64_t tile_begin, int64_t tile_end, void *stream);

/* ---- a1: K hops, device resident -----------------------------------------------------------------------------
 * hops[0..K] are K+1 device pointers to [n, d] slabs with row stride ld; hops[0] holds X on entry, hops[k] receives
 * A^^k X.  Slabs may be column blocks of one [n, (K+1)*d] concat buffer (ld = (K+1)*d).  Requires n_rows==n_cols. */
SGLB200_API int sglb200_propagate(sglb200_graph_t g, float *const *hops, int64_t ld, int d, int K, int mode, void *stream);

/* ---- a1 + a4 + a8-a10/a13 fused: K hops with the degree normalisation and the cross-hop aggregation folded into the row
 * flush of the hop kernel (models/base_model.py:23-36 = propagate, then a second pass over K+1 matrices).
 * X: [n, d] device, row stride ldx.  hops_out: NULL, or K+1 device pointers ([n, d], row stride ld_hops) of which NULL
 * entries are hops the caller does not want stored (hops_out[0] may equal X).  agg_op: -1 none, or an sglb200_agg value:
 * SUM / MEAN / MAX / MIN / WEIGHTED over the hops [agg_start, agg_end) in the reference's left-to-right order (bit-exact;
 * agg_weights: K+1 floats on the HOST indexed by hop), CONCAT (hop k goes straight into column block k-agg_start of
 * agg_out [n, (agg_end-agg_start)*d]), OSD (all K+1 hops), LAST (hop K).  agg_out: [n, d] with row stride ld_out.
 * fuse_norm != 0 and mode FAST and a handle whose values came from sglb200_normalize_values: the hop streams the raw
 * weights (nothing when all are 1) and applies deg^(r-1) / deg^(-r) / the PPR teleport term in the row flush; results
 * then differ from the materialised-values path by float32 rounding only (<= 1e-6 relative).  d <= 512. */
SGLB200_API int sglb200_propagate_fused(sglb200_graph_t g, const float *X, int64_t ldx, float *const *hops_out, int64_t ld_hops,
                            int d, int K, int mode, int agg_op, int agg_start, int agg_end, const float *agg_weights,
                            float *agg_out, int64_t ld_out, int fuse_norm, void *stream);

/* ---- f3: one label-propagation layer (tricks/utils.py:54-56) as ONE hop:  Y = clamp(alpha * (A X) + res, lo, hi) with the
 * scale, the residual add and the clamp in the hop kernel's row flush (res: [n, d] device or NULL; do_clamp == 0: no
 * clamp).  Arithmetic as the reference's torch expression: separately rounded multiply and add.  d <= 512. */
SGLB200_API int sglb200_spmm_axpby(sglb200_graph_t g, const float *X, int64_t ldx, float *Y, int64_t ldy, int d, int mode,
                       float alpha, const float *res, int64_t ld_res, int do_clamp, float lo, float hi, void *stream);

/* same with host buffers: uploads X (host [n,d] contiguous), runs K hops on the device, downloads hop k into
 * hops_out[k-1] (k = 1..K; NULL entries are skipped, e.g. keep only the last hop).  Synchronous. */
SGLB200_API int sglb200_propagate_host(sglb200_graph_t g, const float *X, float *const *hops_out, int d, int K, int mode);

/* ---- a8-a10, a13: cross-hop aggregation ----------------------------------------------------------------------
 * feats: n_feats device pointers to [n, d] slabs (row stride ld_in); out: [n, d] (CONCAT: [n, n_feats*d]) with row
 * stride ld_out; weights: n_feats float32 on the HOST (WEIGHTED only, else NULL).  The caller applies the
 * reference's [start:end] slicing by passing only the selected slabs. */
SGLB200_API int sglb200_aggregate(int op, const float *const *feats, int n_feats, int64_t n, int d, int64_t ld_in,
                      const float *weights, float *out, int64_t ld_out, void *stream);

/* ---- a11: LearnableWeightedMessageOp (per-node kinds), fused forward / backward ------------------------------
 * kind: 2 gate, 3 ori_ref, 4 jk (learnable_weighted_messahe_op.py:68-86; ori_ref/jk use the reference's as-written
 * [-1, K'] view of the hop-major score vector).  The scalar kinds simple / simple_allow_neg have
 * only K' parameters and no per-node work; they go through sglb200_aggregate(WEIGHTED).
 * feats: n_all device pointers [B, d] contiguous (all hops: the jk reference row spans all of them); the op combines
 * hops [start, end), K' = end-start.  w: parameter vector on the device (gate: d; ori_ref: 2d = [ref | hop];
 * jk: (n_all+1)*d = [all hops | hop]); bias: 1 float on the device.
 * forward: scores [K'*B] (hop-major pre-activation, kept for the backward), hop_w [B, K'] (softmax weights),
 * out [B, d]. */
SGLB200_API int sglb200_lw_forward(int kind, const float *const *feats, int n_all, int start, int end, int64_t B, int d,
                       const float *w, const float *bias, float *scores, float *hop_w, float *out, void *stream);
/* backward: grad_out [B, d] -> grad_feats[k] [B, d] for all n_all hops (ACCUMULATED into: caller zeroes), grad_w (same
 * length as w, accumulated), grad_bias (1 float, accumulated).  scratch: K'*B floats. */
SGLB200_API int sglb200_lw_backward(int kind, const float *const *feats, int n_all, int start, int end, int64_t B, int d,
                        const float *w, const float *bias, const float *scores, const float *hop_w,
                        const float *grad_out, float *const *grad_feats, float *grad_w, float *grad_bias,
                        float *scratch, void *stream);

/* ---- a12: IterateLearnableWeightedMessageOp ("recursive") fused forward / backward, ProjectedConcat epilogue ----------
 * (iterate_learnable_weighted_message_op.py:28-51, projected_concat_message_op.py:19-28).
 * feats: n_hops device pointers [B, d] contiguous -- the hops [start, end) the op combines (n_hops <= 16); w: 2d floats
 * [w_a | w_b] of the Linear(2d -> 1), bias: 1 float (device).  forward writes dots [B, 2*n_hops] (kept for the backward),
 * hop_w [B, n_hops] (the final weights) and out [B, d]; backward ACCUMULATES into grad_feats[k] [B, d], grad_w (2d) and
 * grad_bias (1): the caller zeroes them. */
SGLB200_API int sglb200_it_forward(const float *const *feats, int n_hops, int64_t B, int d, const float *w, const float *bias,
                       float *dots, float *hop_w, float *out, void *stream);
SGLB200_API int sglb200_it_backward(const float *const *feats, int n_hops, int64_t B, int d, const float *w, const float *bias,
                        const float *dots, const float *grad_out, float *const *grad_feats, float *grad_w,
                        float *grad_bias, void *stream);
/* out [B, n_hops*h]: column block k = ys[k] (k == 0) or relu(ys[k]) (k > 0), ys[k] [B, h] contiguous; the backward writes
 * grad_ys[k] = grad_out block k (masked by ys[k] > 0 for k > 0). */
SGLB200_API int sglb200_relu_concat(const float *const *ys, int n_hops, int64_t B, int h, float *out, void *stream);
SGLB200_API int sglb200_relu_concat_backward(const float *const *ys, int n_hops, int64_t B, int h, const float *grad_out,
                                 float *const *grad_ys, void *stream);

/* ---- f1: device-resident feature store: out[b, :] = feat[idx[b], :] for every hop in one launch --------------
 * (models/base_model.py:58-61 does a CPU fancy-index + H2D per step).  idx: B int64 on the device. */
SGLB200_API int sglb200_gather_rows(const float *const *feats, int n_feats, int64_t ld_in, const int64_t *idx, int64_t B, int d,
                        float *const *outs, int64_t ld_out, void *stream);

/* ---- (e) halo exchange over NVLink peer memory (row partition, one process per GPU) ---------------------------
 * The reference has no collective on this path; these are the primitives of sgl_b200/dist.py's
 * NCCL-free exchange.  ipc_alloc: cudaMalloc + zero + CUDA IPC handle (64 bytes) that the other ranks of the box open
 * with ipc_open (peer access enabled lazily).  push_rows: dst[i, :] = src[rows[i], :] for i < n_rows with dst in a
 * PEER's memory (128-bit stores over NVLink); rows is a device array.  signal_peers: after everything previously
 * enqueued on `stream`, store `value` (release, system scope) to the n flag words whose (peer) addresses are listed
 * in the device array flag_ptrs_dev.  wait_flags: hold `stream` until the n local flag words are all >= value. */
SGLB200_API int sglb200_ipc_alloc(int64_t bytes, void **ptr, unsigned char handle[64]);
SGLB200_API int sglb200_ipc_open(const unsigned char handle[64], void **ptr);
SGLB200_API int sglb200_ipc_close(void *ptr);
SGLB200_API int sglb200_ipc_free(void *ptr);
SGLB200_API int sglb200_push_rows(const float *src, int64_t ld_src, int d, const int64_t *rows, int64_t n_rows, float *dst,
                      int64_t ld_dst, int max_blocks, void *stream);
SGLB200_API int sglb200_signal_peers(unsigned long long *const *flag_ptrs_dev, int n, unsigned long long value, void *stream);
SGLB200_API int sglb200_wait_flags(const unsigned long long *flags_dev, int n, unsigned long long value, void *stream);
/* wait_flags never spins forever: after SGLB200_PEER_TIMEOUT_MS (default 30000) it gives up and raises an error word that
 * the next sglb200_wait_flags reports as SGLB200_ERR_CUDA; sglb200_peer_status returns (and clears) the mask of flag
 * slots that timed out on the current device, 0 when none. */
SGLB200_API int sglb200_peer_status(void);

/* ---- legacy ABI: drop-in for the reference's two shared objects ----------------------------------------------
 * Same symbols, same signatures, host pointers, `answer` is accumulated into (matmul.c:36-37).  Internally:
 * upload -> EXACT-mode kernel -> download.  int32 offsets as in the reference, but N*d may exceed 2^31.
 * The void symbol has no error channel: on failure it prints to stderr, keeps the text in sglb200_last_error() and fills
 * `answer` with NaN (SGLB200_LEGACY_ABORT=1: abort() instead); it never returns a silently wrong buffer. */
SGLB200_API void FloatCSRMulDenseOMP(float answer[], float data[], int indices[], int indptr[], float mat[], int mat_row,
                         int mat_col);
SGLB200_API int FloatCSRMulDense(float answer[], int data_nnz, float data[], int indices[], int indptr[], float mat[],
                     int mat_row, int mat_col);

#ifdef __cplusplus
}
#endif
#endif /* SGLB200_H_ */
