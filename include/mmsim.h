/* mmsim -- C-ABI of the B200-native metric-learning hot path (libmmsim.so).
 *
 * The reference (johndpope/multimodal_similarity) has no FFI: its boundary is the Python signature of a few
 * leaf functions in src/utils.py and src/networks.py.  Each entry point below replaces one of them; the Python
 * package `multimodal_similarity_b200` mirrors the reference signatures on top of this ABI with ctypes
 * (INTEGRATION.md shows the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative MMSIM_ERR_* code; mmsim_last_error() gives the
 *     thread-local message of the last failure on the calling thread
 *   - all data pointers are DEVICE pointers, row-major contiguous; the caller owns every buffer; the library
 *     never allocates device memory: workspaces are sized by the *_workspace_bytes functions and passed in
 *     (256-byte aligned; cudaMalloc / torch allocations are)
 *   - all work is enqueued on the caller's stream and returns asynchronously; no host synchronisation inside
 *   - there is no CPU fallback: without a CUDA device of compute capability 10.0 the calls fail
 */
#ifndef MMSIM_H_
#define MMSIM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mmsim_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define MMSIM_API __attribute__((visibility("default")))
#else
#define MMSIM_API
#endif

#define MMSIM_OK 0
#define MMSIM_ERR_ARG (-1)
#define MMSIM_ERR_WORKSPACE (-2)
#define MMSIM_ERR_CUDA (-3)
#define MMSIM_ERR_UNSUPPORTED (-4)

#define MMSIM_METRIC_SQEUCLIDEAN 0 /* src/utils.py:335 */
#define MMSIM_METRIC_EUCLIDEAN 1   /* src/utils.py:337  sqrt(sum + 1e-12) */
#define MMSIM_METRIC_L1 2          /* src/utils.py:339 */

#define MMSIM_LOSS_BATCH_HARD 0 /* src/networks.py:797-833 */
#define MMSIM_LOSS_LIFTED 1     /* src/networks.py:835-870 */

#define MMSIM_KNN_MAX_K 112 /* k (+1 with exclude_self) <= 112 on the tcgen05 path */

MMSIM_API int mmsim_version(void);
MMSIM_API const char* mmsim_last_error(void);
/* kernels this library has launched in the calling process so far (reset != 0: return the count and restart it at zero);
 * what bench.py reports as gpu_launches */
MMSIM_API int64_t mmsim_kernel_launches(int reset);

/* out[i*ld + j] = dist(A[i,:], B[j,:]) -- replaces utils.cdist(utils.all_diffs(a, b), metric)
 * (src/utils.py:313-341; the TF twins :302-311,343-360 compute the same values).  fp32, evaluated in NumPy's
 * summation order: bit-identical to the reference's host path. */
MMSIM_API int mmsim_sqdist_f32(const float* A, int64_t M, const float* B, int64_t N, int64_t D, int metric, float* out,
                     int64_t ld, mmsim_stream_t stream);

/* Fused loss forward + backward -- replaces
 *   dists = cdist_tf(all_diffs_tf(E, E)); batch_hard(dists, pids, margin, weighted)   (kind 0)
 *   dists = ...                         ; lifted_loss(dists, pids, margin, weighted)  (kind 1)
 * (src/base_model_batchhard.py:115-124, src/base_model_lifted.py:115-119) and the gradient w.r.t. E that TF
 * autodiff derives.  pids are float32 labels (label_ph is float32 in the reference), 0 = background.
 * soft != 0 selects margin == "soft" (softplus) for batch_hard; otherwise `margin` is the numeric margin.
 * Outputs follow the reference's return tuple: loss[1], num_active[1], diff[N], weights[N],
 * furthest_positive[N], closest_negative[N]; plus the mined column indices pos_idx[N], neg_idx[N]
 * (-1 when the row has no positive / negative; lifted writes -1).  dE[N*D] may be NULL (forward only).
 * Workspace contract: ws must be ZERO-FILLED before its first use with a given (N, D); every successful call leaves the
 * words it depends on zeroed again (the kernel resets its own barrier words), so one call is exactly one cooperative
 * kernel launch -- no memset.  Re-zero (or use a separate workspace) when N or D changes. */
MMSIM_API int mmsim_loss_workspace_bytes(int64_t N, int64_t D, size_t* bytes);
MMSIM_API int mmsim_loss_f32(int kind, const float* E, const float* pids, int64_t N, int64_t D, int soft, float margin,
                   int weighted, float* loss, float* num_active, float* diff, float* weights, float* furthest_positive,
                   float* closest_negative, int32_t* pos_idx, int32_t* neg_idx, float* dE, void* ws, size_t ws_bytes,
                   mmsim_stream_t stream);

/* k nearest gallery rows of every query by the reference's retrieval distance
 *   dist = np.linalg.norm(q - G, axis=1); idx = np.argsort(dist)[:k]        (src/utils.py:73-74)
 * out_dist[nq*k] are the reference's float32 distances bit for bit, out_idx[nq*k] the row indices inside G
 * (a shard holds < 2^31 rows), ordered by (distance, index); -1 / +inf pad when G has fewer than k rows.
 * exclude_self != 0 drops gallery row (self_offset + i) for query i: the leave-one-out of utils.evaluate
 * (src/utils.py:115,172) without the np.delete copy; indices stay in G's numbering.
 * The tcgen05 kernel only filters: every returned distance is recomputed in the reference's arithmetic and the result is
 * either certified or recomputed by the exact fallback (a second tensor-core sweep of the uncertified queries with a
 * threshold derived from their k-th candidate distance; queries without such a bound, or with more ties than a log
 * holds, take a streaming exact scan of the gallery) -- exact for ANY row order and any number of ties.
 * status[8] (device int32): [0] queries that took the exact fallback, [1] queries queued for the streaming scan,
 * [2] how many of those are done.  The call itself runs the first wave of 128; while status[2] < status[1] the caller
 * repeats mmsim_knn_finish_f32 with the same arguments (the Python wrappers do) -- the result is complete when they are
 * equal, which on data without thousands of exact duplicates they are when the call returns.  D <= 256. */
MMSIM_API int mmsim_knn_workspace_bytes(int64_t nq, int64_t ng, int64_t D, int k, size_t* bytes);
MMSIM_API int mmsim_knn_finish_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                         int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status, void* ws, size_t ws_bytes,
                         mmsim_stream_t stream);
MMSIM_API int mmsim_knn_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                  int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status, void* ws, size_t ws_bytes,
                  mmsim_stream_t stream);

/* Measurement hook: the same call restricted to some of its phases, so each kernel can be timed alone with CUDA
 * events on the caller's stream (bench.py, profiles/).  phases is a mask of MMSIM_KNN_PHASE_*; the workspace must
 * hold the state of the earlier phases of a previous call with identical arguments. */
#define MMSIM_KNN_PHASE_PREP 1     /* fp32 -> fp16 operand copies, norms, rounding-error norms */
#define MMSIM_KNN_PHASE_TENSOR 2   /* fused tcgen05 distance + candidate sweep (the dominant kernel) */
#define MMSIM_KNN_PHASE_RERANK 4   /* candidate selection, exact fp32 re-rank, certificate */
#define MMSIM_KNN_PHASE_FALLBACK 8 /* exact recomputation of uncertified queries */
#define MMSIM_KNN_PHASE_PIVOT 16   /* tcgen05 pre-pass over a gallery sample: the 16 smallest sampled keys per query */
#define MMSIM_KNN_PHASE_LADDER 32  /* pivot list -> threshold ladder (after an optional cross-shard merge of the lists) */
#define MMSIM_KNN_PHASE_ALL 63
#define MMSIM_KNN_PHASE_PREP_Q 128 /* mmsim_knn_shard_f32 only: the query half of PREP (operand copies + query grouping) ... */
#define MMSIM_KNN_PHASE_PREP_G 256 /* ... and the gallery half (operand copies, norm pack), each on its own */
MMSIM_API int mmsim_knn_f32_phases(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                         int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status, void* ws, size_t ws_bytes,
                         mmsim_stream_t stream, int phases);

/* The same retrieval with HOST buffers in: what a caller holding NumPy arrays (the reference's retrieve_one loop over
 * embeddings in host memory, src/utils.py:55-81, src/evaluate_model.py) binds.  q_host[nq*D] and g_host[ng*D] are host
 * arrays -- page-locked (cudaHostAlloc / cudaHostRegister) for the transfers to overlap the kernels; pageable memory works
 * but serialises.  The call copies the queries, then a 1/61 sample of the gallery (the pivot pre-pass needs nothing
 * else), then the gallery one split at a time on an internal copy stream; each split is converted and swept while the
 * next one is in flight, so only the first split's transfer is exposed.  q_stage[nq*D] and g_stage[ng*D] are device
 * buffers the caller provides for the float32 copies (the exact re-rank reads them; the library never allocates); the
 * workspace of this call is sized by mmsim_knn_host_workspace_bytes (more gallery splits than the device-resident call).
 * out_dist / out_idx may be device memory or page-locked host memory (the re-rank kernel writes them directly);
 * status is device memory.  Everything is ordered after earlier work on `stream` and complete, as far as `stream` is
 * concerned, when work queued after the call starts.  Results are identical to mmsim_knn_f32 on the same arrays. */
MMSIM_API int mmsim_knn_host_workspace_bytes(int64_t nq, int64_t ng, int64_t D, int k, size_t* bytes);
MMSIM_API int mmsim_knn_host_f32(const float* q_host, int64_t nq, const float* g_host, int64_t ng, int64_t D, int k,
                       int exclude_self, int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status,
                       float* q_stage, float* g_stage, void* ws, size_t ws_bytes, mmsim_stream_t stream);

/* Merge `parts` shard-local results (as produced by mmsim_knn_f32 on each gallery shard and gathered with one
 * NCCL all-gather) into the global top-k ordered by (distance, global index).  Part p's [nq][k] block starts at
 * dist_parts + p * part_stride (same for idx_parts; part_stride in elements, >= nq*k, so a packed
 * [parts][2][nq][k] gather buffer can be merged in place); global index = idx_base[p] + local index.
 * parts * k <= 4096. */
MMSIM_API int mmsim_knn_merge(const float* dist_parts, const int32_t* idx_parts, int64_t part_stride, const int64_t* idx_base,
                    int parts, int64_t nq, int k, float* out_dist, int64_t* out_idx, mmsim_stream_t stream);

/* ---- gallery-sharded retrieval with a reduced re-rank and a GLOBAL certificate ----------------------
 * Per-rank costs that do not shrink with the number of shards (each shard's own threshold estimate, a full 128-row
 * exact re-rank per query and shard) are removed by splitting the call:
 *   1. every rank: mmsim_knn_shard_f32(phases = PREP | PIVOT)   -> its pivot lists in the workspace
 *      (mmsim_knn_pivot_region gives offset/size: [query rows padded to 128][16] float)
 *   2. all-gather the lists, mmsim_knn_merge_pivots(...) back into every rank's workspace: all shards now filter each
 *      query with the same threshold (about 1000 gallery rows below it in TOTAL, not per shard)
 *   3. every rank: mmsim_knn_shard_f32(phases = LADDER | TENSOR | RERANK) with kp << 128: its kp best candidates re-ranked
 *      exactly (out_dist/out_idx [nq*kp], shard-local indices) and out_lb[nq], a lower bound on the true distance of
 *      every row of the shard that was NOT re-ranked
 *   4. exchange, mmsim_knn_merge_certified: global top-k by (distance, global index); status[0] counts queries whose
 *      k-th distance is not below every shard's bound, out_flag[q] (nullable) = that k-th distance for such a query, -1
 *      for a certified one.
 *   5. per-QUERY exact fallback: every rank runs mmsim_knn_shard_fallback_f32 with the flags -- the exact top-k inside its
 *      shard of the flagged queries only (same two tiers as mmsim_knn_f32's fallback; compact [cap, k] blocks, slots in
 *      ascending query order, identical on every rank) -- the blocks are all-gathered and mmsim_knn_merge_patch rewrites
 *      those rows of the merged result.  Only if more than `cap` queries are flagged, or its streaming scan has queries
 *      left (status[1] > status[2]), does the caller repeat the whole call with the plain per-shard mmsim_knn_f32 +
 *      mmsim_knn_merge path (multimodal_similarity_b200/sharded.py).
 * kp is the caller's choice (sharded.py: twice the expected share of the 128 best keys per shard plus a margin).
 * slice_rows > 0 (step 3): the re-rank kernel writes query q's row into block q / slice_rows (blocks slice_stride 32-bit
 * words apart, counted from out_dist / out_idx / out_lb) at row q % slice_rows -- with out_idx = out_dist + slice_rows * kp
 * and out_lb = out_dist + 2 * slice_rows * kp that is exactly the buffer an all-to-all by query slice sends.
 * mmsim_knn_merge_certified writes int64 (out_idx_bits 64) or, when every global index fits 31 bits, int32 indices. */
MMSIM_API int mmsim_knn_shard_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int kp, int exclude_self,
                        int64_t self_offset, float* out_dist, int32_t* out_idx, float* out_lb, int32_t* status, void* ws,
                        size_t ws_bytes, mmsim_stream_t stream, int phases, int64_t slice_rows, int64_t slice_stride);
MMSIM_API int mmsim_knn_pivot_region(int64_t nq, int64_t ng, int64_t D, int k, size_t* offset, size_t* bytes);
/* Introspection (tests): the launch geometry chosen for a problem on a device with num_sms SMs.
 * out[0..12) = padded width, K atoms, 128-query blocks, 256-row gallery tiles, gallery splits, tiles per split, grid,
 * log capacity per (query, split), pivot pre-pass used, tiles of the compact gallery sample, sampled rows, workspace bytes,
 * [12], [13] (if n_out >= 14): workspace offsets of the per-(query, split) candidate counts (int32) and final thresholds;
 * [14], [15]: query-streaming sweep in use, gallery splits of the host-buffer call; [16..20) (if n_out >= 20): workspace
 * offsets of the candidate log (uint2 {key bits, gallery row} [query * splits + split][capacity]), of the fp16 operand
 * copies (queries pre-scaled by -2, then the gallery; rows of `padded width` halves) and of the gallery norm pack
 * ([tile][320] floats, the first 256 the rows' squared norms); [20]: anchors of the query grouping (0: sweep order ==
 * query order). */
MMSIM_API int mmsim_knn_plan(int64_t nq, int64_t ng, int64_t D, int k, int num_sms, int64_t* out, int n_out);
MMSIM_API int mmsim_knn_merge_pivots(const float* parts, int nparts, int64_t part_stride, int64_t rows, float* out,
                           mmsim_stream_t stream);
MMSIM_API int mmsim_knn_merge_certified(const float* dist_parts, const int32_t* idx_parts, int64_t part_stride,
                              const int64_t* idx_base, int parts, int64_t nq, int k_in, int k, const float* lb_parts,
                              int64_t lb_stride, float* out_dist, void* out_idx, int out_idx_bits, int32_t* status,
                              float* out_flag, mmsim_stream_t stream);
MMSIM_API int mmsim_knn_shard_fallback_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                                 int64_t self_offset, const float* flag, int cap, float* out_dist, int32_t* out_idx,
                                 int32_t* out_query, int32_t* status, void* ws, size_t ws_bytes, mmsim_stream_t stream,
                                 int host_layout /* != 0: the workspace was filled through mmsim_knn_shard_host_f32 */);
/* mmsim_knn_shard_f32 for a shard whose rows live in HOST memory (page-locked g_host[ng*D]; g_stage[ng*D] is the device
 * buffer for the float32 copy): Q is on the device already.  The call with MMSIM_KNN_PHASE_PREP_G queues EVERY host->device
 * copy of the shard on an internal copy stream -- the 1/61 sample the pivot pre-pass needs first, then the rows split by
 * split, each split converted as it lands -- and returns; the later phases (PIVOT, then LADDER | TENSOR | RERANK after the
 * pivot exchange) wait only for the pieces they read, so the upload runs under the query preparation, the exchange of the
 * pivot lists and the sweeps of the earlier splits.  Workspace: mmsim_knn_host_workspace_bytes; all phases of one
 * retrieval must go through this entry point (the two layouts differ). */
MMSIM_API int mmsim_knn_shard_host_f32(const float* Q, int64_t nq, const float* g_host, float* g_stage, int64_t ng, int64_t D, int k,
                             int kp, int exclude_self, int64_t self_offset, float* out_dist, int32_t* out_idx, float* out_lb,
                             int32_t* status, void* ws, size_t ws_bytes, mmsim_stream_t stream, int phases, int64_t slice_rows,
                             int64_t slice_stride);
/* rows row_map[s] (s < min(*count, cap); count and row_map on the device) of out_dist / out_idx [*, k] = merge of the
 * `parts` compact lists of slot s (layout as mmsim_knn_merge with nq = cap) */
MMSIM_API int mmsim_knn_merge_patch(const float* dist_parts, const int32_t* idx_parts, int64_t part_stride, const int64_t* idx_base,
                          int parts, int cap, int k, const int32_t* count, const int32_t* row_map, float* out_dist,
                          int64_t* out_idx, mmsim_stream_t stream);

/* Semi-hard (FaceNet) negative mining -- the inner test of utils.select_triplets_facenet (src/utils.py:474-480) for m
 * (anchor, positive) pairs at once.  dist is the [n, n] distance matrix (row pitch ld, e.g. from mmsim_sqdist_f32),
 * labels[n] the int()-converted labels (:445), pairs[2*i], pairs[2*i+1] = anchor, positive.  Row j is a semi-hard
 * negative of pair i iff labels[j] != labels[anchor] (the NaN mask of :475), pos_dist < dist[anchor, j] and
 * fl32(dist[anchor, j] - pos_dist) < alpha (:477-478).  count[i] = how many there are (the reference's len(all_neg));
 * mask (nullable) gets the set itself: bit b of mask[i * ceil(n/32) + w] is row 32 w + b. */
MMSIM_API int mmsim_semihard_mask_f32(const float* dist, int64_t n, int64_t ld, const int32_t* labels, const int32_t* pairs,
                            int64_t m, float alpha, uint32_t* mask, int32_t* count, mmsim_stream_t stream);

/* The reference's all_neg[r] (src/utils.py:484) for p picks: picks[3*i .. 3*i+2] = anchor, positive, r;
 * neg_idx[i] = row of the r-th (0-based, ascending row order = np.where order) semi-hard negative, -1 if r >= count. */
MMSIM_API int mmsim_semihard_pick_f32(const float* dist, int64_t n, int64_t ld, const int32_t* labels, const int32_t* picks,
                            int64_t p, float alpha, int32_t* neg_idx, mmsim_stream_t stream);

/* Leave-one-out retrieval evaluation -- the loop body of utils.evaluate / utils.evaluate_simple
 * (src/utils.py:83-229) for the query rows listed in `queries` (the rows with label > 0, :114,171).
 * labels[N] are the raw int32 labels, cls[N] their dense class ids in [0, C).  Per query q (row i = queries[q]):
 *   ap[q]     float64 sklearn average precision of retrieve_one (:78-79), 0 with npos[q] == 0 when row i has no positive
 *   first[q]  rank (0-based) of the first ranked label equal to the query's (N-1 if none) -> recall_at_K (:257-266)
 *   depth[q], hist[q*C + c]   prefix length walked by precision_at_recall (:231-255) and per-class counts inside it
 *   rank[q*(N-1) + r]         (nullable) the full ranking in the row-i-deleted numbering, ordered by (distance, index)
 * aligned == 0 reproduces the reference's label lookup (full label array indexed by deleted-gallery positions);
 * aligned != 0 uses the deleted labels.  N <= 16385. */
MMSIM_API int mmsim_evaluate_f32(const float* E, const int32_t* labels, const int32_t* cls, int64_t N, int64_t D, int C,
                       const int32_t* queries, int64_t nq, double alpha, int aligned, double* ap, int32_t* npos,
                       int32_t* first, int32_t* depth, int32_t* hist, int32_t* rank, mmsim_stream_t stream);

/* The same records in the workspace form, for galleries of any size: register-tiled exact distances of a batch of queries
 * to all rows (the embeddings are read once per 32 queries, not once per query), then per query
 *   path 1  N <= 24,576: a stable radix sort of the N keys in shared memory + one streaming metrics pass, one CTA per query
 *   path 2  any N: a segmented device radix sort and one streaming pass per query
 * in batches that fit the caller's workspace (mmsim_evaluate_large_workspace_bytes covers both paths; 256-byte aligned).
 * Same arguments and results as mmsim_evaluate_f32.  mmsim_evaluate_ws_f32 takes the path (0 = path 1 when it fits, else 2);
 * mmsim_evaluate_large_f32 is path 2. */
#define MMSIM_EVAL_PATH_AUTO 0
#define MMSIM_EVAL_PATH_SMEM_SORT 1
#define MMSIM_EVAL_PATH_SEGMENTED_SORT 2
MMSIM_API int mmsim_evaluate_large_workspace_bytes(int64_t N, int64_t nq, size_t* bytes);
MMSIM_API int mmsim_evaluate_large_f32(const float* E, const int32_t* labels, const int32_t* cls, int64_t N, int64_t D, int C,
                             const int32_t* queries, int64_t nq, double alpha, int aligned, double* ap, int32_t* npos,
                             int32_t* first, int32_t* depth, int32_t* hist, int32_t* rank, void* workspace,
                             size_t workspace_bytes, mmsim_stream_t stream);
MMSIM_API int mmsim_evaluate_ws_f32(const float* E, const int32_t* labels, const int32_t* cls, int64_t N, int64_t D, int C,
                          const int32_t* queries, int64_t nq, double alpha, int aligned, double* ap, int32_t* npos,
                          int32_t* first, int32_t* depth, int32_t* hist, int32_t* rank, void* workspace, size_t workspace_bytes,
                          mmsim_stream_t stream, int path);

/* The confusion-matrix accumulation of utils.evaluate (src/utils.py:214-220) from the per-query records above:
 * cm[row] += (hist[q] / depth[q]).astype(float32) for every query q with npos[q] > 0 and class qcls[q] == row, in query
 * order (sequential float32 adds, bit-identical to the reference's loop); count[row] = number of such queries.
 * hist [nq, C], depth / npos / qcls [nq], cm [C, C] float32, count [C]; lists [nq] is scratch and list_off[C] the
 * exclusive prefix sum of the number of queries per class (where each class's query list starts in `lists`). */
MMSIM_API int mmsim_evaluate_confusion_f32(const int32_t* hist, const int32_t* depth, const int32_t* npos, const int32_t* qcls,
                                 int64_t nq, int C, float* cm, int32_t* count, int32_t* lists, const int32_t* list_off,
                                 mmsim_stream_t stream);

/* Embedding head: out[r] = l2_normalize(X[r] @ W + b) -- networks.CUBLayer.forward (src/networks.py:376-380, xw_plus_b)
 * followed by tf.nn.l2_normalize(logits, axis=-1, epsilon) (src/base_model_CUB.py:197-201): y * rsqrt(max(sum(y^2), epsilon)).
 * X [N, K], W [K, E], b [E] or NULL, out [N, E], all fp32 row-major; E <= 256; normalized == 0 skips the normalisation
 * (cfg.normalized false). */
MMSIM_API int mmsim_project_normalize_f32(const float* X, int64_t N, int64_t K, const float* W, const float* b, int64_t E,
                                int normalized, float epsilon, float* out, mmsim_stream_t stream);

/* tf.contrib's triplet_semihard_loss (metric_loss_ops.triplet_semihard_loss(labels, embeddings, margin),
 * src/base_CUB.py:163-166), forward + backward: loss[0] = mean over (anchor, positive) pairs of
 * max(margin + D_ap - n*, 0) with n* the smallest negative distance beyond D_ap, else the largest negative distance
 * (squared distances); dE (nullable) = d loss / d E.  labels int32 [N], E [N, D] fp32; N <= 8192.  The TF function is a
 * third-party dependency of the reference that is neither vendored nor installable here: parity unpinned. */
MMSIM_API int mmsim_triplet_semihard_workspace_bytes(int64_t N, size_t* bytes);
MMSIM_API int mmsim_triplet_semihard_f32(const float* E, const int32_t* labels, int64_t N, int64_t D, float margin, float* loss,
                               float* dE, void* workspace, size_t workspace_bytes, mmsim_stream_t stream);

/* tf.contrib's lifted_struct_loss (metric_loss_ops.lifted_struct_loss(labels, embeddings, margin), src/base_CUB.py:167-171),
 * forward + backward: loss[0] = 0.25 * sum over ordered positive pairs of max(log(S_a + S_b) + D_ab, 0)^2 / (P / 2) with
 * S_a = sum over a's negatives of exp(margin - D_aj), Euclidean (not squared) distances; dE (nullable) = d loss / d E.
 * Same arguments as mmsim_triplet_semihard_f32; third-party function, parity unpinned. */
MMSIM_API int mmsim_lifted_struct_workspace_bytes(int64_t N, size_t* bytes);
MMSIM_API int mmsim_lifted_struct_f32(const float* E, const int32_t* labels, int64_t N, int64_t D, float margin, float* loss,
                            float* dE, void* workspace, size_t workspace_bytes, mmsim_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMSIM_H_ */
