// extern "C" surface of libmmsim.so (declared in include/mmsim.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/mmsim.h"
#include "common.cuh"
#include "eval.h"
#include "knn.h"
#include "loss.h"
#include "merge.h"
#include "mining.h"
#include "project.h"
#include "lifted_struct.h"
#include "semihard_loss.h"
#include "sqdist.h"

namespace mmsim {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
cudaError_t launched() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
}  // namespace mmsim

using namespace mmsim;

extern "C" {

MMSIM_API int mmsim_version(void) { return 100; }  // 0.1.0

MMSIM_API const char* mmsim_last_error(void) { return g_err; }

MMSIM_API int64_t mmsim_kernel_launches(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

MMSIM_API int mmsim_sqdist_f32(const float* A, int64_t M, const float* B, int64_t N, int64_t D, int metric, float* out, int64_t ld,
                     mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_sqdist_f32");
  return sqdist::run(A, M, B, N, D, metric, out, ld, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_loss_workspace_bytes(int64_t N, int64_t D, size_t* bytes) {
  MMSIM_REQUIRE(bytes, MMSIM_ERR_ARG, "loss_workspace_bytes: null output");
  MMSIM_REQUIRE(N >= 1 && N <= 1024 && D >= 1 && D <= 512, MMSIM_ERR_UNSUPPORTED,
                "loss: N=%lld (1..1024), D=%lld (1..512) unsupported", (long long)N, (long long)D);
  *bytes = loss::make_layout(N, D).total_bytes;
  return MMSIM_OK;
}

MMSIM_API int mmsim_loss_f32(int kind, const float* E, const float* pids, int64_t N, int64_t D, int soft, float margin, int weighted,
                   float* loss_out, float* num_active, float* diff, float* weights, float* furthest_positive,
                   float* closest_negative, int32_t* pos_idx, int32_t* neg_idx, float* dE, void* ws, size_t ws_bytes,
                   mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_loss_f32");
  return loss::run(kind, E, pids, N, D, soft, margin, weighted, loss_out, num_active, diff, weights, furthest_positive,
                   closest_negative, pos_idx, neg_idx, dE, ws, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

static int knn_ws_bytes(const char* what, int64_t nq, int64_t ng, int64_t D, int k, bool host_mode, size_t* bytes) {
  MMSIM_REQUIRE(bytes, MMSIM_ERR_ARG, "%s: null output", what);
  MMSIM_REQUIRE(nq > 0 && ng > 0 && D > 0 && D <= 256 && k >= 1, MMSIM_ERR_ARG, "%s: bad sizes", what);
  // the layout is sized for the largest grid any device could use, so the answer does not depend on the device
  size_t worst = 0;
  for (int sms = 1; sms <= 192; ++sms) worst = std::max(worst, knn::make_plan(nq, ng, D, k, sms, host_mode).total_bytes);
  *bytes = worst;
  return MMSIM_OK;
}

MMSIM_API int mmsim_knn_workspace_bytes(int64_t nq, int64_t ng, int64_t D, int k, size_t* bytes) {
  return knn_ws_bytes("knn_workspace_bytes", nq, ng, D, k, false, bytes);
}

MMSIM_API int mmsim_knn_host_workspace_bytes(int64_t nq, int64_t ng, int64_t D, int k, size_t* bytes) {
  return knn_ws_bytes("knn_host_workspace_bytes", nq, ng, D, k, true, bytes);
}

MMSIM_API int mmsim_knn_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                  int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status, void* ws, size_t ws_bytes,
                  mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_f32");
  return knn::run(Q, nq, G, ng, D, k, exclude_self, self_offset, out_dist, out_idx, status, ws, ws_bytes,
                  reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_knn_f32_phases(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                         int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status, void* ws, size_t ws_bytes,
                         mmsim_stream_t stream, int phases) {
  MMSIM_RANGE("mmsim_knn_f32_phases");
  MMSIM_REQUIRE(phases > 0 && phases <= knn::kPhaseAll, MMSIM_ERR_ARG, "knn_phases: phases must be a mask in 1..63");
  return knn::run(Q, nq, G, ng, D, k, exclude_self, self_offset, out_dist, out_idx, status, ws, ws_bytes,
                  reinterpret_cast<cudaStream_t>(stream), phases);
}

MMSIM_API int mmsim_knn_host_f32(const float* q_host, int64_t nq, const float* g_host, int64_t ng, int64_t D, int k,
                       int exclude_self, int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status,
                       float* q_stage, float* g_stage, void* ws, size_t ws_bytes, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_host_f32");
  MMSIM_REQUIRE(q_host && g_host && q_stage && g_stage, MMSIM_ERR_ARG, "knn_host: null host or staging pointer");
  const knn::HostPipe hp{q_host, g_host};
  return knn::run(q_stage, nq, g_stage, ng, D, k, exclude_self, self_offset, out_dist, out_idx, status, ws, ws_bytes,
                  reinterpret_cast<cudaStream_t>(stream), knn::kPhaseAll, 0, nullptr, &hp);
}

MMSIM_API int mmsim_knn_finish_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                         int64_t self_offset, float* out_dist, int32_t* out_idx, int32_t* status, void* ws, size_t ws_bytes,
                         mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_finish_f32");
  return knn::run(Q, nq, G, ng, D, k, exclude_self, self_offset, out_dist, out_idx, status, ws, ws_bytes,
                  reinterpret_cast<cudaStream_t>(stream), knn::kPhaseFinish);
}

MMSIM_API int mmsim_knn_shard_fallback_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self,
                                 int64_t self_offset, const float* flag, int cap, float* out_dist, int32_t* out_idx,
                                 int32_t* out_query, int32_t* status, void* ws, size_t ws_bytes, mmsim_stream_t stream,
                                 int host_layout) {
  MMSIM_RANGE("mmsim_knn_shard_fallback_f32");
  return knn::shard_fallback(Q, nq, G, ng, D, k, exclude_self, self_offset, flag, cap, out_dist, out_idx, out_query, status, ws,
                             ws_bytes, reinterpret_cast<cudaStream_t>(stream), host_layout != 0);
}

MMSIM_API int mmsim_knn_merge_patch(const float* dist_parts, const int32_t* idx_parts, int64_t part_stride, const int64_t* idx_base,
                          int parts, int cap, int k, const int32_t* count, const int32_t* row_map, float* out_dist,
                          int64_t* out_idx, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_merge_patch");
  MMSIM_REQUIRE(count && row_map, MMSIM_ERR_ARG, "knn_merge_patch: null pointer argument");
  return merge::run(dist_parts, idx_parts, part_stride, idx_base, parts, cap, k, k, nullptr, 0, out_dist, out_idx, nullptr,
                    reinterpret_cast<cudaStream_t>(stream), nullptr, count, row_map);
}

MMSIM_API int mmsim_knn_merge(const float* dist_parts, const int32_t* idx_parts, int64_t part_stride, const int64_t* idx_base,
                    int parts, int64_t nq, int k, float* out_dist, int64_t* out_idx, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_merge");
  return merge::run(dist_parts, idx_parts, part_stride, idx_base, parts, nq, k, k, nullptr, 0, out_dist, out_idx, nullptr,
                    reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_evaluate_f32(const float* E, const int32_t* labels, const int32_t* cls, int64_t N, int64_t D, int C,
                       const int32_t* queries, int64_t nq, double alpha, int aligned, double* ap, int32_t* npos,
                       int32_t* first, int32_t* depth, int32_t* hist, int32_t* rank, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_evaluate_f32");
  return eval::run(E, labels, cls, N, D, C, queries, nq, alpha, aligned, ap, npos, first, depth, hist, rank,
                   reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_knn_shard_f32(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int kp, int exclude_self,
                        int64_t self_offset, float* out_dist, int32_t* out_idx, float* out_lb, int32_t* status, void* ws,
                        size_t ws_bytes, mmsim_stream_t stream, int phases, int64_t slice_rows, int64_t slice_stride) {
  MMSIM_RANGE("mmsim_knn_shard_f32");
  MMSIM_REQUIRE(phases > 0 && (phases & ~(knn::kPhaseAll | knn::kPhasePrepQ | knn::kPhasePrepG)) == 0, MMSIM_ERR_ARG,
                "knn_shard: phases must be a mask of MMSIM_KNN_PHASE_*");
  MMSIM_REQUIRE(kp >= 1 && kp <= knn::KP, MMSIM_ERR_ARG, "knn_shard: kp must be in 1..%d", knn::KP);
  return knn::run(Q, nq, G, ng, D, k, exclude_self, self_offset, out_dist, out_idx, status, ws, ws_bytes,
                  reinterpret_cast<cudaStream_t>(stream), phases, kp, out_lb, nullptr, slice_rows, slice_stride);
}

MMSIM_API int mmsim_knn_shard_host_f32(const float* Q, int64_t nq, const float* g_host, float* g_stage, int64_t ng, int64_t D, int k,
                             int kp, int exclude_self, int64_t self_offset, float* out_dist, int32_t* out_idx, float* out_lb,
                             int32_t* status, void* ws, size_t ws_bytes, mmsim_stream_t stream, int phases, int64_t slice_rows,
                             int64_t slice_stride) {
  MMSIM_RANGE("mmsim_knn_shard_host_f32");
  MMSIM_REQUIRE(g_host && g_stage, MMSIM_ERR_ARG, "knn_shard_host: null host gallery or staging pointer");
  MMSIM_REQUIRE(phases > 0 && (phases & ~(knn::kPhaseAll | knn::kPhasePrepQ | knn::kPhasePrepG)) == 0, MMSIM_ERR_ARG,
                "knn_shard_host: phases must be a mask of MMSIM_KNN_PHASE_*");
  MMSIM_REQUIRE(kp >= 1 && kp <= knn::KP, MMSIM_ERR_ARG, "knn_shard_host: kp must be in 1..%d", knn::KP);
  const knn::HostPipe hp{nullptr, g_host};
  return knn::run(Q, nq, g_stage, ng, D, k, exclude_self, self_offset, out_dist, out_idx, status, ws, ws_bytes,
                  reinterpret_cast<cudaStream_t>(stream), phases, kp, out_lb, &hp, slice_rows, slice_stride);
}

MMSIM_API int mmsim_knn_pivot_region(int64_t nq, int64_t ng, int64_t D, int k, size_t* offset, size_t* bytes) {
  MMSIM_REQUIRE(offset && bytes, MMSIM_ERR_ARG, "knn_pivot_region: null output");
  MMSIM_REQUIRE(nq > 0 && ng > 0 && D > 0 && D <= 256 && k >= 1, MMSIM_ERR_ARG, "knn_pivot_region: bad sizes");
  const knn::Plan p = knn::make_plan(nq, ng, D, k, 148);   // the region does not depend on the SM count
  *offset = p.off_piv16;
  *bytes = size_t(p.n_qblocks) * 128 * knn::kPivotsPerRow * sizeof(float);
  return MMSIM_OK;
}

MMSIM_API int mmsim_knn_plan(int64_t nq, int64_t ng, int64_t D, int k, int num_sms, int64_t* out, int n_out) {
  MMSIM_REQUIRE(out && n_out >= 12, MMSIM_ERR_ARG, "knn_plan: need an output array of at least 12 int64");
  MMSIM_REQUIRE(nq > 0 && ng > 0 && D > 0 && D <= 256 && k >= 1 && num_sms >= 1, MMSIM_ERR_ARG, "knn_plan: bad sizes");
  const knn::Plan p = knn::make_plan(nq, ng, D, k, num_sms);
  // [12..] workspace offsets for tests that inspect a finished call (tests/test_gpu_certificate.py): candidate-log counters,
  // thresholds and entries, the fp16 operand copies and the gallery norm pack
  const int64_t v[21] = {p.Dp, p.katoms, p.n_qblocks, p.n_tiles, p.n_splits, p.tiles_per_split, p.grid, p.logcap,
                         p.use_pivots, p.n_sample_tiles, p.n_sample, int64_t(p.total_bytes),
                         int64_t(p.off_log_cnt), int64_t(p.off_log_tau), p.sweepq, p.host_splits, int64_t(p.off_log),
                         int64_t(p.off_qh), int64_t(p.off_gh), int64_t(p.off_gpack), p.n_anchor};
  for (int i = 0; i < 21 && i < n_out; ++i) out[i] = v[i];
  return MMSIM_OK;
}

MMSIM_API int mmsim_knn_merge_pivots(const float* parts, int nparts, int64_t part_stride, int64_t rows, float* out,
                           mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_merge_pivots");
  return knn::merge_pivots(parts, nparts, part_stride, rows, out, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_knn_merge_certified(const float* dist_parts, const int32_t* idx_parts, int64_t part_stride,
                              const int64_t* idx_base, int parts, int64_t nq, int k_in, int k, const float* lb_parts,
                              int64_t lb_stride, float* out_dist, void* out_idx, int out_idx_bits, int32_t* status,
                              float* out_flag, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_knn_merge_certified");
  MMSIM_REQUIRE(out_idx_bits == 32 || out_idx_bits == 64, MMSIM_ERR_ARG, "knn_merge_certified: out_idx_bits must be 32 or 64");
  return merge::run(dist_parts, idx_parts, part_stride, idx_base, parts, nq, k_in, k, lb_parts, lb_stride, out_dist, out_idx,
                    status, reinterpret_cast<cudaStream_t>(stream), out_flag, nullptr, nullptr, out_idx_bits == 32);
}

MMSIM_API int mmsim_semihard_mask_f32(const float* dist, int64_t n, int64_t ld, const int32_t* labels, const int32_t* pairs,
                            int64_t m, float alpha, uint32_t* mask, int32_t* count, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_semihard_mask_f32");
  return mining::run_mask(dist, n, ld, labels, pairs, m, alpha, mask, count, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_semihard_pick_f32(const float* dist, int64_t n, int64_t ld, const int32_t* labels, const int32_t* picks,
                            int64_t p, float alpha, int32_t* neg_idx, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_semihard_pick_f32");
  return mining::run_pick(dist, n, ld, labels, picks, p, alpha, neg_idx, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_evaluate_large_workspace_bytes(int64_t N, int64_t nq, size_t* bytes) {
  return eval::large_workspace_bytes(N, nq, bytes);
}

MMSIM_API int mmsim_evaluate_large_f32(const float* E, const int32_t* labels, const int32_t* cls, int64_t N, int64_t D, int C,
                             const int32_t* queries, int64_t nq, double alpha, int aligned, double* ap, int32_t* npos,
                             int32_t* first, int32_t* depth, int32_t* hist, int32_t* rank, void* workspace,
                             size_t workspace_bytes, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_evaluate_large_f32");
  return eval::run_large(E, labels, cls, N, D, C, queries, nq, alpha, aligned, ap, npos, first, depth, hist, rank, workspace,
                         workspace_bytes, reinterpret_cast<cudaStream_t>(stream), eval::kPathSegmentedSort);
}

MMSIM_API int mmsim_evaluate_ws_f32(const float* E, const int32_t* labels, const int32_t* cls, int64_t N, int64_t D, int C,
                          const int32_t* queries, int64_t nq, double alpha, int aligned, double* ap, int32_t* npos,
                          int32_t* first, int32_t* depth, int32_t* hist, int32_t* rank, void* workspace, size_t workspace_bytes,
                          mmsim_stream_t stream, int path) {
  MMSIM_RANGE("mmsim_evaluate_ws_f32");
  return eval::run_large(E, labels, cls, N, D, C, queries, nq, alpha, aligned, ap, npos, first, depth, hist, rank, workspace,
                         workspace_bytes, reinterpret_cast<cudaStream_t>(stream), path);
}

MMSIM_API int mmsim_evaluate_confusion_f32(const int32_t* hist, const int32_t* depth, const int32_t* npos, const int32_t* qcls,
                                 int64_t nq, int C, float* cm, int32_t* count, int32_t* lists, const int32_t* list_off,
                                 mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_evaluate_confusion_f32");
  return eval::confusion(hist, depth, npos, qcls, nq, C, cm, count, lists, list_off, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_project_normalize_f32(const float* X, int64_t N, int64_t K, const float* W, const float* b, int64_t E,
                                int normalized, float epsilon, float* out, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_project_normalize_f32");
  return project::run(X, N, K, W, b, E, normalized, epsilon, out, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_triplet_semihard_workspace_bytes(int64_t N, size_t* bytes) { return semihard_loss::workspace_bytes(N, bytes); }

MMSIM_API int mmsim_triplet_semihard_f32(const float* E, const int32_t* labels, int64_t N, int64_t D, float margin, float* loss,
                               float* dE, void* workspace, size_t workspace_bytes, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_triplet_semihard_f32");
  return semihard_loss::run(E, labels, N, D, margin, loss, dE, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

MMSIM_API int mmsim_lifted_struct_workspace_bytes(int64_t N, size_t* bytes) { return lifted_struct::workspace_bytes(N, bytes); }

MMSIM_API int mmsim_lifted_struct_f32(const float* E, const int32_t* labels, int64_t N, int64_t D, float margin, float* loss,
                            float* dE, void* workspace, size_t workspace_bytes, mmsim_stream_t stream) {
  MMSIM_RANGE("mmsim_lifted_struct_f32");
  return lifted_struct::run(E, labels, N, D, margin, loss, dE, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
