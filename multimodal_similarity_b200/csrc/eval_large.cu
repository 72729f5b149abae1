// K5b: leave-one-out retrieval evaluation in the workspace form -- the same per-query records as eval.cu (K5), i.e. the
// loop body of
//   utils.evaluate / utils.evaluate_simple            src/utils.py:83-229
// for a batch of B queries at a time:
//   1. eval_tile_dist_kernel (eval_fast.cu)   exact fp32 distances of B queries to all N rows -> [B][N] key bits, the
//                                             query's own row as 0xffffffff (sorts last)
//   2a. N <= 24,576: eval_sort_metrics_kernel (eval_fast.cu) -- radix sort in shared memory + metrics, one CTA per query
//   2b. larger N   : a segmented radix sort of every row by key bits (stable, rows start ascending => ordered by
//                    (distance, index)); the sort is CUB's DeviceSegmentedRadixSort -- a library call for a plain sort, like
//                    cuBLAS for a plain GEMM -- then eval_metrics_kernel: one CTA per query streams its sorted row once
//                    (eval_metrics.cuh)
// SURVEY.md 8(f) row 2.  Path 2b is HBM-bound: the sort moves 4 passes x 16 bytes per (query, row) pair.
#include <cuda_runtime.h>

#include <algorithm>

#include <cub/device/device_segmented_radix_sort.cuh>

#include "common.cuh"
#include "eval.h"
#include "eval_metrics.cuh"

namespace mmsim {
namespace eval {

constexpr int LT = 1024;   // threads of the metrics kernel

__global__ void seg_offsets_kernel(int* __restrict__ off, int b, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= b) off[i] = i * n;
}

__global__ void __launch_bounds__(LT)
eval_metrics_kernel(const uint32_t* __restrict__ sd, const int* __restrict__ sv, const int* __restrict__ labels,
                    const int* __restrict__ cls, int N, int C, const int* __restrict__ queries, double alpha, int aligned,
                    double* __restrict__ out_ap, int* __restrict__ out_npos, int* __restrict__ out_first,
                    int* __restrict__ out_depth, int* __restrict__ out_hist, int* __restrict__ out_rank) {
  extern __shared__ int lhist[];                       // [C]
  const int qn = blockIdx.x;
  loo_metrics<LT, int>(sd + size_t(qn) * N, sv + size_t(qn) * N, labels, cls, N, C, queries[qn], qn, alpha, aligned, lhist, out_ap,
                       out_npos, out_first, out_depth, out_hist, out_rank);
}

static size_t sort_temp_bytes(int64_t items, int segs) {
  size_t bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                           static_cast<const int*>(nullptr), static_cast<int*>(nullptr), items, segs,
                                           static_cast<const int*>(nullptr), static_cast<const int*>(nullptr));
  return bytes;
}

constexpr int64_t kMaxBatchItems = int64_t(1) << 30;   // (query, row) pairs per batch: int offsets, bounded workspace

static int batch_rows_segmented(int64_t N, int64_t nq) {
  return int(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nq, 1024), kMaxBatchItems / N / 4)));
}
static int64_t key_pitch(int64_t N) { return (N + 31) / 32 * 32; }
static int batch_rows_sorted(int64_t N, int64_t nq) {   // shared-memory sort path: one [B][pitch] block of keys, <= 256 MB
  const int64_t cap = std::max<int64_t>(32, (int64_t(1) << 26) / key_pitch(N) / 32 * 32);
  return int(std::min<int64_t>(std::max<int64_t>(nq, 1), cap));
}
static size_t segmented_bytes(int64_t N, int64_t nq) {
  const int B = batch_rows_segmented(N, nq);
  const size_t items = size_t(B) * size_t(N);
  return align_up(items * 4, 256) * 4 + align_up(size_t(B + 1) * 4, 256) + align_up(sort_temp_bytes(int64_t(items), B), 256) + 1024;
}

int large_workspace_bytes(int64_t N, int64_t nq, size_t* out) {
  MMSIM_REQUIRE(out && N >= 2 && nq >= 0 && N < (int64_t(1) << 28), MMSIM_ERR_ARG, "evaluate_large: bad sizes");
  size_t need = segmented_bytes(N, nq);
  if (N <= kSortMaxN) need = std::max(need, align_up(size_t(batch_rows_sorted(N, nq)) * size_t(key_pitch(N)) * 4, 256) + 1024);
  *out = need;
  return MMSIM_OK;
}

int run_large(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
              double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, void* ws,
              size_t ws_bytes, cudaStream_t s, int path) {
  MMSIM_REQUIRE(E && labels && cls && queries && ap && npos && first && depth && hist && ws, MMSIM_ERR_ARG,
                "evaluate_large: null pointer argument");
  MMSIM_REQUIRE(N >= 2 && N < (int64_t(1) << 28) && D >= 1 && D <= 4096 && C >= 1 && C <= 8192 && nq >= 0, MMSIM_ERR_ARG,
                "evaluate_large: bad sizes N=%lld D=%lld C=%d", (long long)N, (long long)D, C);
  MMSIM_REQUIRE(path == kPathAuto || path == kPathSmemSort || path == kPathSegmentedSort, MMSIM_ERR_ARG,
                "evaluate_large: path must be 0 (auto), 1 (shared-memory sort) or 2 (segmented sort)");
  size_t need = 0;
  large_workspace_bytes(N, nq, &need);
  MMSIM_REQUIRE(ws_bytes >= need, MMSIM_ERR_WORKSPACE, "evaluate_large: workspace too small (%zu < %zu)", ws_bytes, need);
  MMSIM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MMSIM_ERR_WORKSPACE, "evaluate_large: workspace must be 256-byte aligned");
  const bool fits = sort_path_fits(N, C);
  MMSIM_REQUIRE(path != kPathSmemSort || fits, MMSIM_ERR_UNSUPPORTED,
                "evaluate_large: N=%lld, C=%d does not fit the shared-memory sort (N <= %d)", (long long)N, C, kSortMaxN);
  if (nq == 0) return MMSIM_OK;
  const int n = int(N - 1);
  uint8_t* w = static_cast<uint8_t*>(ws);

  if (fits && path != kPathSegmentedSort) {
    const int B = batch_rows_sorted(N, nq);
    const int64_t ldk = key_pitch(N);
    uint32_t* keys = reinterpret_cast<uint32_t*>(w);
    for (int64_t q0 = 0; q0 < nq; q0 += B) {
      const int b = int(std::min<int64_t>(B, nq - q0));
      int rc = launch_distances(E, N, D, queries + q0, b, keys, ldk, nullptr, s);
      if (rc != MMSIM_OK) return rc;
      rc = launch_sort_metrics(keys, ldk, labels, cls, N, C, queries + q0, b, alpha, aligned, ap + q0, npos + q0, first + q0,
                               depth + q0, hist + size_t(q0) * C, rank ? rank + size_t(q0) * n : nullptr, s);
      if (rc != MMSIM_OK) return rc;
    }
    return MMSIM_OK;
  }

  const int B = batch_rows_segmented(N, nq);
  const size_t items = size_t(B) * size_t(N), seg = align_up(items * 4, 256);
  uint32_t* k_in = reinterpret_cast<uint32_t*>(w);
  uint32_t* k_out = reinterpret_cast<uint32_t*>(w + seg);
  int* v_in = reinterpret_cast<int*>(w + 2 * seg);
  int* v_out = reinterpret_cast<int*>(w + 3 * seg);
  int* off = reinterpret_cast<int*>(w + 4 * seg);
  void* tmp = w + 4 * seg + align_up(size_t(B + 1) * 4, 256);
  size_t tmp_bytes = sort_temp_bytes(int64_t(items), B);
  seg_offsets_kernel<<<(B + 256) / 256, 256, 0, s>>>(off, B, int(N));
  MMSIM_CUDA_CHECK(::mmsim::launched());
  for (int64_t q0 = 0; q0 < nq; q0 += B) {
    const int b = int(std::min<int64_t>(B, nq - q0));
    const int rc = launch_distances(E, N, D, queries + q0, b, k_in, N, v_in, s);
    if (rc != MMSIM_OK) return rc;
    MMSIM_CUDA_CHECK(cub::DeviceSegmentedRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, int64_t(b) * N, b, off,
                                                              off + 1, 0, 32, s));
    eval_metrics_kernel<<<b, LT, size_t(C) * 4, s>>>(k_out, v_out, labels, cls, int(N), C, queries + q0, alpha, aligned, ap + q0,
                                                     npos + q0, first + q0, depth + q0, hist + size_t(q0) * C,
                                                     rank ? rank + size_t(q0) * n : nullptr);
    MMSIM_CUDA_CHECK(::mmsim::launched());
  }
  return MMSIM_OK;
}

}  // namespace eval
}  // namespace mmsim
