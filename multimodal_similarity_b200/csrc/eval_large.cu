// K5b: leave-one-out retrieval evaluation at gallery scale (N > 16385, where one query's full ranking no longer fits in
// shared memory) -- the same per-query records as eval.cu (K5), i.e. the loop body of
//   utils.evaluate / utils.evaluate_simple            src/utils.py:83-229
// for a batch of B queries at a time:
//   1. eval_dist_kernel     exact fp32 distances (NumPy arithmetic, exact.cuh) of B queries to the other N-1 rows, in the
//                           row-i-deleted numbering  -> [B][N-1] keys + positions
//   2. segmented radix sort of every row by distance; stable, positions start ascending => ordered by (distance, index)
//                           exactly like K5's bitonic sort.  The sort is CUB's DeviceSegmentedRadixSort -- a library
//                           call for a plain sort, like cuBLAS for a plain GEMM; kernels 1 and 3 are ours.
//   3. eval_metrics_kernel  one CTA per query streams its sorted row once: sklearn-compatible AP (thresholds at distinct
//                           scores fl32(max(dist) - dist), ties grouped), first-match rank, precision@recall depth, then
//                           the per-class histogram of the walked prefix
// SURVEY.md 8(f) row 2.  HBM-bound: the sort moves 4 passes x 16 bytes per (query, row) pair.
#include <cuda_runtime.h>

#include <cub/device/device_segmented_radix_sort.cuh>

#include "common.cuh"
#include "eval.h"
#include "exact.cuh"

namespace mmsim {
namespace eval {

constexpr int LT = 1024;   // threads of the metrics kernel

__global__ void __launch_bounds__(256)
eval_dist_kernel(const float* __restrict__ E, int N, int D, const int* __restrict__ queries, float* __restrict__ keys,
                 int* __restrict__ vals) {
  extern __shared__ float dq[];
  const int i = queries[blockIdx.y], n = N - 1;
  for (int c = threadIdx.x; c < D; c += blockDim.x) dq[c] = E[size_t(i) * D + c];
  __syncthreads();
  for (int jp = blockIdx.x * blockDim.x + threadIdx.x; jp < n; jp += gridDim.x * blockDim.x) {
    const int j = jp + (jp >= i ? 1 : 0);
    keys[size_t(blockIdx.y) * n + jp] = exact_l2(dq, E + size_t(j) * D, D);
    vals[size_t(blockIdx.y) * n + jp] = jp;
  }
}

__global__ void seg_offsets_kernel(int* __restrict__ off, int b, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= b) off[i] = i * n;
}

// block-wide inclusive sum scan / exclusive max scan of one int per thread (LT threads); `carry` is added / maxed in and
// the block total (sum) or block maximum (max) comes back through it
__device__ __forceinline__ int block_scan_sum(int x, int* wsum, int& carry) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const int out = incl + (warp ? wsum[warp - 1] : 0) + carry;
  const int total = wsum[31];
  __syncthreads();
  carry += total;
  return out;
}
__device__ __forceinline__ int block_scan_max_excl(int x, int* wmax, int& carry) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = max(incl, y);
  }
  if (lane == 31) wmax[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wmax[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w = max(w, y);
    }
    wmax[lane] = w;
  }
  __syncthreads();
  int excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
  excl = max(excl, max(warp ? wmax[warp - 1] : 0, carry));
  const int total = wmax[31];
  __syncthreads();
  carry = max(carry, total);
  return excl;
}

__global__ void __launch_bounds__(LT)
eval_metrics_kernel(const float* __restrict__ sd, const int* __restrict__ sv, const int* __restrict__ labels,
                    const int* __restrict__ cls, int N, int C, const int* __restrict__ queries, double alpha, int aligned,
                    double* __restrict__ out_ap, int* __restrict__ out_npos, int* __restrict__ out_first,
                    int* __restrict__ out_depth, int* __restrict__ out_hist, int* __restrict__ out_rank) {
  extern __shared__ int lhist[];                       // [C]
  __shared__ int wbuf[32];
  __shared__ double dred[LT / 32];
  __shared__ int s_first, s_depth, s_cnt, s_m0;
  const int t = threadIdx.x, qn = blockIdx.x;
  const int i = queries[qn], n = N - 1, ql = labels[i];
  const float* d = sd + size_t(qn) * n;
  const int* v = sv + size_t(qn) * n;
  for (int c = t; c < C; c += LT) lhist[c] = 0;
  if (t == 0) { s_first = n; s_depth = n; s_cnt = 0; s_m0 = 0; }
  __syncthreads();

  // how many ranked labels equal the query's, as the reference forms them (src/utils.py:185-186: the FULL label array
  // indexed by positions of the gallery-with-row-i-deleted unless aligned): independent of the ranking
  {
    int c = 0;
    for (int jp = t; jp < n; jp += LT) c += labels[aligned ? jp + (jp >= i ? 1 : 0) : jp] == ql ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((t & 31) == 0) atomicAdd(&s_cnt, c);
  }
  __syncthreads();
  const int target = int(alpha * double(s_cnt));
  const float dmax = n > 0 ? d[n - 1] : 0.f;

  int carry_pos = 0, carry_m = 0, carry_end = 0;
  double acc = 0.0;
  int my_first = n;
  for (int base = 0; base < n; base += LT) {
    const int r = base + t;
    const bool in = r < n;
    int pos = 0, m = 0;
    bool is_end = false;
    if (in) {
      const int jp = v[r];
      if (out_rank) out_rank[size_t(qn) * n + r] = jp;
      const int jt = jp + (jp >= i ? 1 : 0);
      pos = labels[jt] == ql ? 1 : 0;                                     // AP uses the correctly deleted labels (:78)
      m = labels[aligned ? jt : jp] == ql ? 1 : 0;
      const float s = __fsub_rn(dmax, d[r]);                              // score = max(dist) - dist in float32 (:79)
      is_end = r == n - 1 || __fsub_rn(dmax, d[r + 1]) != s;             // last rank of its tie group
      if (m) my_first = min(my_first, r);
      if (r == 0 && m) s_m0 = 1;
    }
    const int cp = block_scan_sum(pos, wbuf, carry_pos);                  // positives among ranks 0..r
    const int cm = block_scan_sum(m, wbuf, carry_m);
    const int prev_end = block_scan_max_excl(is_end ? cp : 0, wbuf, carry_end);   // positives up to the previous group end
    if (in && is_end && cp > prev_end) acc += double(cp - prev_end) * (double(cp) / double(r + 1));
    if (in && target > 0 && m && cm == target) s_depth = r + 1;           // exactly one rank satisfies this
  }
  // reductions
  my_first = __reduce_min_sync(0xffffffffu, my_first);
  if ((t & 31) == 0) atomicMin(&s_first, my_first);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((t & 31) == 0) dred[t >> 5] = acc;
  __syncthreads();
  const int npos = carry_pos;
  if (t == 0) {
    double a = 0.0;
    for (int w = 0; w < LT / 32; ++w) a += dred[w];
    out_ap[qn] = npos > 0 ? a / double(npos) : 0.0;
    out_npos[qn] = npos;
    if (target == 0) s_depth = s_m0 ? n : min(n, 1);                     // src/utils.py:240-250 with int(alpha * #pos) == 0
  }
  __syncthreads();
  const int depth = s_depth;
  for (int r = t; r < depth; r += LT) {
    const int jp = v[r];
    atomicAdd(&lhist[cls[aligned ? jp + (jp >= i ? 1 : 0) : jp]], 1);
  }
  __syncthreads();
  for (int c = t; c < C; c += LT) out_hist[size_t(qn) * C + c] = lhist[c];
  if (t == 0) {
    out_first[qn] = s_first;
    out_depth[qn] = depth;
  }
}

static size_t sort_temp_bytes(int64_t items, int segs) {
  size_t bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, static_cast<const float*>(nullptr), static_cast<float*>(nullptr),
                                           static_cast<const int*>(nullptr), static_cast<int*>(nullptr), items, segs,
                                           static_cast<const int*>(nullptr), static_cast<const int*>(nullptr));
  return bytes;
}

constexpr int64_t kMaxBatchItems = int64_t(1) << 30;   // (query, row) pairs per batch: int offsets, bounded workspace

static int batch_rows(int64_t N, int64_t nq) {
  const int64_t n = N - 1;
  return int(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nq, 1024), kMaxBatchItems / n / 4)));
}

int large_workspace_bytes(int64_t N, int64_t nq, size_t* out) {
  MMSIM_REQUIRE(out && N >= 2 && nq >= 0 && N < (int64_t(1) << 28), MMSIM_ERR_ARG, "evaluate_large: bad sizes");
  const int B = batch_rows(N, nq);
  const size_t items = size_t(B) * size_t(N - 1);
  *out = align_up(items * 4, 256) * 4 + align_up(size_t(B + 1) * 4, 256) + align_up(sort_temp_bytes(int64_t(items), B), 256) + 1024;
  return MMSIM_OK;
}

int run_large(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
              double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, void* ws,
              size_t ws_bytes, cudaStream_t s) {
  MMSIM_REQUIRE(E && labels && cls && queries && ap && npos && first && depth && hist && ws, MMSIM_ERR_ARG,
                "evaluate_large: null pointer argument");
  MMSIM_REQUIRE(N >= 2 && N < (int64_t(1) << 28) && D >= 1 && D <= 4096 && C >= 1 && C <= 8192 && nq >= 0, MMSIM_ERR_ARG,
                "evaluate_large: bad sizes N=%lld D=%lld C=%d", (long long)N, (long long)D, C);
  size_t need = 0;
  large_workspace_bytes(N, nq, &need);
  MMSIM_REQUIRE(ws_bytes >= need, MMSIM_ERR_WORKSPACE, "evaluate_large: workspace too small (%zu < %zu)", ws_bytes, need);
  MMSIM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MMSIM_ERR_WORKSPACE, "evaluate_large: workspace must be 256-byte aligned");
  if (nq == 0) return MMSIM_OK;
  const int B = batch_rows(N, nq);
  const int n = int(N - 1);
  const size_t items = size_t(B) * n, seg = align_up(items * 4, 256);
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* k_in = reinterpret_cast<float*>(w);
  float* k_out = reinterpret_cast<float*>(w + seg);
  int* v_in = reinterpret_cast<int*>(w + 2 * seg);
  int* v_out = reinterpret_cast<int*>(w + 3 * seg);
  int* off = reinterpret_cast<int*>(w + 4 * seg);
  void* tmp = w + 4 * seg + align_up(size_t(B + 1) * 4, 256);
  size_t tmp_bytes = sort_temp_bytes(int64_t(items), B);
  seg_offsets_kernel<<<(B + 256) / 256, 256, 0, s>>>(off, B, n);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  for (int64_t q0 = 0; q0 < nq; q0 += B) {
    const int b = int(std::min<int64_t>(B, nq - q0));
    dim3 grid(unsigned(std::min<int64_t>((n + 255) / 256, 4096)), unsigned(b));
    eval_dist_kernel<<<grid, 256, size_t(D) * 4, s>>>(E, int(N), int(D), queries + q0, k_in, v_in);
    MMSIM_CUDA_CHECK(::mmsim::launched());
    MMSIM_CUDA_CHECK(cub::DeviceSegmentedRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, int64_t(b) * n, b, off,
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
