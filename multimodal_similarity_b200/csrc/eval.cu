// K5: leave-one-out retrieval evaluation, one CTA per query -- the GPU form of the loop body of
//   utils.evaluate / utils.evaluate_simple            src/utils.py:83-229
//     retrieve_one(emb[i], np.delete(emb, i, 0), labels[i], np.delete(labels, i))      :55-81
//     precision_at_recall(labels[sorted_idx], labels[i], alpha)                        :231-255
//     recall_at_K(labels[sorted_idx], labels[i], K)                                    :257-266
// Per query: exact fp32 distances to the other N-1 rows (NumPy arithmetic, exact.cuh), a full bitonic sort by
// (distance, index) in shared memory, then
//   ap            sklearn average_precision_score on score = fl32(max(dist) - dist): thresholds at distinct scores
//   first_match   rank of the first ranked item whose label equals the query's  -> recall@K for every K
//   depth, hist   the prefix length precision_at_recall walks and the per-class counts inside it
// The host (retrieval.py) only averages these per-query records in the reference's order.
//
// REFERENCE QUIRK kept for parity (aligned == 0): the reference indexes the FULL label array with positions of
// the gallery-with-row-i-deleted (src/utils.py:128,132,185,190), so precision@recall / recall@K see the label of
// the predecessor for every item at original position >= i.  AP uses the correctly deleted labels (:78).
#include <cuda_runtime.h>

#include "common.cuh"
#include "eval.h"
#include "exact.cuh"

namespace mmsim {
namespace eval {

constexpr int THREADS = 256;

__device__ __forceinline__ void block_inclusive_scan(int* a, int n, int* partial) {
  const int t = threadIdx.x;
  const int chunk = (n + THREADS - 1) / THREADS;
  const int lo = min(n, t * chunk), hi = min(n, lo + chunk);
  int s = 0;
  for (int x = lo; x < hi; ++x) s += a[x];
  partial[t] = s;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int x = 0; x < THREADS; ++x) {
      const int v = partial[x];
      partial[x] = run;
      run += v;
    }
  }
  __syncthreads();
  int run = partial[t];
  for (int x = lo; x < hi; ++x) {
    run += a[x];
    a[x] = run;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS)
eval_rank_kernel(const float* __restrict__ E, const int* __restrict__ labels, const int* __restrict__ cls, int N, int D,
                 int C, const int* __restrict__ queries, int P, double alpha, int aligned,
                 double* __restrict__ out_ap, int* __restrict__ out_npos, int* __restrict__ out_first,
                 int* __restrict__ out_depth, int* __restrict__ out_hist, int* __restrict__ out_rank /* nullable [nq][N-1] */) {
  extern __shared__ __align__(16) unsigned char esm[];
  float* sk = reinterpret_cast<float*>(esm);          // [P] distances
  int* sv = reinterpret_cast<int*>(sk + P);           // [P] index in the row-i-deleted numbering
  int* sc = sv + P;                                   // [P] scan buffer
  float* qs = reinterpret_cast<float*>(sc + P);       // [D] query
  int* hist = reinterpret_cast<int*>(qs + D);         // [C]
  __shared__ int partial[THREADS];
  __shared__ double dred[THREADS];
  __shared__ int s_first, s_depth;

  const int t = threadIdx.x;
  const int qn = blockIdx.x;
  const int i = queries[qn];
  const int n = N - 1;
  const int ql = labels[i];

  for (int c = t; c < D; c += THREADS) qs[c] = E[size_t(i) * D + c];
  for (int c = t; c < C; c += THREADS) hist[c] = 0;
  if (t == 0) { s_first = n; s_depth = n; }
  __syncthreads();

  // ---- distances to every other row
  for (int jp = t; jp < P; jp += THREADS) {
    float d = kInf;
    int v = 0x7fffffff;
    if (jp < n) {
      const int j = jp + (jp >= i ? 1 : 0);
      d = exact_l2(qs, E + size_t(j) * D, D);
      v = jp;
    }
    sk[jp] = d;
    sv[jp] = v;
  }
  __syncthreads();

  // ---- full ranking: bitonic sort by (distance, index)
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int x = t; x < P / 2; x += THREADS) {
        const int lo = 2 * x - (x & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const float a = sk[lo], b = sk[hi];
        const int ia = sv[lo], ib = sv[hi];
        const bool gt = (a > b) || (a == b && ia > ib);
        if (gt == asc) {
          sk[lo] = b; sk[hi] = a;
          sv[lo] = ib; sv[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (out_rank)
    for (int r = t; r < n; r += THREADS) out_rank[size_t(qn) * n + r] = sv[r];

  // ---- average precision (correct labels): positives' precision at the end of their tie group
  for (int r = t; r < n; r += THREADS) {
    const int jp = sv[r];
    sc[r] = labels[jp + (jp >= i ? 1 : 0)] == ql ? 1 : 0;
  }
  __syncthreads();
  block_inclusive_scan(sc, n, partial);
  const int npos = n > 0 ? sc[n - 1] : 0;
  double acc = 0.0;
  if (npos > 0) {
    const float dmax = sk[n - 1];
    for (int r = t; r < n; r += THREADS) {
      const int prev = r ? sc[r - 1] : 0;
      if (sc[r] == prev) continue;                      // not a positive
      const float s = __fsub_rn(dmax, sk[r]);           // score = max(dist) - dist in float32 (src/utils.py:79)
      int e = r;
      while (e + 1 < n && __fsub_rn(dmax, sk[e + 1]) == s) ++e;   // tied scores share one threshold
      acc += double(sc[e]) / double(e + 1);
    }
  }
  dred[t] = acc;
  __syncthreads();
  for (int o = THREADS / 2; o > 0; o >>= 1) {
    if (t < o) dred[t] += dred[t + o];
    __syncthreads();
  }
  if (t == 0) {
    out_ap[qn] = npos > 0 ? dred[0] / double(npos) : 0.0;
    out_npos[qn] = npos;
  }
  __syncthreads();

  // ---- ranked labels as the reference forms them (quirk unless aligned)
  for (int r = t; r < n; r += THREADS) {
    const int jp = sv[r];
    const int src = aligned ? jp + (jp >= i ? 1 : 0) : jp;
    const int m = labels[src] == ql ? 1 : 0;
    sc[r] = m;
    if (m) atomicMin(&s_first, r);
  }
  __syncthreads();
  const int first_is_pos = n > 0 ? sc[0] : 0;
  block_inclusive_scan(sc, n, partial);
  const int cnt = n > 0 ? sc[n - 1] : 0;
  const int target = int(alpha * double(cnt));
  if (target == 0) {
    if (t == 0) s_depth = first_is_pos ? n : min(n, 1);
  } else {
    for (int r = t; r < n; r += THREADS) {
      const int prev = r ? sc[r - 1] : 0;
      if (sc[r] == target && prev == target - 1) s_depth = r + 1;   // exactly one r satisfies this
    }
  }
  __syncthreads();
  const int depth = s_depth;
  for (int r = t; r < depth; r += THREADS) {
    const int jp = sv[r];
    const int src = aligned ? jp + (jp >= i ? 1 : 0) : jp;
    atomicAdd(&hist[cls[src]], 1);
  }
  __syncthreads();
  for (int c = t; c < C; c += THREADS) out_hist[size_t(qn) * C + c] = hist[c];
  if (t == 0) {
    out_first[qn] = s_first;
    out_depth[qn] = depth;
  }
}

int run(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
        double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, cudaStream_t s) {
  MMSIM_REQUIRE(E && labels && cls && queries && ap && npos && first && depth && hist, MMSIM_ERR_ARG, "evaluate: null pointer argument");
  MMSIM_REQUIRE(N >= 2 && D >= 1 && C >= 1 && nq >= 0, MMSIM_ERR_ARG, "evaluate: bad sizes N=%lld D=%lld C=%d", (long long)N, (long long)D, C);
  int P = 32;
  while (P < N - 1) P <<= 1;
  const size_t smem = size_t(P) * 12 + size_t(D) * 4 + size_t(C) * 4;
  MMSIM_REQUIRE(smem <= 200 * 1024, MMSIM_ERR_UNSUPPORTED,
                "evaluate: N=%lld, D=%lld, C=%d needs %zu bytes of shared memory per query (limit 200 KiB: N <= 16385); "
                "use retrieve() + AP@k for larger galleries", (long long)N, (long long)D, C, smem);
  if (nq == 0) return MMSIM_OK;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(eval_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  eval_rank_kernel<<<unsigned(nq), THREADS, smem, s>>>(E, labels, cls, int(N), int(D), C, queries, P, alpha, aligned, ap, npos,
                                                       first, depth, hist, rank);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace eval
}  // namespace mmsim
