// K6: merge per-shard top-k lists into the global top-k (no reference counterpart: the reference is single
// process; this is the exchange step of the gallery-sharded retrieval).
// Input parts are the shard-local results of mmsim_knn_f32, (dist f32, idx i32 local) sorted by (dist, idx);
// output is ordered by (dist, global idx) so every rank computes bit-identical results and the merged result
// equals the single-shard result.
#include <cuda_runtime.h>

#include "common.cuh"
#include "merge.h"

namespace mmsim {
namespace merge {

constexpr int WARPS = 4;

// One warp per query.  All parts * k pairs go to shared memory and are bitonic-sorted by (dist, global idx).
__global__ void __launch_bounds__(WARPS * 32)
knn_merge_kernel(const float* __restrict__ dist_parts, const int* __restrict__ idx_parts,
                 int64_t part_stride, const int64_t* __restrict__ idx_base, int parts, int64_t nq, int k_in, int k, int P,
                 const float* __restrict__ lb_parts, int64_t lb_stride, float* __restrict__ out_dist,
                 int64_t* __restrict__ out_idx, int* __restrict__ status) {
  extern __shared__ __align__(8) unsigned char msm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = int64_t(blockIdx.x) * WARPS + warp;
  if (qi >= nq) return;
  int64_t* sv = reinterpret_cast<int64_t*>(msm) + size_t(warp) * P;
  float* sk = reinterpret_cast<float*>(reinterpret_cast<int64_t*>(msm) + size_t(WARPS) * P) + size_t(warp) * P;
  const int total = parts * k_in;
  for (int x = lane; x < P; x += 32) {
    float d = kInf;
    int64_t g = INT64_MAX;
    if (x < total) {
      const int part = x / k_in, r = x - part * k_in;
      const size_t src = size_t(part) * part_stride + size_t(qi) * k_in + r;
      const int li = idx_parts[src];
      if (li >= 0) {
        d = dist_parts[src];
        g = idx_base[part] + li;
      }
    }
    sk[x] = d;
    sv[x] = g;
  }
  __syncwarp();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < P / 2; t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const float a = sk[lo], b = sk[hi];
        const int64_t ia = sv[lo], ib = sv[hi];
        const bool gt = (a > b) || (a == b && ia > ib);
        if (gt == asc) {
          sk[lo] = b; sk[hi] = a;
          sv[lo] = ib; sv[hi] = ia;
        }
      }
      __syncwarp();
    }
  }
  for (int r = lane; r < k; r += 32) {
    out_dist[qi * k + r] = sk[r];
    out_idx[qi * k + r] = sv[r] == INT64_MAX ? -1 : sv[r];
  }
  // global certificate of the reduced-candidate protocol: every row a shard did NOT re-rank lies at distance >= that
  // shard's lower bound, so the merged top-k is exact iff its k-th distance is below every shard's bound
  if (lb_parts && lane == 0) {
    float lb = kInf;
    for (int p = 0; p < parts; ++p) lb = fminf(lb, lb_parts[size_t(p) * lb_stride + qi]);
    if (!(sk[k - 1] < lb)) atomicAdd(&status[0], 1);
  }
}

int run(const float* dist_parts, const int* idx_parts, int64_t part_stride, const int64_t* idx_base, int parts, int64_t nq,
        int k_in, int k, const float* lb_parts, int64_t lb_stride, float* out_dist, int64_t* out_idx, int* status,
        cudaStream_t s) {
  MMSIM_REQUIRE(dist_parts && idx_parts && idx_base && out_dist && out_idx, MMSIM_ERR_ARG, "knn_merge: null pointer argument");
  MMSIM_REQUIRE(parts >= 1 && k >= 1 && k_in >= 1 && nq >= 0 && part_stride >= nq * k_in && (!lb_parts || status), MMSIM_ERR_ARG, "knn_merge: bad sizes parts=%d nq=%lld k=%d", parts, (long long)nq, k);
  int P = 32;
  while (P < parts * k_in || P < k) P <<= 1;
  MMSIM_REQUIRE(P <= 4096, MMSIM_ERR_UNSUPPORTED, "knn_merge: parts*k = %d exceeds 4096", parts * k_in);
  if (nq == 0) return MMSIM_OK;
  const size_t smem = size_t(WARPS) * P * 12;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  knn_merge_kernel<<<unsigned((nq + WARPS - 1) / WARPS), WARPS * 32, smem, s>>>(dist_parts, idx_parts, part_stride, idx_base, parts, nq, k_in, k, P,
                                                                              lb_parts, lb_stride, out_dist, out_idx, status);
  MMSIM_CUDA_CHECK(cudaGetLastError());
  return MMSIM_OK;
}

}  // namespace merge
}  // namespace mmsim
