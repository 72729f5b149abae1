// K6: merge per-shard top-k lists into the global top-k (no reference counterpart: the reference is single
// process; this is the exchange step of the gallery-sharded retrieval).
// Input parts are shard-local results, (dist f32, idx i32 local) sorted by (dist, idx); output is ordered by
// (dist, global idx) so every rank computes bit-identical results and the merged result equals the single-shard one.
//
// One warp per query.  The parts are sorted, so the k-th smallest distance overall is bounded by
// T = max_p part_p[ceil(k / parts) - 1]: only entries <= T can make the top-k.  They are compacted into shared memory
// (32-bit payload = slot in the gathered buffer; the global index is looked up only for exact-distance ties and for the
// k winners) and bitonic-sorted at the reduced size.
#include <cuda_runtime.h>

#include "common.cuh"
#include "merge.h"

namespace mmsim {
namespace merge {

constexpr int WARPS = 4;

__global__ void __launch_bounds__(WARPS * 32)
knn_merge_kernel(const float* __restrict__ dist_parts, const int* __restrict__ idx_parts,
                 int64_t part_stride, const int64_t* __restrict__ idx_base, int parts, int64_t nq, int k_in, int k, int P,
                 const float* __restrict__ lb_parts, int64_t lb_stride, float* __restrict__ out_dist,
                 void* __restrict__ out_idx_v, int idx32, int* __restrict__ status, float* __restrict__ out_flag,
                 const int* __restrict__ count_ptr, const int* __restrict__ row_map) {
  int64_t* out_idx = static_cast<int64_t*>(out_idx_v);
  int* out_idx32 = static_cast<int*>(out_idx_v);         // idx32: global indices fit 31 bits (half the bytes on the wire)
  extern __shared__ __align__(8) unsigned char msm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = int64_t(blockIdx.x) * WARPS + warp;
  if (qi >= nq || (count_ptr && qi >= *count_ptr)) return;
  const int64_t qo = row_map ? row_map[qi] : qi;      // output row (patch mode: the query a compact slot belongs to)
  float* sk = reinterpret_cast<float*>(msm) + size_t(warp) * 2 * P;   // [P] distance
  int* sv = reinterpret_cast<int*>(sk + P);                           // [P] slot = part * k_in + r
  const int total = parts * k_in;

  // ---- upper bound on the k-th smallest distance: the first `need` entries of every part are >= k entries in all
  const int need = (k + parts - 1) / parts;
  float T = kInf;
  if (need <= k_in) {
    float t = -kInf;
    for (int p = lane; p < parts; p += 32) {
      const size_t src = size_t(p) * part_stride + size_t(qi) * k_in + (need - 1);
      t = fmaxf(t, idx_parts[src] >= 0 ? dist_parts[src] : kInf);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    T = t;
  }

  // ---- compact the entries <= T
  int filled = 0;
  for (int x0 = 0; x0 < total; x0 += 32) {
    const int x = x0 + lane;
    float d = kInf;
    bool take = false;
    if (x < total) {
      const int part = x / k_in, r = x - part * k_in;
      const size_t src = size_t(part) * part_stride + size_t(qi) * k_in + r;
      if (idx_parts[src] >= 0) {
        d = dist_parts[src];
        take = d <= T;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, take);
    const int pos = filled + __popc(bal & ((1u << lane) - 1u));
    if (take && pos < P) {
      sk[pos] = d;
      sv[pos] = x;
    }
    filled += __popc(bal);
  }
  filled = min(filled, P);
  int S = 32;
  while (S < filled) S <<= 1;              // sort size for this query (warp-uniform)
  for (int x = filled + lane; x < S; x += 32) {
    sk[x] = kInf;
    sv[x] = 0x7fffffff;
  }
  __syncwarp();

  auto gidx = [&](int slot) -> int64_t {
    if (slot == 0x7fffffff) return INT64_MAX;
    const int part = slot / k_in, r = slot - part * k_in;
    return idx_base[part] + idx_parts[size_t(part) * part_stride + size_t(qi) * k_in + r];
  };
  for (int size = 2; size <= S; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < S / 2; t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const float a = sk[lo], b = sk[hi];
        const int ia = sv[lo], ib = sv[hi];
        // exact-distance ties (rare) are ordered by global index
        const bool gt = (a > b) || (a == b && ia != ib && gidx(ia) > gidx(ib));
        if (gt == asc) {
          sk[lo] = b; sk[hi] = a;
          sv[lo] = ib; sv[hi] = ia;
        }
      }
      __syncwarp();
    }
  }
  for (int r = lane; r < k; r += 32) {
    const bool ok = r < S && sv[r] != 0x7fffffff;
    out_dist[qo * k + r] = ok ? sk[r] : kInf;
    const int64_t gi = ok ? gidx(sv[r]) : -1;
    if (idx32) out_idx32[qo * k + r] = int(gi); else out_idx[qo * k + r] = gi;
  }
  // global certificate of the reduced-candidate protocol: every row a shard did NOT re-rank lies at distance >= that
  // shard's lower bound, so the merged top-k is exact iff its k-th distance is below every shard's bound
  if (lb_parts && lane == 0) {
    float lb = kInf;
    for (int p = 0; p < parts; ++p) lb = fminf(lb, lb_parts[size_t(p) * lb_stride + qi]);
    const float dk = (k - 1 < S) ? sk[k - 1] : kInf;
    const bool unc = !(dk < lb);
    if (unc) atomicAdd(&status[0], 1);
    // per-query flag for the per-query exact fallback: the merged k-th distance (an upper bound of the true one, +inf
    // when fewer than k candidates were merged) of an uncertified query, -1 for a certified one
    if (out_flag) out_flag[qi] = unc ? (dk == dk ? dk : kInf) : -1.f;
  }
}

int run(const float* dist_parts, const int* idx_parts, int64_t part_stride, const int64_t* idx_base, int parts, int64_t nq,
        int k_in, int k, const float* lb_parts, int64_t lb_stride, float* out_dist, void* out_idx, int* status,
        cudaStream_t s, float* out_flag, const int* count_ptr, const int* row_map, int idx32) {
  MMSIM_REQUIRE(dist_parts && idx_parts && idx_base && out_dist && out_idx, MMSIM_ERR_ARG, "knn_merge: null pointer argument");
  MMSIM_REQUIRE(parts >= 1 && k >= 1 && k_in >= 1 && nq >= 0 && part_stride >= nq * k_in && (!lb_parts || status), MMSIM_ERR_ARG,
                "knn_merge: bad sizes parts=%d nq=%lld k=%d", parts, (long long)nq, k);
  int P = 32;
  while (P < parts * k_in || P < k) P <<= 1;
  MMSIM_REQUIRE(P <= 4096, MMSIM_ERR_UNSUPPORTED, "knn_merge: parts*k = %d exceeds 4096", parts * k_in);
  if (nq == 0) return MMSIM_OK;
  const size_t smem = size_t(WARPS) * P * 8;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  knn_merge_kernel<<<unsigned((nq + WARPS - 1) / WARPS), WARPS * 32, smem, s>>>(dist_parts, idx_parts, part_stride, idx_base, parts, nq, k_in, k, P,
                                                                              lb_parts, lb_stride, out_dist, out_idx, idx32, status, out_flag,
                                                                              count_ptr, row_map);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace merge
}  // namespace mmsim
