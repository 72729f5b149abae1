// Semi-hard (FaceNet) negative mining on the device -- the inner test of the reference's host miner
//   utils.select_triplets_facenet            src/utils.py:430-496 (and its six near-copies)
// For every requested (anchor, positive) pair, the set
//   { n : label[n] != label[anchor],  pos_dist < dist[anchor, n],  fl32(dist[anchor, n] - pos_dist) < fl32(alpha) }
// (src/utils.py:476-479: rows of the anchor's class are NaN-masked, both comparisons in float32): its size (the
// reference's len(all_neg)), optionally the set itself as a bitmask, and -- second entry point -- its r-th member in
// ascending row order (the reference's all_neg[r]).  The RNG-coupled pair order and the random draws stay on the host
// (multimodal_similarity_b200/mining.py); they depend on the counts only, so neither the distance matrix nor the masks
// ever leave the device and the same seeds give the same triplets as the reference.
#include <cuda_runtime.h>

#include "common.cuh"
#include "mining.h"

namespace mmsim {
namespace mining {

__device__ __forceinline__ bool semihard(float nd, float pd, float alpha) {
  return (__fsub_rn(nd, pd) < alpha) && (pd < nd);   // NaN distances compare false, like the reference's masked rows
}

// one warp per pair; a warp reads 128 contiguous bytes of the anchor's row per step
__global__ void __launch_bounds__(128)
semihard_mask_kernel(const float* __restrict__ dist, int64_t n, int64_t ld, const int* __restrict__ labels,
                     const int* __restrict__ pairs, int64_t m, float alpha, uint32_t* __restrict__ mask,
                     int* __restrict__ count) {
  const int64_t pi = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pi >= m) return;
  const int an = pairs[2 * pi], pos = pairs[2 * pi + 1];
  const float* row = dist + int64_t(an) * ld;
  const float pd = row[pos];
  const int key = labels[an];
  const int64_t words = (n + 31) / 32;
  int total = 0;
  for (int64_t w = 0; w < words; ++w) {
    const int64_t j = w * 32 + lane;
    const bool ok = j < n && labels[j] != key && semihard(row[j], pd, alpha);
    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
    if (mask != nullptr && lane == 0) mask[pi * words + w] = bal;
    total += __popc(bal);
  }
  if (lane == 0) count[pi] = total;
}

// one warp per pick {anchor, positive, r}: row index of the r-th (0-based) semi-hard negative, -1 if there are <= r
__global__ void __launch_bounds__(128)
semihard_pick_kernel(const float* __restrict__ dist, int64_t n, int64_t ld, const int* __restrict__ labels,
                     const int* __restrict__ picks, int64_t p, float alpha, int* __restrict__ neg_idx) {
  const int64_t qi = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= p) return;
  const int an = picks[3 * qi], pos = picks[3 * qi + 1];
  int r = picks[3 * qi + 2];
  const float* row = dist + int64_t(an) * ld;
  const float pd = row[pos];
  const int key = labels[an];
  const int64_t words = (n + 31) / 32;
  int found = -1;
  for (int64_t w = 0; w < words; ++w) {
    const int64_t j = w * 32 + lane;
    const bool ok = j < n && labels[j] != key && semihard(row[j], pd, alpha);
    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
    const int c = __popc(bal);
    if (r < c) {
      found = int(w * 32) + __fns(bal, 0, r + 1);   // position of the (r+1)-th set bit
      break;
    }
    r -= c;
  }
  if (lane == 0) neg_idx[qi] = found;
}

static int check_common(const void* dist, int64_t n, int64_t ld, const void* labels, const void* list, const void* out,
                        int64_t m, const char* who) {
  MMSIM_REQUIRE(dist && labels && (m == 0 || (list && out)), MMSIM_ERR_ARG, "%s: null pointer argument", who);
  MMSIM_REQUIRE(n >= 1 && ld >= n && m >= 0 && m < (int64_t(1) << 26), MMSIM_ERR_ARG, "%s: bad sizes n=%lld ld=%lld m=%lld",
                who, (long long)n, (long long)ld, (long long)m);
  return MMSIM_OK;
}

int run_mask(const float* dist, int64_t n, int64_t ld, const int* labels, const int* pairs, int64_t m, float alpha,
             uint32_t* mask, int* count, cudaStream_t s) {
  if (int rc = check_common(dist, n, ld, labels, pairs, count, m, "semihard_mask")) return rc;
  if (m == 0) return MMSIM_OK;
  semihard_mask_kernel<<<unsigned((m * 32 + 127) / 128), 128, 0, s>>>(dist, n, ld, labels, pairs, m, alpha, mask, count);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

int run_pick(const float* dist, int64_t n, int64_t ld, const int* labels, const int* picks, int64_t p, float alpha,
             int* neg_idx, cudaStream_t s) {
  if (int rc = check_common(dist, n, ld, labels, picks, neg_idx, p, "semihard_pick")) return rc;
  if (p == 0) return MMSIM_OK;
  semihard_pick_kernel<<<unsigned((p * 32 + 127) / 128), 128, 0, s>>>(dist, n, ld, labels, picks, p, alpha, neg_idx);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace mining
}  // namespace mmsim
