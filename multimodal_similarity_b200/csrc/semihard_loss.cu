// K8: tf.contrib's triplet_semihard_loss, forward + backward -- the loss the reference's CUB
// trainers select with --loss triplet:
//   metric_loss_ops.triplet_semihard_loss(labels, embeddings, margin)      src/base_CUB.py:163-166, src/debug_CUB.py:211-214
// The function lives in TensorFlow 1.x (tensorflow/contrib/losses/python/metric_learning/metric_loss_ops.py), which is
// not vendored by the reference and not installable here: PARITY UNPINNED.  Its published algorithm, restated:
//   D      = pairwise SQUARED distances (diagonal 0)
//   for every anchor i and positive k (same label, k != i):
//     n*   = min { D_ij : label_j != label_i, D_ij > D_ik }         (the "semi-hard" negative, outside the positive)
//            or, if that set is empty, max { D_ij : label_j != label_i }   ("negatives_inside": the largest negative;
//            with no negative at all masked_maximum degenerates to the row minimum, i.e. 0)
//     l_ik = max(margin + D_ik - n*, 0)
//   loss   = sum l_ik / #(i, k)
// Here: D from the exact difference-form kernel (sqdist.cu; TF uses the Gram form and clamps at 0 -- the difference form is
// that without the cancellation), then one CTA per anchor: the anchor's row of D and the labels in shared memory, one warp
// per positive scanning the row, the loss term and the three-row sparse gradient (dD_ij/de_i = 2 (e_i - e_j)) per
// active pair.  Deterministic loss: per-anchor partial sums, summed in order by a one-block kernel.
#include <cuda_runtime.h>

#include "common.cuh"
#include "semihard_loss.h"
#include "sqdist.h"

namespace mmsim {
namespace semihard_loss {

constexpr int T = 128;

// ws[0] = number of (anchor, positive) pairs, as float
__global__ void __launch_bounds__(1024) count_pairs_kernel(const int* __restrict__ labels, int N, float* __restrict__ num_pos) {
  __shared__ int red[32];
  int c = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int li = labels[i];
    for (int j = 0; j < N; ++j) c += (labels[j] == li && j != i) ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) tot += red[w];
    *num_pos = float(tot);
  }
}

__global__ void __launch_bounds__(T)
anchor_kernel(const float* __restrict__ E, const int* __restrict__ labels, const float* __restrict__ Dm, int N, int D, float margin,
              const float* __restrict__ num_pos, float* __restrict__ partial, float* __restrict__ dE) {
  extern __shared__ float sm[];
  float* row = sm;                                  // [N] D_i.
  int* lab = reinterpret_cast<int*>(sm + N);        // [N]
  __shared__ float wsum[T / 32];
  const int i = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int j = t; j < N; j += T) { row[j] = Dm[size_t(i) * N + j]; lab[j] = labels[j]; }
  __syncthreads();
  const int li = lab[i];
  const float scale = 2.0f / *num_pos;              // d(loss)/d(l_ik) = 1 / #pairs, dD/de carries the 2
  float acc = 0.f;
  for (int k = warp; k < N; k += T / 32) {          // warp-uniform: one positive per warp and step
    if (k == i || lab[k] != li) continue;
    const float dik = row[k];
    float best_out = kInf, best_in = -kInf;         // smallest negative beyond the positive / largest negative
    int j_out = -1, j_in = -1;
    for (int j = lane; j < N; j += 32) {
      if (lab[j] == li) continue;
      const float d = row[j];
      if (d > dik && (d < best_out)) { best_out = d; j_out = j; }       // ascending j: ties keep the smallest index
      if (d > best_in) { best_in = d; j_in = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float bo = __shfl_xor_sync(0xffffffffu, best_out, o), bi = __shfl_xor_sync(0xffffffffu, best_in, o);
      const int jo = __shfl_xor_sync(0xffffffffu, j_out, o), ji = __shfl_xor_sync(0xffffffffu, j_in, o);
      if (bo < best_out || (bo == best_out && jo >= 0 && (j_out < 0 || jo < j_out))) { best_out = bo; j_out = jo; }
      if (bi > best_in || (bi == best_in && ji >= 0 && (j_in < 0 || ji < j_in))) { best_in = bi; j_in = ji; }
    }
    const int jn = j_out >= 0 ? j_out : j_in;       // -1: the anchor has no negative at all
    const float nstar = jn >= 0 ? (j_out >= 0 ? best_out : best_in) : 0.f;   // masked_maximum with an empty mask = row minimum = D_ii = 0
    const float l = margin + dik - nstar;
    if (l > 0.f) {
      if (lane == 0) acc += l;
      if (dE) {
        const float* ei = E + size_t(i) * D;
        const float* ek = E + size_t(k) * D;
        const float* en = E + size_t(jn >= 0 ? jn : i) * D;
        for (int d = lane; d < D; d += 32) {
          const float vi = ei[d], vk = ek[d];
          float gi = scale * (vi - vk);                        // + dD_ik
          atomicAdd(&dE[size_t(k) * D + d], scale * (vk - vi));
          if (jn >= 0) {
            const float vn = en[d];
            gi -= scale * (vi - vn);                           // - dD_in
            atomicAdd(&dE[size_t(jn) * D + d], scale * (vi - vn));
          }
          atomicAdd(&dE[size_t(i) * D + d], gi);
        }
      }
    }
  }
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (t == 0) {
    float s = 0.f;
    for (int w = 0; w < T / 32; ++w) s += wsum[w];
    partial[i] = s;
  }
}

__global__ void __launch_bounds__(256) finish_kernel(const float* __restrict__ partial, int N, const float* __restrict__ num_pos,
                                                     float* __restrict__ loss) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] / *num_pos;   // 0/0 = NaN without any positive pair, like the TF graph
}

int workspace_bytes(int64_t N, size_t* out) {
  MMSIM_REQUIRE(out && N >= 1 && N <= 8192, MMSIM_ERR_ARG, "triplet_semihard: N must be in [1, 8192]");
  *out = align_up(size_t(N) * N * 4, 256) + align_up(size_t(N) * 4, 256) + 256;
  return MMSIM_OK;
}

int run(const float* E, const int* labels, int64_t N, int64_t D, float margin, float* loss, float* dE, void* ws, size_t ws_bytes,
        cudaStream_t s) {
  MMSIM_REQUIRE(E && labels && loss && ws, MMSIM_ERR_ARG, "triplet_semihard: null pointer argument");
  MMSIM_REQUIRE(N >= 1 && N <= 8192 && D >= 1, MMSIM_ERR_UNSUPPORTED, "triplet_semihard: N=%lld D=%lld unsupported (N <= 8192)",
                (long long)N, (long long)D);
  size_t need = 0;
  workspace_bytes(N, &need);
  MMSIM_REQUIRE(ws_bytes >= need, MMSIM_ERR_WORKSPACE, "triplet_semihard: workspace too small (%zu < %zu)", ws_bytes, need);
  MMSIM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MMSIM_ERR_WORKSPACE, "triplet_semihard: workspace must be 256-byte aligned");
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* Dm = reinterpret_cast<float*>(w);
  float* partial = reinterpret_cast<float*>(w + align_up(size_t(N) * N * 4, 256));
  float* num_pos = reinterpret_cast<float*>(w + align_up(size_t(N) * N * 4, 256) + align_up(size_t(N) * 4, 256));
  if (int rc = sqdist::run(E, N, E, N, D, 0, Dm, N, s)) return rc;
  count_pairs_kernel<<<1, 1024, 0, s>>>(labels, int(N), num_pos);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  if (dE) MMSIM_CUDA_CHECK(cudaMemsetAsync(dE, 0, size_t(N) * D * 4, s));
  const size_t smem = size_t(N) * 8;
  if (smem > 48 * 1024) MMSIM_CUDA_CHECK(cudaFuncSetAttribute(anchor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  anchor_kernel<<<unsigned(N), T, smem, s>>>(E, labels, Dm, int(N), int(D), margin, num_pos, partial, dE);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  finish_kernel<<<1, 256, 0, s>>>(partial, int(N), num_pos, loss);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace semihard_loss
}  // namespace mmsim
