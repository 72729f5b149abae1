// K9: tf.contrib's lifted_struct_loss, forward + backward -- what the reference's CUB trainers select
// with --loss lifted:
//   metric_loss_ops.lifted_struct_loss(labels, embeddings, margin)        src/base_CUB.py:167-171, src/debug_CUB.py:215-219
// Third-party (TensorFlow 1.x contrib, not vendored, not installable here): PARITY UNPINNED.  Its published algorithm
// (Song et al., "Deep Metric Learning via Lifted Structured Feature Embedding"), restated:
//   D      = pairwise Euclidean distances (NOT squared; diagonal 0)
//   S_a    = sum over negatives j of a of exp(margin - D_aj)
//   J_ab   = log(S_a + S_b) + D_ab                 for every ordered positive pair (a, b), a != b
//   loss   = 0.25 * sum_(a,b) max(J_ab, 0)^2 / (P / 2),   P = number of ordered positive pairs
// (TF evaluates log(S_a + S_b) as M_ab + log(sum exp(. - M_ab)) with M_ab = max(m_a, m_b), m_a the largest
// margin - D_aj over a's negatives -- the same value; here each S_a is kept as (max, scaled sum).)
// Gradient: c_ab = max(J_ab, 0) / P per ordered pair;  dL/dD_ab += c_ab (+ c_ba);  dL/dD_aj -= exp(margin - D_aj) * W_a for
// every negative j of a, W_a = sum_b 2 c_ab / (S_a + S_b);  dD_ij/de_i = (e_i - e_j) / D_ij.
// Kernels: exact squared distances (sqdist.cu) -> per-row S (one warp per row) -> per-row pair terms, loss partial and
// log W (one CTA per row) -> per-row gradient gather (one CTA per row: no atomics, deterministic).
#include <cuda_runtime.h>

#include "common.cuh"
#include "lifted_struct.h"
#include "sqdist.h"

namespace mmsim {
namespace lifted_struct {

constexpr int T = 128;

__device__ __forceinline__ float dist_of(float sq) { return sq > 0.f ? sqrtf(sq) : 0.f; }
// log(exp(a) + exp(b)) for possibly -inf arguments
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -kInf) return -kInf;
  return m + logf(expf(a - m) + expf(b - m));
}

// logS[a] = log sum_{j: label_j != label_a} exp(margin - D_aj)  (-inf without negatives); num_pos = ordered positive pairs
__global__ void __launch_bounds__(T)
row_sums_kernel(const float* __restrict__ Dm, const int* __restrict__ labels, int N, float margin, float* __restrict__ logS,
                float* __restrict__ num_pos_rows) {
  const int a = (blockIdx.x * T + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= N) return;
  const int la = labels[a];
  float m = -kInf, s = 0.f;
  int pos = 0;
  for (int j = lane; j < N; j += 32) {
    if (labels[j] == la) { pos += j != a ? 1 : 0; continue; }
    const float x = margin - dist_of(Dm[size_t(a) * N + j]);
    if (x <= m) s += expf(x - m); else { s = s * expf(m - x) + 1.f; m = x; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    if (mm > -kInf) { s = s * expf(m - mm) + s2 * expf(m2 - mm); m = mm; }
    pos += __shfl_xor_sync(0xffffffffu, pos, o);
  }
  if (lane == 0) {
    logS[a] = s > 0.f ? m + logf(s) : -kInf;
    num_pos_rows[a] = float(pos);
  }
}

__global__ void __launch_bounds__(1024) total_kernel(const float* __restrict__ rows, int N, float* __restrict__ total) {
  __shared__ float red[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += 1024) s += rows[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = red[0];
}

// per row a: loss partial over its positives b, and logW[a] = log sum_b 2 c_ab / (S_a + S_b)   (-inf when nothing is active)
__global__ void __launch_bounds__(T)
pair_kernel(const float* __restrict__ Dm, const int* __restrict__ labels, int N, const float* __restrict__ logS,
            const float* __restrict__ num_pos, float* __restrict__ partial, float* __restrict__ logW) {
  __shared__ float r_l[T], r_m[T], r_s[T];
  const int a = blockIdx.x, t = threadIdx.x;
  const int la = labels[a];
  const float P = *num_pos, lsa = logS[a];
  float l = 0.f, m = -kInf, s = 0.f;
  for (int b = t; b < N; b += T) {
    if (b == a || labels[b] != la) continue;
    const float lab_ = lse2(lsa, logS[b]);
    const float J = lab_ + dist_of(Dm[size_t(a) * N + b]);
    if (J > 0.f) {                                   // (J = -inf when neither row has a negative)
      l += J * J;
      const float x = logf(2.f * J / P) - lab_;      // log of 2 c_ab / (S_a + S_b)
      if (x <= m) s += expf(x - m); else { s = s * expf(m - x) + 1.f; m = x; }
    }
  }
  r_l[t] = l; r_m[t] = m; r_s[t] = s;
  __syncthreads();
  if (t == 0) {                                      // fixed order: deterministic
    float L = 0.f, M = -kInf, S = 0.f;
    for (int x = 0; x < T; ++x) {
      L += r_l[x];
      const float mm = fmaxf(M, r_m[x]);
      if (mm > -kInf) { S = S * expf(M - mm) + r_s[x] * expf(r_m[x] - mm); M = mm; }
    }
    partial[a] = L;
    logW[a] = S > 0.f ? M + logf(S) : -kInf;
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
  if (threadIdx.x == 0) *loss = 0.25f * red[0] / (*num_pos * 0.5f);
}

// dE[a] = sum_j T_aj (e_a - e_j) / D_aj,  T_aj = dL/dD_aj of the unordered pair:
//   positives: c_aj + c_ja = 2 max(J_aj, 0) / P;   negatives: -exp(margin - D_aj) (W_a + W_j)
__global__ void __launch_bounds__(T)
grad_kernel(const float* __restrict__ E, const float* __restrict__ Dm, const int* __restrict__ labels, int N, int D, float margin,
            const float* __restrict__ logS, const float* __restrict__ logW, const float* __restrict__ num_pos,
            float* __restrict__ dE) {
  extern __shared__ float coef[];                   // [N] T_aj / D_aj
  const int a = blockIdx.x, t = threadIdx.x;
  const int la = labels[a];
  const float P = *num_pos, lsa = logS[a], lwa = logW[a];
  for (int j = t; j < N; j += T) {
    float c = 0.f;
    if (j != a) {
      const float d = dist_of(Dm[size_t(a) * N + j]);
      if (d > 0.f) {
        if (labels[j] == la) {
          const float J = lse2(lsa, logS[j]) + d;
          if (J > 0.f) c = 2.f * J / P / d;
        } else {
          const float w = lse2(lwa, logW[j]);        // log(W_a + W_j)
          if (w > -kInf) c = -expf(margin - d + w) / d;
        }
      }
    }
    coef[j] = c;
  }
  __syncthreads();
  for (int x = t; x < D; x += T) {
    const float ea = E[size_t(a) * D + x];
    float g = 0.f;
    for (int j = 0; j < N; ++j) g = fmaf(coef[j], ea - E[size_t(j) * D + x], g);
    dE[size_t(a) * D + x] = g;
  }
}

int workspace_bytes(int64_t N, size_t* out) {
  MMSIM_REQUIRE(out && N >= 1 && N <= 8192, MMSIM_ERR_ARG, "lifted_struct: N must be in [1, 8192]");
  *out = align_up(size_t(N) * N * 4, 256) + 4 * align_up(size_t(N) * 4, 256) + 256;
  return MMSIM_OK;
}

int run(const float* E, const int* labels, int64_t N, int64_t D, float margin, float* loss, float* dE, void* ws, size_t ws_bytes,
        cudaStream_t s) {
  MMSIM_REQUIRE(E && labels && loss && ws, MMSIM_ERR_ARG, "lifted_struct: null pointer argument");
  MMSIM_REQUIRE(N >= 1 && N <= 8192 && D >= 1, MMSIM_ERR_UNSUPPORTED, "lifted_struct: N=%lld D=%lld unsupported (N <= 8192)",
                (long long)N, (long long)D);
  size_t need = 0;
  workspace_bytes(N, &need);
  MMSIM_REQUIRE(ws_bytes >= need, MMSIM_ERR_WORKSPACE, "lifted_struct: workspace too small (%zu < %zu)", ws_bytes, need);
  MMSIM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MMSIM_ERR_WORKSPACE, "lifted_struct: workspace must be 256-byte aligned");
  uint8_t* w = static_cast<uint8_t*>(ws);
  const size_t vec = align_up(size_t(N) * 4, 256);
  float* Dm = reinterpret_cast<float*>(w);
  uint8_t* v0 = w + align_up(size_t(N) * N * 4, 256);
  float* logS = reinterpret_cast<float*>(v0);
  float* pos_rows = reinterpret_cast<float*>(v0 + vec);
  float* partial = reinterpret_cast<float*>(v0 + 2 * vec);
  float* logW = reinterpret_cast<float*>(v0 + 3 * vec);
  float* num_pos = reinterpret_cast<float*>(v0 + 4 * vec);
  if (int rc = sqdist::run(E, N, E, N, D, 0, Dm, N, s)) return rc;
  row_sums_kernel<<<unsigned((N * 32 + T - 1) / T), T, 0, s>>>(Dm, labels, int(N), margin, logS, pos_rows);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  total_kernel<<<1, 1024, 0, s>>>(pos_rows, int(N), num_pos);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  pair_kernel<<<unsigned(N), T, 0, s>>>(Dm, labels, int(N), logS, num_pos, partial, logW);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  finish_kernel<<<1, 256, 0, s>>>(partial, int(N), num_pos, loss);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  if (dE) {
    grad_kernel<<<unsigned(N), T, size_t(N) * 4, s>>>(E, Dm, labels, int(N), int(D), margin, logS, logW, num_pos, dE);
    MMSIM_CUDA_CHECK(::mmsim::launched());
  }
  return MMSIM_OK;
}

}  // namespace lifted_struct
}  // namespace mmsim
