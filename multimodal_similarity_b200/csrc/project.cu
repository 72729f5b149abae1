// K7: embedding head -- projection + bias + L2 normalisation in one kernel, the step right before
// the hot path in the CUB pipeline:
//   networks.CUBLayer.forward        src/networks.py:376-380   logits = tf.nn.xw_plus_b(x, W, b)
//   base_model_CUB.py:197-201        embedding = tf.nn.l2_normalize(logits, axis=-1, epsilon=1e-10)   (if cfg.normalized)
// i.e. out[r] = y * rsqrt(max(sum(y^2), eps)), y = x[r] @ W + b.  fp32 FFMA tiles (the head is 0.8 GFLOP at
// 5,924 x 1024 -> 128, microseconds either way; fp32 keeps it within 1e-6 of the fp32 TF graph, a 16-bit tensor-core GEMM
// would not); a CTA owns 32 complete output rows, so the row norm is a warp reduction in the epilogue and the logits
// never reach memory.
#include <cuda_runtime.h>

#include "common.cuh"
#include "project.h"

namespace mmsim {
namespace project {

constexpr int PM = 32, PK = 32, PT = 256;   // rows per CTA, K step, threads (8 warps x 4 rows each)
constexpr int MAXC = 8;                     // output columns per thread: E <= 256

template <int NC>
__global__ void __launch_bounds__(PT)
project_normalize_kernel(const float* __restrict__ X, int64_t N, int K, const float* __restrict__ W, const float* __restrict__ b,
                         int E, int normalized, float eps, float* __restrict__ out) {
  extern __shared__ __align__(16) float psm[];
  float* Xs = psm;                   // [PM][PK + 1]
  float* Ws = Xs + PM * (PK + 1);    // [PK][NC * 32]
  const int t = threadIdx.x, ty = t >> 5, tx = t & 31;
  const int64_t r0 = int64_t(blockIdx.x) * PM;
  const int EC = NC * 32;            // padded output width held in shared memory
  float acc[4][NC];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[a][c] = 0.f;
  for (int k0 = 0; k0 < K; k0 += PK) {
    for (int x = t; x < PM * PK; x += PT) {
      const int r = x / PK, k = x - r * PK;
      Xs[r * (PK + 1) + k] = (r0 + r < N && k0 + k < K) ? X[(r0 + r) * K + k0 + k] : 0.f;
    }
    for (int x = t; x < PK * EC; x += PT) {
      const int k = x / EC, e = x - k * EC;
      Ws[x] = (k0 + k < K && e < E) ? W[size_t(k0 + k) * E + e] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < PK; ++k) {
      float xv[4], wv[NC];
#pragma unroll
      for (int a = 0; a < 4; ++a) xv[a] = Xs[(ty * 4 + a) * (PK + 1) + k];
#pragma unroll
      for (int c = 0; c < NC; ++c) wv[c] = Ws[k * EC + tx + 32 * c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[a][c] = fmaf(xv[a], wv[c], acc[a][c]);
    }
    __syncthreads();
  }
  // epilogue: bias, row norm (the 32 lanes of a warp hold one row's E columns), scale, store
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t r = r0 + ty * 4 + a;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int e = tx + 32 * c;
      if (e < E) {
        if (b) acc[a][c] += b[e];
        ss = fmaf(acc[a][c], acc[a][c], ss);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = normalized ? rsqrtf(fmaxf(ss, eps)) : 1.f;
    if (r < N) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int e = tx + 32 * c;
        if (e < E) out[r * E + e] = acc[a][c] * scale;
      }
    }
  }
}

template <int NC>
static int launch(const float* X, int64_t N, int K, const float* W, const float* b, int E, int normalized, float eps, float* out,
                  cudaStream_t s) {
  const size_t smem = (size_t(PM) * (PK + 1) + size_t(PK) * NC * 32) * 4;
  project_normalize_kernel<NC><<<unsigned((N + PM - 1) / PM), PT, smem, s>>>(X, N, K, W, b, E, normalized, eps, out);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

int run(const float* X, int64_t N, int64_t K, const float* W, const float* b, int64_t E, int normalized, float eps, float* out,
        cudaStream_t s) {
  MMSIM_REQUIRE(X && W && out, MMSIM_ERR_ARG, "project_normalize: null pointer argument");
  MMSIM_REQUIRE(N >= 0 && K >= 1 && K < (int64_t(1) << 30) && E >= 1, MMSIM_ERR_ARG, "project_normalize: bad sizes N=%lld K=%lld E=%lld",
                (long long)N, (long long)K, (long long)E);
  MMSIM_REQUIRE(E <= 32 * MAXC, MMSIM_ERR_UNSUPPORTED, "project_normalize: output width E=%lld > %d is not supported", (long long)E, 32 * MAXC);
  if (N == 0) return MMSIM_OK;
  switch ((E + 31) / 32) {
    case 1: return launch<1>(X, N, int(K), W, b, int(E), normalized, eps, out, s);
    case 2: return launch<2>(X, N, int(K), W, b, int(E), normalized, eps, out, s);
    case 3: case 4: return launch<4>(X, N, int(K), W, b, int(E), normalized, eps, out, s);
    default: return launch<8>(X, N, int(K), W, b, int(E), normalized, eps, out, s);
  }
}

}  // namespace project
}  // namespace mmsim
