// K7: embedding head -- projection + bias + L2 normalisation in one kernel, the step right before
// the hot path in the CUB pipeline:
//   networks.CUBLayer.forward        src/networks.py:376-380   logits = tf.nn.xw_plus_b(x, W, b)
//   base_model_CUB.py:197-201        embedding = tf.nn.l2_normalize(logits, axis=-1, epsilon=1e-10)   (if cfg.normalized)
// i.e. out[r] = y * rsqrt(max(sum(y^2), eps)), y = x[r] @ W + b.  fp32 FFMA tiles (the head is 1.55 GFLOP at
// 5,924 x 1024 -> 128, microseconds either way; fp32 keeps it within 1e-6 of the fp32 TF graph, a 16-bit tensor-core GEMM
// would not); a CTA owns 16 complete output rows, so the row norm is a warp reduction in the epilogue and the logits
// never reach memory.
#include <cuda_runtime.h>

#include "common.cuh"
#include "project.h"

namespace mmsim {
namespace project {

// Round 2: register tile of 4 rows x 4 NC4 columns per thread fed by 128-bit shared-memory loads (8 FMAs per LDS.128; the
// first version issued one 32-bit load per two FMAs and ran at 13 TFLOP/s = 117 us at 5,924 x 1,024 -> 128), 16 rows per
// CTA so that 371 CTAs spread over the 148 SMs, 128-bit global loads when the rows allow it.
constexpr int PM = 16, PK = 32, PT = 128;   // rows per CTA, K step, threads (4 warps x 4 rows each)
constexpr int MAXC = 8;                     // output columns per thread / 32: E <= 256

template <int NC4>                          // float4 column groups per lane: E <= 128 NC4
__global__ void __launch_bounds__(PT)
project_normalize_kernel(const float* __restrict__ X, int64_t N, int K, const float* __restrict__ W, const float* __restrict__ b,
                         int E, int normalized, float eps, float* __restrict__ out) {
  extern __shared__ __align__(16) float psm[];
  constexpr int EC = NC4 * 128;      // padded output width held in shared memory
  float* Xs = psm;                   // [PM][PK]      (a warp reads one row at a time: broadcast)
  float* Ws = Xs + PM * PK;          // [PK][EC]      (lane tx reads the float4 groups tx, tx + 32, ...)
  const int t = threadIdx.x, ty = t >> 5, tx = t & 31;
  const int64_t r0 = int64_t(blockIdx.x) * PM;
  const bool vx = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  const bool vw = (E & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  float acc[4][NC4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < NC4; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][c][j] = 0.f;
  // global -> registers -> shared memory, one K step ahead: the loads of step k0 + PK are in flight while step k0 is multiplied
  // (without this prefetch every one of the K / 32 steps waited a full memory round trip: 107 us at 5,924 x 1,024 -> 128)
  constexpr int WV = PK * (EC / 4) / PT;          // float4 of the W tile per thread
  float4 xr, wr[WV];
  auto fetch = [&](int k0) {
    {
      const int r = t >> 3, k = (t & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < N) {
        const float* src = X + (r0 + r) * K + k0 + k;
        if (vx && k0 + k + 3 < K) v = *reinterpret_cast<const float4*>(src);
        else {
          if (k0 + k < K) v.x = src[0];
          if (k0 + k + 1 < K) v.y = src[1];
          if (k0 + k + 2 < K) v.z = src[2];
          if (k0 + k + 3 < K) v.w = src[3];
        }
      }
      xr = v;
    }
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int x = t + i * PT;
      const int k = x / (EC / 4), e = (x - k * (EC / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < K) {
        const float* src = W + size_t(k0 + k) * E + e;
        if (vw && e + 3 < E) v = *reinterpret_cast<const float4*>(src);
        else {
          if (e < E) v.x = src[0];
          if (e + 1 < E) v.y = src[1];
          if (e + 2 < E) v.z = src[2];
          if (e + 3 < E) v.w = src[3];
        }
      }
      wr[i] = v;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += PK) {
    *reinterpret_cast<float4*>(Xs + (t >> 3) * PK + (t & 7) * 4) = xr;
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int x = t + i * PT;
      const int k = x / (EC / 4), e = (x - k * (EC / 4)) * 4;
      *reinterpret_cast<float4*>(Ws + k * EC + e) = wr[i];
    }
    __syncthreads();
    if (k0 + PK < K) fetch(k0 + PK);
#pragma unroll 2
    for (int k = 0; k < PK; k += 4) {
      float4 xv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xv[a] = *reinterpret_cast<const float4*>(Xs + (ty * 4 + a) * PK + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int c = 0; c < NC4; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(Ws + (k + kk) * EC + (c * 32 + tx) * 4);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const float x = kk == 0 ? xv[a].x : kk == 1 ? xv[a].y : kk == 2 ? xv[a].z : xv[a].w;
            acc[a][c][0] = fmaf(x, w.x, acc[a][c][0]);
            acc[a][c][1] = fmaf(x, w.y, acc[a][c][1]);
            acc[a][c][2] = fmaf(x, w.z, acc[a][c][2]);
            acc[a][c][3] = fmaf(x, w.w, acc[a][c][3]);
          }
        }
      }
    }
    __syncthreads();
  }
  // epilogue: bias, row norm (the 32 lanes of a warp hold one row's E columns), scale, store
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t r = r0 + ty * 4 + a;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NC4; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = (c * 32 + tx) * 4 + j;
        if (e < E) {
          if (b) acc[a][c][j] += b[e];
          ss = fmaf(acc[a][c][j], acc[a][c][j], ss);
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = normalized ? rsqrtf(fmaxf(ss, eps)) : 1.f;
    if (r < N) {
#pragma unroll
      for (int c = 0; c < NC4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = (c * 32 + tx) * 4 + j;
          if (e < E) out[r * E + e] = acc[a][c][j] * scale;
        }
    }
  }
}

template <int NC4>
static int launch(const float* X, int64_t N, int K, const float* W, const float* b, int E, int normalized, float eps, float* out,
                  cudaStream_t s) {
  const size_t smem = (size_t(PM) * PK + size_t(PK) * NC4 * 128) * 4;
  project_normalize_kernel<NC4><<<unsigned((N + PM - 1) / PM), PT, smem, s>>>(X, N, K, W, b, E, normalized, eps, out);
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
  return E <= 128 ? launch<1>(X, N, int(K), W, b, int(E), normalized, eps, out, s)
                  : launch<2>(X, N, int(K), W, b, int(E), normalized, eps, out, s);
}

}  // namespace project
}  // namespace mmsim
