// K1 (exact form): pairwise distance matrix, bit-identical to the reference's NumPy twin
//   utils.cdist(utils.all_diffs(a, b), metric)          src/utils.py:313-341
// for metric in {squaredeuclidean, euclidean, l1}.  One thread per (i, j) pair evaluates the NumPy summation
// order (exact.cuh); the two row panels are staged in shared memory with a +1 pitch so a warp's 32 different
// b-rows hit 32 different banks and the a-row is a broadcast.  This is the materialising debug / mining path
// (the host-side triplet miners of the reference consume the full matrix, src/base_model.py:271-272); the
// loss and retrieval paths never materialise it.
#include <cuda_runtime.h>

#include "common.cuh"
#include "exact.cuh"
#include "sqdist.h"

namespace mmsim {
namespace sqdist {

constexpr int TI = 16, TJ = 64, THREADS = 256;

template <int METRIC, bool STAGED>
__global__ void __launch_bounds__(THREADS)
sqdist_exact_kernel(const float* __restrict__ A, int64_t M, const float* __restrict__ B, int64_t N, int D,
                    float* __restrict__ out, int64_t ld) {
  extern __shared__ float sm[];
  const int P = D + 1;
  const int64_t i0 = int64_t(blockIdx.y) * TI, j0 = int64_t(blockIdx.x) * TJ;
  if (STAGED) {
    for (int x = threadIdx.x; x < TI * D; x += THREADS) {
      const int r = x / D, c = x - r * D;
      sm[r * P + c] = i0 + r < M ? A[(i0 + r) * D + c] : 0.f;
    }
    for (int x = threadIdx.x; x < TJ * D; x += THREADS) {
      const int r = x / D, c = x - r * D;
      sm[(TI + r) * P + c] = j0 + r < N ? B[(j0 + r) * D + c] : 0.f;
    }
    __syncthreads();
  }
  const int tj = threadIdx.x & (TJ - 1);
  for (int ti = threadIdx.x / TJ; ti < TI; ti += THREADS / TJ) {
    const int64_t i = i0 + ti, j = j0 + tj;
    if (i >= M || j >= N) continue;
    const float* a = STAGED ? sm + ti * P : A + i * D;
    const float* b = STAGED ? sm + (TI + tj) * P : B + j * D;
    out[i * ld + j] = exact_finish<METRIC>(exact_reduce<METRIC>(a, b, D));
  }
}

template <int METRIC>
static int launch(const float* A, int64_t M, const float* B, int64_t N, int D, float* out, int64_t ld, cudaStream_t s) {
  const dim3 grid(unsigned((N + TJ - 1) / TJ), unsigned((M + TI - 1) / TI));
  const size_t smem = size_t(TI + TJ) * (D + 1) * 4;
  if (smem <= 96 * 1024) {
    MMSIM_CUDA_CHECK(cudaFuncSetAttribute(sqdist_exact_kernel<METRIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    sqdist_exact_kernel<METRIC, true><<<grid, THREADS, smem, s>>>(A, M, B, N, D, out, ld);
  } else {
    sqdist_exact_kernel<METRIC, false><<<grid, THREADS, 0, s>>>(A, M, B, N, D, out, ld);
  }
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

int run(const float* A, int64_t M, const float* B, int64_t N, int64_t D, int metric, float* out, int64_t ld, cudaStream_t s) {
  MMSIM_REQUIRE(A && B && out, MMSIM_ERR_ARG, "sqdist: null pointer argument");
  MMSIM_REQUIRE(M >= 0 && N >= 0 && D >= 0 && ld >= N, MMSIM_ERR_ARG, "sqdist: bad shape M=%lld N=%lld D=%lld ld=%lld",
                (long long)M, (long long)N, (long long)D, (long long)ld);
  MMSIM_REQUIRE(D < (1 << 24), MMSIM_ERR_ARG, "sqdist: D too large");
  MMSIM_REQUIRE((M + TI - 1) / TI <= 65535, MMSIM_ERR_ARG, "sqdist: M too large for one launch (chunk it)");
  if (M == 0 || N == 0) return MMSIM_OK;
  switch (metric) {
    case kSquaredEuclidean: return launch<kSquaredEuclidean>(A, M, B, N, int(D), out, ld, s);
    case kEuclidean: return launch<kEuclidean>(A, M, B, N, int(D), out, ld, s);
    case kL1: return launch<kL1>(A, M, B, N, int(D), out, ld, s);
  }
  set_error("sqdist: unknown metric %d (0 squaredeuclidean, 1 euclidean, 2 l1)", metric);
  return MMSIM_ERR_ARG;
}

}  // namespace sqdist
}  // namespace mmsim
