#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace mmsim {
namespace sqdist {
int run(const float* A, int64_t M, const float* B, int64_t N, int64_t D, int metric, float* out, int64_t ld, cudaStream_t s);
}
}  // namespace mmsim
