#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace mmsim {
namespace project {
int run(const float* X, int64_t N, int64_t K, const float* W, const float* b, int64_t E, int normalized, float eps, float* out,
        cudaStream_t s);
}
}  // namespace mmsim
