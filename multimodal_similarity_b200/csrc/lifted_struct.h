#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
namespace mmsim {
namespace lifted_struct {
int workspace_bytes(int64_t N, size_t* out);
int run(const float* E, const int* labels, int64_t N, int64_t D, float margin, float* loss, float* dE, void* ws, size_t ws_bytes,
        cudaStream_t s);
}
}  // namespace mmsim
