#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace mmsim {
namespace merge {
int run(const float* dist_parts, const int* idx_parts, int64_t part_stride, const int64_t* idx_base, int parts, int64_t nq,
        int k_in, int k, const float* lb_parts, int64_t lb_stride, float* out_dist, void* out_idx, int* status,
        cudaStream_t s, float* out_flag = nullptr, const int* count_ptr = nullptr, const int* row_map = nullptr, int idx32 = 0);
}
}  // namespace mmsim
