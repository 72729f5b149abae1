#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace mmsim {
namespace mining {
int run_mask(const float* dist, int64_t n, int64_t ld, const int* labels, const int* pairs, int64_t m, float alpha,
             uint32_t* mask, int* count, cudaStream_t s);
int run_pick(const float* dist, int64_t n, int64_t ld, const int* labels, const int* picks, int64_t p, float alpha,
             int* neg_idx, cudaStream_t s);
}
}  // namespace mmsim
