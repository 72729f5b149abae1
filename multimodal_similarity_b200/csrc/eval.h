#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
namespace mmsim {
namespace eval {
int run(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
        double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, cudaStream_t s);
int large_workspace_bytes(int64_t N, int64_t nq, size_t* out);
int run_large(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
              double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, void* ws,
              size_t ws_bytes, cudaStream_t s);
}
}  // namespace mmsim
