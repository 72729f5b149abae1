#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
namespace mmsim {
namespace eval {
// K5 (eval.cu): one CTA per query, distances + bitonic sort + metrics fused; N <= 16385, no workspace
int run(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
        double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, cudaStream_t s);

// workspace form (eval_large.cu): tiled exact distances for a batch of queries, then per query either the shared-memory
// radix sort + metrics kernel (eval_fast.cu, N <= kSortMaxN) or a segmented device sort + the streaming metrics kernel
enum : int { kPathAuto = 0, kPathSmemSort = 1, kPathSegmentedSort = 2 };
constexpr int kSortMaxN = 24576;
int large_workspace_bytes(int64_t N, int64_t nq, size_t* out);
int run_large(const float* E, const int* labels, const int* cls, int64_t N, int64_t D, int C, const int* queries, int64_t nq,
              double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank, void* ws,
              size_t ws_bytes, cudaStream_t s, int path = kPathAuto);

// pieces of the workspace form (eval_fast.cu)
struct LeafPlan { int n; int lo[2]; int len[2]; };     // NumPy's pairwise-summation tree when it has at most two leaves
bool leaf_plan(int D, LeafPlan* lp);
int launch_distances(const float* E, int64_t N, int64_t D, const int* queries, int nqb, uint32_t* keys, int64_t ldk, int* vals,
                     cudaStream_t s);
bool sort_path_fits(int64_t N, int C);
int launch_sort_metrics(const uint32_t* keys, int64_t ldk, const int* labels, const int* cls, int64_t N, int C, const int* queries,
                        int nqb, double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank,
                        cudaStream_t s);
int confusion(const int* hist, const int* depth, const int* npos, const int* qcls, int64_t nq, int C, float* cm, int* count,
              int* lists, const int* list_off, cudaStream_t s);
}
}  // namespace mmsim
