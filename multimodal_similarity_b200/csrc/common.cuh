// Shared host/device helpers for the mmsim C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include <nvtx3/nvToolsExt.h>

#include "../../include/mmsim.h"

namespace mmsim {

// Error codes (MMSIM_OK, MMSIM_ERR_*) come from the public header.
void set_error(const char* fmt, ...);
// every kernel launch of the library is followed by MMSIM_CUDA_CHECK(launched()): cudaGetLastError() + one tick of the
// process-wide launch counter (mmsim_kernel_launches; bench.py reports it as gpu_launches)
cudaError_t launched();

#define MMSIM_CUDA_CHECK(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::mmsim::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MMSIM_ERR_CUDA;                                                           \
    }                                                                                           \
  } while (0)

#define MMSIM_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ::mmsim::set_error(__VA_ARGS__);   \
      return (code);                     \
    }                                    \
  } while (0)

// NVTX range over the host-side enqueue of an entry point or a phase (header-only NVTX 3: a no-op costing one indirect call
// unless a tool -- nsys, ncu --nvtx -- injected itself); the kernels launched inside inherit it in the tool's timeline.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define MMSIM_CAT2(a, b) a##b
#define MMSIM_CAT(a, b) MMSIM_CAT2(a, b)
#define MMSIM_RANGE(name) ::mmsim::NvtxRange MMSIM_CAT(_mmsim_range_, __LINE__)(name)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr float kInf = __builtin_huge_valf();

// ---- warp-level sorting primitives on (key, payload) pairs held one per lane -----------------------------
// Compare-exchange with the lane `lane ^ mask`; `up` = this lane keeps the smaller element.
__device__ __forceinline__ void cx_pair(float& k, int& v, uint32_t mask, bool keep_small) {
  const float ok = __shfl_xor_sync(0xffffffffu, k, mask);
  const int ov = __shfl_xor_sync(0xffffffffu, v, mask);
  const bool other_smaller = (ok < k) || (ok == k && ov < v);
  if (other_smaller == keep_small) {
    k = ok;
    v = ov;
  }
}
// Full bitonic sort of 32 pairs across the warp, ascending by (key, payload).
__device__ __forceinline__ void warp_sort32(float& k, int& v, uint32_t lane) {
#pragma unroll
  for (uint32_t size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      const bool asc = (lane & size) == 0 || size == 32;
      const bool lower = (lane & stride) == 0;
      cx_pair(k, v, stride, asc == lower);
    }
  }
}
// Sort a bitonic sequence of 32 pairs ascending (5 butterfly stages).
__device__ __forceinline__ void warp_bitonic_merge32(float& k, int& v, uint32_t lane) {
#pragma unroll
  for (uint32_t stride = 16; stride > 0; stride >>= 1) cx_pair(k, v, stride, (lane & stride) == 0);
}
// Merge-split: X (row, ascending across lanes) and Y (carry, ascending) -> X = smallest 32 (ascending),
// Y = largest 32 (ascending).  Ties prefer the element already in X.
__device__ __forceinline__ void warp_merge_split32(float& xk, int& xv, float& yk, int& yv, uint32_t lane) {
  const float rk = __shfl_sync(0xffffffffu, yk, 31 - lane);
  const int rv = __shfl_sync(0xffffffffu, yv, 31 - lane);
  const bool take_x = (xk < rk) || (xk == rk && xv <= rv);
  const float lo_k = take_x ? xk : rk, hi_k = take_x ? rk : xk;
  const int lo_v = take_x ? xv : rv, hi_v = take_x ? rv : xv;
  xk = lo_k; xv = lo_v; yk = hi_k; yv = hi_v;
  warp_bitonic_merge32(xk, xv, lane);
  warp_bitonic_merge32(yk, yv, lane);
}

}  // namespace mmsim
