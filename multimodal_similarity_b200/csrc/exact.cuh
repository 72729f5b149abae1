// The canonical exact fp32 distance of the reference's NumPy path, bit for bit.
//
// Reference: utils.cdist(utils.all_diffs(a, b)) (src/utils.py:313-341) and
// np.linalg.norm(q[None] - db, axis=1) (src/utils.py:73) both evaluate, in float32 and without FMA,
//   d_k = fl(a_k - b_k);  s_k = fl(d_k * d_k);  S = pairwise(s, D);   [dist = fl(sqrt(S))]
// where pairwise() is NumPy's add.reduce order:
//   n < 8    : sequential
//   n <= 128 : 8 strided accumulators r[k] += s[i+k], combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n%8 tail
//   n > 128  : n2 = n/2 - (n/2)%8 ; pairwise(s[:n2]) + pairwise(s[n2:])
// __fsub_rn/__fmul_rn/__fadd_rn are never contracted into FMAs, so the bits match NumPy's.
#pragma once
#include <cuda_runtime.h>

namespace mmsim {

enum Metric { kSquaredEuclidean = 0, kEuclidean = 1, kL1 = 2 };

template <int METRIC>
__device__ __forceinline__ float exact_term(float a, float b) {
  const float d = __fsub_rn(a, b);
  if (METRIC == kL1) return fabsf(d);
  return __fmul_rn(d, d);
}

// One leaf (n <= 128) of the pairwise tree.  `a` and `b` may live in any address space.
template <int METRIC>
__device__ __forceinline__ float exact_leaf(const float* __restrict__ a, const float* __restrict__ b, int n) {
  if (n < 8) {
    float r = 0.f;  // NumPy starts from s[0]; 0 + s[0] == s[0] exactly (s >= 0, and -0 cannot occur for squares/abs)
    if (n > 0) r = exact_term<METRIC>(a[0], b[0]);
    for (int i = 1; i < n; ++i) r = __fadd_rn(r, exact_term<METRIC>(a[i], b[i]));
    return r;
  }
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = exact_term<METRIC>(a[k], b[k]);
  int i = 8;
  for (; i + 8 <= n; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], exact_term<METRIC>(a[i + k], b[i + k]));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, exact_term<METRIC>(a[i], b[i]));
  return res;
}

// Full pairwise tree, iterative (explicit stack; depth <= 24 covers any int n).
template <int METRIC>
__device__ float exact_reduce(const float* __restrict__ a, const float* __restrict__ b, int n) {
  if (n <= 128) return exact_leaf<METRIC>(a, b, n);
  // Post-order evaluation of the binary split tree.  Each stack frame is a pending right half.
  int off_stack[24], len_stack[24];
  float val_stack[24];
  unsigned char state[24];
  int sp = 0;
  off_stack[0] = 0;
  len_stack[0] = n;
  state[0] = 0;
  float ret = 0.f;
  while (sp >= 0) {
    const int off = off_stack[sp], len = len_stack[sp];
    if (len <= 128) {
      ret = exact_leaf<METRIC>(a + off, b + off, len);
      --sp;
      continue;
    }
    int n2 = len / 2;
    n2 -= n2 % 8;
    if (state[sp] == 0) {  // descend left
      state[sp] = 1;
      ++sp;
      off_stack[sp] = off;
      len_stack[sp] = n2;
      state[sp] = 0;
    } else if (state[sp] == 1) {  // left done -> descend right
      val_stack[sp] = ret;
      state[sp] = 2;
      ++sp;
      off_stack[sp] = off + n2;
      len_stack[sp] = len - n2;
      state[sp] = 0;
    } else {  // both done
      ret = __fadd_rn(val_stack[sp], ret);
      --sp;
    }
  }
  return ret;
}


// ---- cooperative form: 8 adjacent lanes (sub = lane & 7) evaluate one pair; lane `sub` owns NumPy's accumulator r[sub].
// Loads are 32-byte-sector coalesced across the 8 lanes and independent of each other (memory-level parallelism), the
// summation order -- and therefore every bit -- is the same as exact_reduce().  All 32 lanes of the warp must call
// these together (full-mask shuffles); every lane of a group returns the result.
template <int METRIC>
__device__ __forceinline__ float exact_leaf_8(const float* __restrict__ a, const float* __restrict__ b, int n, int sub) {
  if (n < 8) {
    float r = 0.f;
    if (n > 0) r = exact_term<METRIC>(a[0], b[0]);
    for (int i = 1; i < n; ++i) r = __fadd_rn(r, exact_term<METRIC>(a[i], b[i]));
    return r;
  }
  float r = exact_term<METRIC>(a[sub], b[sub]);
  int i = 8;
#pragma unroll 8
  for (; i + 8 <= n; i += 8) r = __fadd_rn(r, exact_term<METRIC>(a[i + sub], b[i + sub]));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));   // (r0+r1) (r2+r3) (r4+r5) (r6+r7)
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));   // ((r0+r1)+(r2+r3)) ((r4+r5)+(r6+r7))
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
  for (; i < n; ++i) r = __fadd_rn(r, exact_term<METRIC>(a[i], b[i]));
  return r;
}

template <int METRIC>
__device__ float exact_reduce_8(const float* __restrict__ a, const float* __restrict__ b, int n, int sub) {
  if (n <= 128) return exact_leaf_8<METRIC>(a, b, n, sub);
  int off_stack[24], len_stack[24];
  float val_stack[24];
  unsigned char state[24];
  int sp = 0;
  off_stack[0] = 0;
  len_stack[0] = n;
  state[0] = 0;
  float ret = 0.f;
  while (sp >= 0) {
    const int off = off_stack[sp], len = len_stack[sp];
    if (len <= 128) {
      ret = exact_leaf_8<METRIC>(a + off, b + off, len, sub);
      --sp;
      continue;
    }
    int n2 = len / 2;
    n2 -= n2 % 8;
    if (state[sp] == 0) {
      state[sp] = 1;
      ++sp;
      off_stack[sp] = off;
      len_stack[sp] = n2;
      state[sp] = 0;
    } else if (state[sp] == 1) {
      val_stack[sp] = ret;
      state[sp] = 2;
      ++sp;
      off_stack[sp] = off + n2;
      len_stack[sp] = len - n2;
      state[sp] = 0;
    } else {
      ret = __fadd_rn(val_stack[sp], ret);
      --sp;
    }
  }
  return ret;
}

template <int METRIC>
__device__ __forceinline__ float exact_finish(float s) {
  if (METRIC == kEuclidean) return __fsqrt_rn(__fadd_rn(s, 1e-12f));  // src/utils.py:337 (float32 + python float -> float32)
  return s;
}

// Retrieval distance: sqrt of the squared sum (np.linalg.norm, src/utils.py:73).
__device__ __forceinline__ float exact_l2(const float* __restrict__ a, const float* __restrict__ b, int n) {
  return __fsqrt_rn(exact_reduce<kSquaredEuclidean>(a, b, n));
}

}  // namespace mmsim
