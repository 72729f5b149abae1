// Per-query metrics of the leave-one-out evaluation, computed in ONE streaming pass over a query's ranking -- shared by the
// shared-memory sort path (eval_fast.cu, ranking in shared memory) and the segmented-sort path (eval_large.cu, ranking in
// global memory).  The loop body of utils.evaluate / utils.evaluate_simple (src/utils.py:83-229):
//   ap            sklearn average_precision_score of score = fl32(max(dist) - dist) (:78-79): thresholds at distinct scores,
//                 so every positive counts with the precision at the END of its tie group
//   first         rank of the first ranked label equal to the query's                      -> recall_at_K (:257-266)
//   depth, hist   the prefix length precision_at_recall walks (:231-255) and the per-class counts inside it
// The ranking holds all N rows in the ORIGINAL numbering: N - 1 real rows ordered by (distance, index), then the query's
// own row (key 0xffffffff).  Positions in the gallery-with-row-i-deleted are jp = j - (j > i).
//
// REFERENCE QUIRK kept for parity (aligned == 0): the reference indexes the FULL label array with positions of the
// deleted gallery (src/utils.py:128,132,185,190), so precision@recall / recall@K see the label of row jp, not row j.
// AP uses the correctly deleted labels (:78).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mmsim {
namespace eval {

// block-wide inclusive sum scan / exclusive max scan of one int per thread; `carry` is added / maxed in and the block
// total (sum) or block maximum (max) comes back through it.  NT threads, all of them must call.
template <int NT>
__device__ __forceinline__ int block_scan_sum(int x, int* wsum, int& carry) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < NW ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const int out = incl + (warp ? wsum[warp - 1] : 0) + carry;
  const int total = wsum[31];
  __syncthreads();
  carry += total;
  return out;
}
template <int NT>
__device__ __forceinline__ int block_scan_max_excl(int x, int* wmax, int& carry) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = max(incl, y);
  }
  if (lane == 31) wmax[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < NW ? wmax[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w = max(w, y);
    }
    wmax[lane] = w;
  }
  __syncthreads();
  int excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
  excl = max(excl, max(warp ? wmax[warp - 1] : 0, carry));
  const int total = wmax[31];
  __syncthreads();
  carry = max(carry, total);
  return excl;
}

// kb / v: the query's ranking (key bits of the float32 distances, original row index), N entries, in shared or global
// memory.  lhist: [C] ints of shared memory.  Every thread of the NT-thread block calls this.
template <int NT, typename IdxT>
__device__ __forceinline__ void loo_metrics(const uint32_t* kb, const IdxT* v, const int* __restrict__ labels,
                                            const int* __restrict__ cls, int N, int C, int i, int qn, double alpha, int aligned,
                                            int* lhist, double* __restrict__ out_ap, int* __restrict__ out_npos,
                                            int* __restrict__ out_first, int* __restrict__ out_depth, int* __restrict__ out_hist,
                                            int* __restrict__ out_rank) {
  __shared__ int wbuf[32];
  __shared__ double dred[NT / 32];
  __shared__ int s_first, s_depth, s_cnt, s_m0;
  const int t = threadIdx.x;
  const int n = N - 1, ql = labels[i];
  for (int c = t; c < C; c += NT) lhist[c] = 0;
  if (t == 0) { s_first = n; s_depth = n; s_cnt = 0; s_m0 = 0; }
  __syncthreads();

  // how many ranked labels equal the query's, as the reference forms them (src/utils.py:185-186): independent of the ranking
  {
    int c = 0;
    for (int jp = t; jp < n; jp += NT) c += labels[aligned ? jp + (jp >= i ? 1 : 0) : jp] == ql ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((t & 31) == 0 && c) atomicAdd(&s_cnt, c);
  }
  __syncthreads();
  const int target = int(alpha * double(s_cnt));
  const float dmax = n > 0 ? __uint_as_float(kb[n - 1]) : 0.f;

  int carry_pos = 0, carry_m = 0, carry_end = 0;
  double acc = 0.0;
  int my_first = n;
  for (int base = 0; base < n; base += NT) {
    const int r = base + t;
    const bool in = r < n;
    int pos = 0, m = 0;
    bool is_end = false;
    if (in) {
      const int j = int(v[r]);
      const int jp = j - (j > i ? 1 : 0);
      if (out_rank) out_rank[size_t(qn) * n + r] = jp;
      pos = labels[j] == ql ? 1 : 0;                                       // AP uses the correctly deleted labels (:78)
      m = labels[aligned ? j : jp] == ql ? 1 : 0;
      const float s = __fsub_rn(dmax, __uint_as_float(kb[r]));             // score = max(dist) - dist in float32 (:79)
      is_end = r == n - 1 || __fsub_rn(dmax, __uint_as_float(kb[r + 1])) != s;   // last rank of its tie group
      if (m) my_first = min(my_first, r);
      if (r == 0 && m) s_m0 = 1;
    }
    const int cp = block_scan_sum<NT>(pos, wbuf, carry_pos);               // positives among ranks 0..r
    const int cm = block_scan_sum<NT>(m, wbuf, carry_m);
    const int prev_end = block_scan_max_excl<NT>(is_end ? cp : 0, wbuf, carry_end);   // positives up to the previous group end
    if (in && is_end && cp > prev_end) acc += double(cp - prev_end) * (double(cp) / double(r + 1));
    if (in && target > 0 && m && cm == target) s_depth = r + 1;            // exactly one rank satisfies this
  }
  my_first = __reduce_min_sync(0xffffffffu, my_first);
  if ((t & 31) == 0) atomicMin(&s_first, my_first);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((t & 31) == 0) dred[t >> 5] = acc;
  __syncthreads();
  const int npos = carry_pos;
  if (t == 0) {
    double a = 0.0;
    for (int w = 0; w < NT / 32; ++w) a += dred[w];
    out_ap[qn] = npos > 0 ? a / double(npos) : 0.0;
    out_npos[qn] = npos;
    if (target == 0) s_depth = s_m0 ? n : min(n, 1);                      // src/utils.py:240-250 with int(alpha * #pos) == 0
  }
  __syncthreads();
  const int depth = s_depth;
  // per-class counts of the walked prefix; a warp adds each class it sees once (few classes => same-address atomics)
  for (int base = 0; base < depth; base += NT) {
    const int r = base + t;
    int c = -1;
    if (r < depth) {
      const int j = int(v[r]);
      c = cls[aligned ? j : j - (j > i ? 1 : 0)];
    }
    const unsigned same = __match_any_sync(0xffffffffu, c);
    if (c >= 0 && (t & 31) == __ffs(same) - 1) atomicAdd(&lhist[c], __popc(same));
  }
  __syncthreads();
  for (int c = t; c < C; c += NT) out_hist[size_t(qn) * C + c] = lhist[c];
  if (t == 0) {
    out_first[qn] = s_first;
    out_depth[qn] = depth;
  }
}

}  // namespace eval
}  // namespace mmsim
