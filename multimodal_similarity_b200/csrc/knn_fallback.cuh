// Exact fallback of the kNN retrieval (included by knn_tc.cu): what happens to the queries the certificate of the
// re-rank kernel could not prove exact.  Reference semantics: np.argsort(np.linalg.norm(q - G, axis=1))[:k]
// (src/utils.py:73-74) -- the result must be that no matter how the gallery is ordered or how many ties it holds.
//
//   tier 1  (common)  the k-th exact candidate distance dk of an uncertified query is an UPPER bound of its true k-th
//           distance (the candidates are real rows), so every row of the true top-k has an approximate key below
//           tau' = (dk + rounding-error norms)^2 - |q~|^2 + delta.  The uncertified queries are compacted into their own
//           operand block and swept AGAIN on the tensor cores (knn_tc_kernel<MODE_RESWEEP>, geometry derived on the
//           device from the uncertified count: no host round trip) with the fixed threshold tau'; fb_select_kernel then
//           recomputes EVERY logged row exactly and keeps the k smallest (distance, index) pairs.  No certificate is
//           needed: the logged set contains the true top-k by construction.
//   tier 2  (dk = +inf: fewer than k candidates were found; or a tier-1 log overflowed: massive ties / duplicates)
//           streaming exact top-k over the whole gallery, no capacity limit anywhere: FB2_CHUNKS CTAs per query each keep
//           the k best of their chunk (block-level threshold + periodic compaction), a merge CTA combines them.  One
//           wave of FB2_WAVE queries is enqueued with every call; status[1] - status[2] tells the host how many are left,
//           and mmsim_knn_finish_f32 runs further waves (the host wrappers loop until none are left).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "exact.cuh"

namespace mmsim {
namespace knn {

constexpr int FB_THREADS = 256;
constexpr int FB_CAP = 2048;       // staging entries of the block-level exact top-k
constexpr int FB_BATCH = 512;      // pushes between two capacity checks (FB_CAP - FB_BATCH >= any k)
constexpr int FB2_CHUNKS = 32;     // tier 2: gallery chunks (CTAs) per query
constexpr int FB2_WAVE = 128;      // tier 2: queries per wave

// Geometry of the tier-1 re-sweep, derived from the uncertified count U on the device (sweep kernel and select kernel
// must agree): query blocks of 128, and as many gallery splits as fill the persistent grid in one wave.
struct DynGeom {
  int nq, n_qblocks, n_splits, tiles_per_split;
};
__host__ __device__ inline DynGeom resweep_geometry(int U, int n_tiles, int num_ctas, int64_t slot_budget) {
  DynGeom g;
  g.nq = U;
  g.n_qblocks = (U + 127) / 128;
  if (g.n_qblocks == 0) {
    g.n_splits = 1;
    g.tiles_per_split = n_tiles;
    return g;
  }
  int64_t s = num_ctas / g.n_qblocks;                               // one wave
  const int64_t by_budget = slot_budget / (int64_t(g.n_qblocks) * 128);   // log / counter slots of the main plan
  if (s > by_budget) s = by_budget;
  if (s > n_tiles / 4) s = n_tiles / 4;                             // at least 4 tiles per split
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  g.tiles_per_split = int((n_tiles + s - 1) / s);
  g.n_splits = (n_tiles + g.tiles_per_split - 1) / g.tiles_per_split;
  return g;
}

// ------------------------------------------------------------------------------------------------ block-level top-k
// The k smallest (distance, index) pairs of a stream, one CTA.  Pairs below the current threshold are appended to a
// staging array; when it could overflow the array is sorted, cut to k, and the k-th pair becomes the threshold.
struct TopKSmem {
  float key[FB_CAP];
  int idx[FB_CAP];
  int count;
  float thr_d;
  int thr_i;
};

__device__ __forceinline__ void topk_reset(TopKSmem& s) {
  __syncthreads();
  if (threadIdx.x == 0) {
    s.count = 0;
    s.thr_d = kInf;
    s.thr_i = 0x7fffffff;
  }
  __syncthreads();
}

// any thread; at most FB_BATCH calls per CTA between two topk_maybe_compact
__device__ __forceinline__ void topk_push(TopKSmem& s, float d, int i) {
  const float td = s.thr_d;
  if (d < td || (d == td && d < kInf && i < s.thr_i)) {
    const int pos = atomicAdd(&s.count, 1);
    s.key[pos] = d;
    s.idx[pos] = i;
  }
}

// all threads of the CTA: sort the staged pairs by (distance, index), keep the k smallest, set the threshold
__device__ void topk_compact(TopKSmem& s, int k) {
  __syncthreads();
  const int n = s.count;
  int P = 32;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) {
    s.key[i] = kInf;
    s.idx[i] = 0x7fffffff;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const float a = s.key[lo], b = s.key[hi];
        const int ia = s.idx[lo], ib = s.idx[hi];
        const bool gt = (a > b) || (a == b && ia > ib);
        if (gt == asc) {
          s.key[lo] = b; s.key[hi] = a;
          s.idx[lo] = ib; s.idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    if (n >= k) {
      s.thr_d = s.key[k - 1];
      s.thr_i = s.idx[k - 1];
      s.count = k;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void topk_maybe_compact(TopKSmem& s, int k) {
  __syncthreads();
  if (s.count > FB_CAP - FB_BATCH) topk_compact(s, k);   // uniform: every thread reads the same count after the barrier
}

// ------------------------------------------------------------------------------------------------ tier 1: prepare
// One warp per uncertified query: fp16 operand row (same arithmetic as prep_rows_kernel), the fixed threshold tau' of the
// re-sweep, and the routing of queries without a finite bound to tier 2.
struct FbLists {
  const int* count;        // [1] uncertified queries U (device)
  const int* query;        // [U] caller's query index
  float* bound;            // [U] k-th exact candidate distance (upper bound of the true k-th distance); +inf: none
  int* fb2_list;           // tier-2 queue of slots
  int* status;             // status[1] = tier-2 queue length, status[2] = tier-2 slots processed
  int cap;                 // capacity of the lists
};

__global__ void __launch_bounds__(FB_THREADS)
fb_prepare_kernel(const float* __restrict__ Q, int D, int Dp, FbLists L, const float* __restrict__ gstats, float delta_coeff,
                  __half* __restrict__ fb_qh, float* __restrict__ fb_ladder,
                  int force_tier2 /* test hook (MMSIM_KNN_FORCE_TIER2=1): every query takes the streaming scan */) {
  const int U = min(*L.count, L.cap);
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int rows = (U + 127) / 128 * 128;
  for (int slot = warp0; slot < rows; slot += nwarps) {
    __half* hr = fb_qh + size_t(slot) * Dp;
    if (slot >= U) {                               // padding rows of the last block: zero operands, closed threshold
      for (int c = lane; c < Dp; c += 32) hr[c] = __float2half_rn(0.f);
      if (lane == 0) *reinterpret_cast<float4*>(fb_ladder + size_t(slot) * 4) = make_float4(-kInf, -kInf, -kInf, -kInf);
      continue;
    }
    const float* xr = Q + size_t(L.query[slot]) * D;
    float s = 0.f, e = 0.f;
    for (int c = lane; c < Dp; c += 32) {
      const float v = c < D ? xr[c] : 0.f;
      const __half h = __float2half_rn(v);
      const float r = __half2float(h);
      s = fmaf(r, r, s);
      const float d = v - r;
      e = fmaf(d, d, e);
      if (isinf(r * -2.0f) || isinf(r)) e = kInf;
      hr[c] = __float2half_rn(r * -2.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if (lane == 0) {
      const float en = sqrtf(e) * 1.0001f;
      const float dk = L.bound[slot];
      // rows with exact distance <= dk:  |q~ - g~| <= d (1 + 1e-5) + |q - q~| + max |g - g~|  and
      // key = |g~|^2 - 2 q~.g~ <= |q~ - g~|^2 - |q~|^2 + delta   (delta: fp32 accumulation error, as in the certificate)
      const float reach = dk * 1.00001f + en + gstats[0];
      float tau = reach * reach * 1.000001f - s + delta_coeff * (s + gstats[1]);
      tau += 1e-6f * fabsf(tau) + 1e-30f;
      if (!(tau < kInf) || force_tier2) {           // no finite bound (fewer than k candidates, values outside fp16): tier 2
        tau = -kInf;
        L.bound[slot] = kInf;
        const int pos = atomicAdd(&L.status[1], 1);
        if (pos < L.cap) L.fb2_list[pos] = slot;
      }
      *reinterpret_cast<float4*>(fb_ladder + size_t(slot) * 4) = make_float4(-kInf, -kInf, -kInf, tau);
    }
  }
}

// Where the exact top-k of a fallback query goes: row `query` of the caller's [nq, k] arrays, or row `slot` of a compact
// [cap, k] block (gallery-shard mode: the per-shard lists are exchanged and merged afterwards).
struct FbOut {
  float* dist;
  int* idx;
  int compact;
};

__device__ __forceinline__ void fb_write_out(const TopKSmem& s, const FbOut& o, int row, int k) {
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const bool ok = r < s.count;
    o.dist[size_t(row) * k + r] = ok ? s.key[r] : kInf;
    o.idx[size_t(row) * k + r] = ok ? s.idx[r] : -1;
  }
}

// Exact distances of up to 32 rows per CTA round (8 lanes per row: full 32-byte sectors), pushed into the block top-k.
// row_of(e) gives the gallery row of entry e (or -1); every thread of the CTA must take part.
template <typename RowOf>
__device__ __forceinline__ void fb_stream_rows(TopKSmem& s, const float* qs, const float* __restrict__ G, int D, int64_t n_entries,
                                               int self, int k, RowOf row_of) {
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;      // 32 groups of 8 lanes
  constexpr int PER_ROUND = FB_THREADS / 8;
  int since = 0;
  for (int64_t e0 = 0; e0 < n_entries; e0 += PER_ROUND) {
    const int64_t e = e0 + grp;
    const int row = e < n_entries ? row_of(e) : -1;
    const bool live = row >= 0 && row != self;
    const float d2 = exact_reduce_8<kSquaredEuclidean>(qs, G + size_t(live ? row : 0) * D, D, sub);
    if (live && sub == 0) topk_push(s, __fsqrt_rn(d2), row);
    since += PER_ROUND;
    if (since > FB_BATCH - PER_ROUND) {
      topk_maybe_compact(s, k);
      since = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ tier 1: select
__global__ void __launch_bounds__(FB_THREADS)
fb_select_kernel(const float* __restrict__ Q, const float* __restrict__ G, int D, const uint2* __restrict__ log, int logcap,
                 const int* __restrict__ log_cnt, int n_tiles, int num_ctas, int64_t slot_budget, FbLists L, int k,
                 int exclude_self, int64_t self_offset, FbOut out) {
  __shared__ TopKSmem s;
  extern __shared__ float fb_qs[];
  const int U = min(*L.count, L.cap);
  const DynGeom geo = resweep_geometry(U, n_tiles, num_ctas, slot_budget);
  for (int slot = blockIdx.x; slot < U; slot += gridDim.x) {
    if (!(L.bound[slot] < kInf)) continue;          // queued for tier 2 by fb_prepare_kernel
    const size_t l0 = size_t(slot) * geo.n_splits;
    bool overflow = false;
    for (int sp = 0; sp < geo.n_splits; ++sp) overflow |= log_cnt[l0 + sp] > logcap;
    if (overflow) {                                  // uniform over the CTA
      if (threadIdx.x == 0) {
        const int pos = atomicAdd(&L.status[1], 1);
        if (pos < L.cap) L.fb2_list[pos] = slot;
      }
      continue;
    }
    const int qo = L.query[slot];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) fb_qs[c] = Q[size_t(qo) * D + c];
    topk_reset(s);
    const int self = exclude_self ? int(self_offset + qo) : -1;
    for (int sp = 0; sp < geo.n_splits; ++sp) {
      const uint2* ls = log + (l0 + sp) * logcap;
      fb_stream_rows(s, fb_qs, G, D, log_cnt[l0 + sp], self, k, [&](int64_t e) { return int(__ldg(&ls[e].y)); });
    }
    topk_compact(s, k);
    fb_write_out(s, out, out.compact ? slot : qo, k);
  }
}

// ------------------------------------------------------------------------------------------------ tier 2
__device__ __forceinline__ int fb2_wave(const int* status, int cap) { return min(FB2_WAVE, min(status[1], cap) - status[2]); }

// grid (FB2_CHUNKS, FB2_WAVE): CTA (c, y) keeps the k best rows of chunk c for the y-th query of this wave
__global__ void __launch_bounds__(FB_THREADS)
fb2_scan_kernel(const float* __restrict__ Q, const float* __restrict__ G, int64_t ng, int D, FbLists L, int k, int exclude_self,
                int64_t self_offset, float* __restrict__ part_dist, int* __restrict__ part_idx) {
  __shared__ TopKSmem s;
  extern __shared__ float fb_qs[];
  if (int(blockIdx.y) >= fb2_wave(L.status, L.cap)) return;
  const int slot = L.fb2_list[L.status[2] + blockIdx.y];
  const int qo = L.query[slot];
  for (int c = threadIdx.x; c < D; c += blockDim.x) fb_qs[c] = Q[size_t(qo) * D + c];
  topk_reset(s);
  const int64_t per = (ng + FB2_CHUNKS - 1) / FB2_CHUNKS;
  const int64_t g0 = min(ng, int64_t(blockIdx.x) * per), g1 = min(ng, g0 + per);
  const int self = exclude_self ? int(self_offset + qo) : -1;
  fb_stream_rows(s, fb_qs, G, D, g1 - g0, self, k, [&](int64_t e) { return int(g0 + e); });
  topk_compact(s, k);
  const size_t o = (size_t(blockIdx.y) * FB2_CHUNKS + blockIdx.x) * k;
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const bool ok = r < s.count;
    part_dist[o + r] = ok ? s.key[r] : kInf;
    part_idx[o + r] = ok ? s.idx[r] : -1;
  }
}

// grid FB2_WAVE: combine the chunks' lists of one query
__global__ void __launch_bounds__(FB_THREADS)
fb2_merge_kernel(FbLists L, int k, const float* __restrict__ part_dist, const int* __restrict__ part_idx, FbOut out) {
  __shared__ TopKSmem s;
  if (int(blockIdx.x) >= fb2_wave(L.status, L.cap)) return;
  const int slot = L.fb2_list[L.status[2] + blockIdx.x];
  topk_reset(s);
  const size_t o = size_t(blockIdx.x) * FB2_CHUNKS * k;
  const int n = FB2_CHUNKS * k;
  int since = 0;
  for (int e0 = 0; e0 < n; e0 += FB_THREADS) {
    const int e = e0 + threadIdx.x;
    if (e < n && part_idx[o + e] >= 0) topk_push(s, part_dist[o + e], part_idx[o + e]);
    since += FB_THREADS;
    if (since > FB_BATCH - FB_THREADS) {
      topk_maybe_compact(s, k);
      since = 0;
    }
  }
  topk_compact(s, k);
  fb_write_out(s, out, out.compact ? slot : L.query[slot], k);
}

__global__ void fb2_advance_kernel(int* status, int cap) { status[2] += max(0, fb2_wave(status, cap)); }

// ------------------------------------------------------------------------------------------------ shard mode helpers
// flag[q] >= 0 (the merged k-th distance of an uncertified query; +inf allowed) -> ordered list (ascending query index: every
// rank builds the SAME list, the compact blocks they exchange line up slot by slot).  One block.
__global__ void __launch_bounds__(1024)
fb_list_from_flags_kernel(const float* __restrict__ flag, int nq, int cap, int* __restrict__ query, float* __restrict__ bound,
                          int* __restrict__ status) {
  __shared__ int part[1024];
  const int per = (nq + 1023) / 1024, lo = min(nq, int(threadIdx.x) * per), hi = min(nq, lo + per);
  int mine = 0;
  for (int i = lo; i < hi; ++i) mine += flag[i] >= 0.f ? 1 : 0;
  part[threadIdx.x] = mine;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int x = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += x;
    __syncthreads();
  }
  int pos = part[threadIdx.x] - mine;
  for (int i = lo; i < hi; ++i) {
    if (flag[i] >= 0.f) {
      if (pos < cap) {
        query[pos] = i;
        bound[pos] = flag[i];
      }
      ++pos;
    }
  }
  if (threadIdx.x == 1023) {
    status[0] = part[1023];            // may exceed cap: the caller then falls back for the whole call
    status[1] = 0;
    status[2] = 0;
  }
}

}  // namespace knn
}  // namespace mmsim
