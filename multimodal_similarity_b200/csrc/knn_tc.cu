// K4: query x gallery k-nearest-neighbour retrieval for sm_100a.
//
// Replaces the per-query loop of the reference (src/utils.py:55-81 `retrieve_one`: fp32 L2 distance to every
// gallery row + full argsort) with
//   1. prep      fp32 -> fp16 operand copies (+ squared norms of the rounded rows and the rounding error norms)
//   2. knn_tc    persistent, warp-specialised tcgen05 kernel: TMA-staged gallery tiles, UMMA 128x256x16 into
//                double-buffered TMEM accumulators, epilogue warps stream the accumulator out of TMEM and keep a
//                threshold-gated per-query candidate list (KP = 128 best approximate keys) -- the Q x G matrix
//                never exists in memory
//   3. rerank    exact fp32 distance in NumPy summation order (exact.cuh) for the KP candidates, sort by
//                (distance, index), emit top-k, and *certify* the result: every non-candidate's true distance
//                is provably larger than the k-th exact distance (rounding-error norms + triangle inequality)
//   4. fallback  queries that could not be certified are recomputed exactly against the whole gallery
// so the returned distances are bit-identical to the reference's and indices are exact outside exact ties.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "common.cuh"
#include "exact.cuh"
#include "knn.h"
#include "ptx.cuh"

namespace mmsim {
namespace knn {

// ------------------------------------------------------------------------------------------------ prep
// One warp per row.  xh[row, 0..Dp) = fp16(x) * scale (zero padded), norm[row] = sum fp16(x)^2 (unscaled, fp32),
// err[row] = ||x - fp16(x)||_2.  Rows in [n, n_pad) get norm = +inf (masks padded gallery columns).
__global__ void prep_rows_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int D, int Dp, float scale,
                                 __half* __restrict__ xh, float* __restrict__ norm, float* __restrict__ err,
                                 unsigned int* __restrict__ max_stats /* [0]=max err bits, [1]=max norm bits, or null */) {
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  if (row >= n) {
    if (lane == 0) norm[row] = kInf;
    return;
  }
  const float* xr = x + row * D;
  __half* hr = xh + row * Dp;
  float s = 0.f, e = 0.f;
  for (int c = lane; c < Dp; c += 32) {
    float v = c < D ? xr[c] : 0.f;
    const __half h = __float2half_rn(v);
    const float r = __half2float(h);
    s = fmaf(r, r, s);
    const float d = v - r;
    e = fmaf(d, d, e);
    hr[c] = __float2half_rn(r * scale);  // scale is a power of two: exact unless it overflows (then err = inf below)
    if (isinf(r * scale) || isinf(r)) e = kInf;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if (lane == 0) {
    // inflate by a few ulps so the values are upper bounds despite fp32 summation error
    const float en = sqrtf(e) * 1.0001f;
    norm[row] = s;
    if (err) err[row] = en;
    if (max_stats) {
      atomicMax(&max_stats[0], __float_as_uint(en));
      atomicMax(&max_stats[1], __float_as_uint(s));
    }
  }
}

// ------------------------------------------------------------------------------------------------ fused kernel
constexpr int BM = 128;                    // queries per CTA tile (UMMA M, one TMEM lane each)
constexpr int BN = 256;                    // gallery rows per tile (UMMA N)
constexpr int KATOM = 64;                  // fp16 elements per 128-byte swizzle atom
constexpr int NS = 4;                      // gallery smem stages (one K atom of one tile each)
constexpr int A_ATOM_BYTES = BM * 128;     // 16 KiB
constexpr int B_STAGE_BYTES = BN * 128;    // 32 KiB
constexpr int NUM_THREADS = 192;           // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: epilogue

template <int KATOMS>
struct Smem {
  static constexpr int NT = 4;                       // norm ring slots, recycled through their own empty barriers
  static constexpr int CB = KATOMS == 4 ? 24 : 32;   // candidate buffer entries per row
  static constexpr int CBP = CB + 1;                 // padded pitch: conflict-free append and row read
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + KATOMS * A_ATOM_BYTES;
  static constexpr int NORM_OFF = B_OFF + NS * B_STAGE_BYTES;
  static constexpr int CK_OFF = NORM_OFF + NT * BN * 4;
  static constexpr int CI_OFF = CK_OFF + BM * CBP * 4;
  static constexpr int BAR_OFF = CI_OFF + BM * CBP * 4;  // 8-byte aligned: all terms are multiples of 8? checked below
  static constexpr int NUM_BARS = 2 * NS + 2 + 2 + 2 + NT;
  static constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int TOTAL = TMEM_PTR_OFF + 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-byte alignment of the base
  static_assert(BAR_OFF % 8 == 0, "barriers must be 8-byte aligned");
  static_assert(DYN_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

// Merge the row's candidate buffer into its sorted KP-list (global, L2 resident); returns the new threshold
// (the KP-th smallest key).  Executed by the whole warp for one row at a time.
__device__ __noinline__ float flush_row(const float* __restrict__ bufk, const int* __restrict__ bufi, int n,
                                        float* __restrict__ lk, int* __restrict__ li, uint32_t lane) {
  float xk[KP / 32];
  int xv[KP / 32];
#pragma unroll
  for (int j = 0; j < KP / 32; ++j) {  // issue the list loads first, they are the long-latency part
    xk[j] = lk[j * 32 + lane];
    xv[j] = li[j * 32 + lane];
  }
  float yk = int(lane) < n ? bufk[lane] : kInf;
  int yv = int(lane) < n ? bufi[lane] : -1;
  warp_sort32(yk, yv, lane);
#pragma unroll
  for (int j = 0; j < KP / 32; ++j) {
    warp_merge_split32(xk[j], xv[j], yk, yv, lane);
    lk[j * 32 + lane] = xk[j];
    li[j * 32 + lane] = xv[j];
  }
  return __shfl_sync(0xffffffffu, xk[KP / 32 - 1], 31);
}

template <int KATOMS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_g,
              const float* __restrict__ gnorm, int nq, int n_qblocks, int n_splits, int tiles_per_split, int n_tiles,
              float* __restrict__ cand_key, int* __restrict__ cand_idx) {
  using S = Smem<KATOMS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* smem_a = smem + S::A_OFF;
  uint8_t* smem_b = smem + S::B_OFF;
  float* norm_ring = reinterpret_cast<float*>(smem + S::NORM_OFF);
  float* bufk = reinterpret_cast<float*>(smem + S::CK_OFF);
  int* bufi = reinterpret_cast<int*>(smem + S::CI_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* tfull = bars + 2 * NS;       // [2]   MMA -> epilogue (accumulator ready)
  uint64_t* tempty = bars + 2 * NS + 2;  // [2]   epilogue -> MMA (accumulator drained)
  uint64_t* afull = bars + 2 * NS + 4;   //       query tile landed
  uint64_t* aempty = bars + 2 * NS + 5;  //       all MMAs of the item done, query tile may be overwritten
  uint64_t* nempty = bars + 2 * NS + 6;  // [NT]  epilogue finished a tile: its norm slot may be refilled
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::TMEM_PTR_OFF);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_g);
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], BM);
    }
    ptx::mbar_init(afull, 1);
    ptx::mbar_init(aempty, 1);
    for (int i = 0; i < S::NT; ++i) ptx::mbar_init(&nempty[i], BM);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 2 * BN);  // 512 columns: two 128x256 fp32 accumulators
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_items = n_qblocks * n_splits;

  if (warp == 0) {
    // =============================================================== TMA producer (one thread)
    if (lane == 0) {
      uint32_t it = 0, tc = 0, ic = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
        const int split = item / n_qblocks, qb = item - split * n_qblocks;
        const int t0 = split * tiles_per_split;
        const int t1 = min(n_tiles, t0 + tiles_per_split);
        ptx::mbar_wait(aempty, (ic & 1) ^ 1);
        ptx::mbar_expect_tx(afull, KATOMS * A_ATOM_BYTES);
        for (int ka = 0; ka < KATOMS; ++ka)
          ptx::tma_load_2d(smem_a + ka * A_ATOM_BYTES, &tm_q, ka * KATOM, qb * BM, afull);
        for (int t = t0; t < t1; ++t, ++tc) {
          for (int ka = 0; ka < KATOMS; ++ka, ++it) {
            const uint32_t stage = it % NS, phase = (it / NS) & 1;
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], B_STAGE_BYTES + (ka == 0 ? BN * 4 : 0));
            ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tm_g, ka * KATOM, t * BN, &full[stage]);
            if (ka == 0) {
              // the epilogue reads this tile's norms long after it released the accumulator: separate hand-back
              ptx::mbar_wait(&nempty[tc % S::NT], ((tc / S::NT) & 1) ^ 1);
              ptx::bulk_load_1d(norm_ring + (tc % S::NT) * BN, gnorm + size_t(t) * BN, BN * 4, &full[stage]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer (one thread issues, warp waits)
    const uint32_t idesc = ptx::umma_idesc_f16(BM, BN);
    const uint64_t adesc0 = ptx::umma_desc_k128(ptx::smem_u32(smem_a));
    const uint64_t bdesc0 = ptx::umma_desc_k128(ptx::smem_u32(smem_b));
    uint32_t it = 0, tc = 0, ic = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
      const int split = item / n_qblocks;
      const int t0 = split * tiles_per_split;
      const int t1 = min(n_tiles, t0 + tiles_per_split);
      ptx::mbar_wait(afull, ic & 1);
      ptx::tc_fence_after();
      for (int t = t0; t < t1; ++t, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait(&tempty[as], ((tc >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        for (int ka = 0; ka < KATOMS; ++ka, ++it) {
          const uint32_t stage = it % NS, phase = (it / NS) & 1;
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < KATOM / 16; ++k) {
              const uint64_t ad = adesc0 + uint64_t(ka * (A_ATOM_BYTES >> 4) + k * 2);
              const uint64_t bd = bdesc0 + uint64_t(stage * (B_STAGE_BYTES >> 4) + k * 2);
              ptx::umma_f16(tmem_base + as * BN, ad, bd, idesc, (ka | k) != 0);
            }
            ptx::umma_commit(&empty[stage]);              // smem stage reusable once these MMAs retire
            if (ka == KATOMS - 1) ptx::umma_commit(&tfull[as]);  // accumulator complete
          }
          __syncwarp();
        }
      }
      if (lane == 0) ptx::umma_commit(aempty);
      __syncwarp();
    }
  } else {
    // =============================================================== epilogue: TMEM -> threshold-gated top-KP
    const uint32_t q = warp & 3;              // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;            // row of the CTA tile == TMEM lane
    float* my_bk = bufk + row * S::CBP;
    int* my_bi = bufi + row * S::CBP;
    uint32_t tc = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int split = item / n_qblocks, qb = item - split * n_qblocks;
      const int t0 = split * tiles_per_split;
      const int t1 = min(n_tiles, t0 + tiles_per_split);
      // candidate lists of this warp's 32 rows: [(qb*128 + q*32 + r) * n_splits + split][KP]
      const size_t list0 = (size_t(qb * BM + q * 32) * n_splits + split) * KP;
      const size_t list_stride = size_t(n_splits) * KP;
      for (int r = 0; r < 32; ++r) {
#pragma unroll
        for (int j = 0; j < KP / 32; ++j) {
          cand_key[list0 + r * list_stride + j * 32 + lane] = kInf;
          cand_idx[list0 + r * list_stride + j * 32 + lane] = -1;
        }
      }
      __syncwarp();
      float thr = (qb * BM + row < nq) ? kInf : -kInf;   // rows past the last query never accept a candidate
      int cnt = 0;

      auto flush_rows = [&](uint32_t need) {
        while (need) {
          const int r = __ffs(need) - 1;
          need &= need - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, r);
          __syncwarp();
          const float nt = flush_row(bufk + (q * 32 + r) * S::CBP, bufi + (q * 32 + r) * S::CBP, n,
                                     cand_key + list0 + r * list_stride, cand_idx + list0 + r * list_stride, lane);
          if (int(lane) == r) {
            thr = nt;
            cnt = 0;
          }
        }
      };

      for (int t = t0; t < t1; ++t, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait(&tfull[as], (tc >> 1) & 1);
        ptx::tc_fence_after();
        const float* nrm = norm_ring + (tc % S::NT) * BN;
        const uint32_t taddr = tmem_base + ((q * 32) << 16) + as * BN;
        const int col0 = t * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          ptx::tmem_ld32(taddr + c * 32, v);
          ptx::tmem_ld_wait(v);
          if (c == BN / 32 - 1) {  // accumulator fully copied to registers: hand the TMEM stage back to the MMA warp
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty[as]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 n0 = *reinterpret_cast<const float4*>(nrm + c * 32 + g * 8);
            const float4 n1 = *reinterpret_cast<const float4*>(nrm + c * 32 + g * 8 + 4);
            float key[8];
            // the query operand was pre-scaled by -2, so acc = -2 q.g and key = |g|^2 - 2 q.g
            key[0] = v[g * 8 + 0] + n0.x; key[1] = v[g * 8 + 1] + n0.y;
            key[2] = v[g * 8 + 2] + n0.z; key[3] = v[g * 8 + 3] + n0.w;
            key[4] = v[g * 8 + 4] + n1.x; key[5] = v[g * 8 + 5] + n1.y;
            key[6] = v[g * 8 + 6] + n1.z; key[7] = v[g * 8 + 7] + n1.w;
            const float m = fminf(fminf(fminf(key[0], key[1]), fminf(key[2], key[3])),
                                  fminf(fminf(key[4], key[5]), fminf(key[6], key[7])));
            if (__any_sync(0xffffffffu, m < thr)) {
              flush_rows(__ballot_sync(0xffffffffu, cnt > S::CB - 8));
              const int cbase = col0 + c * 32 + g * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (key[j] < thr) {
                  my_bk[cnt] = key[j];
                  my_bi[cnt] = cbase + j;
                  ++cnt;
                }
              }
            }
          }
        }
        ptx::mbar_arrive(&nempty[tc % S::NT]);
      }
      __syncwarp();
      flush_rows(__ballot_sync(0xffffffffu, cnt > 0));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

// ------------------------------------------------------------------------------------------------ rerank
// One warp per query: merge the per-split candidate lists, recompute the KP candidates' distances exactly,
// sort by (distance, index), emit top-k, certify.
constexpr int RR_WARPS = 4;

__global__ void __launch_bounds__(RR_WARPS * 32)
knn_rerank_kernel(const float* __restrict__ Q, const float* __restrict__ G, int nq, int64_t ng, int D,
                  const float* __restrict__ cand_key, const int* __restrict__ cand_idx, int n_splits,
                  const float* __restrict__ qnorm, const float* __restrict__ qerr, const float* __restrict__ gstats,
                  float delta_coeff, int k, int exclude_self, int64_t self_offset,
                  float* __restrict__ out_dist, int* __restrict__ out_idx,
                  int* __restrict__ status, int* __restrict__ unc_query, float* __restrict__ unc_bound, int unc_cap) {
  extern __shared__ float rr_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * RR_WARPS + warp;
  if (qi >= nq) return;
  float* qs = rr_smem + warp * (D + 2 * KP);      // query vector, then sort scratch
  float* sk = qs + D;
  int* sv = reinterpret_cast<int*>(sk + KP);

  for (int c = lane; c < D; c += 32) qs[c] = Q[size_t(qi) * D + c];

  // ---- merge split lists by approximate key
  const size_t base = size_t(qi) * n_splits * KP;
  float xk[KP / 32];
  int xv[KP / 32];
#pragma unroll
  for (int j = 0; j < KP / 32; ++j) {
    xk[j] = cand_key[base + j * 32 + lane];
    xv[j] = cand_idx[base + j * 32 + lane];
  }
  for (int s = 1; s < n_splits; ++s) {
    for (int j2 = 0; j2 < KP / 32; ++j2) {
      float yk = cand_key[base + size_t(s) * KP + j2 * 32 + lane];
      int yv = cand_idx[base + size_t(s) * KP + j2 * 32 + lane];
#pragma unroll
      for (int j = 0; j < KP / 32; ++j) warp_merge_split32(xk[j], xv[j], yk, yv, lane);
    }
  }
  const float tau = __shfl_sync(0xffffffffu, xk[KP / 32 - 1], 31);
  __syncwarp();

  // ---- exact distances (reference arithmetic), one candidate per lane at a time
  const int self = exclude_self ? int(self_offset + qi) : -1;
#pragma unroll
  for (int j = 0; j < KP / 32; ++j) {
    const int idx = xv[j];
    float d = kInf;
    if (idx >= 0 && idx != self) d = exact_l2(qs, G + size_t(idx) * D, D);
    sk[j * 32 + lane] = d;
    sv[j * 32 + lane] = (idx >= 0 && idx != self) ? idx : 0x7fffffff;
  }
  __syncwarp();

  // ---- bitonic sort of KP pairs in shared memory by (distance, index)
  for (int size = 2; size <= KP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < KP / 2; t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const float a = sk[lo], b = sk[hi];
        const int ia = sv[lo], ib = sv[hi];
        const bool gt = (a > b) || (a == b && ia > ib);
        if (gt == asc) {
          sk[lo] = b; sk[hi] = a;
          sv[lo] = ib; sv[hi] = ia;
        }
      }
      __syncwarp();
    }
  }

  for (int r = lane; r < k; r += 32) {
    const float d = sk[r];
    out_dist[size_t(qi) * k + r] = d;
    out_idx[size_t(qi) * k + r] = (d < kInf) ? sv[r] : -1;
  }

  // ---- certificate
  if (lane == 0) {
    const float dk = sk[k - 1];
    bool ok;
    if (!(tau < kInf)) {
      // the list never filled: every gallery row with a finite key is a candidate
      ok = (gstats[0] < kInf) && (gstats[1] < kInf) && (qerr[qi] < kInf);
    } else {
      const float qn = qnorm[qi];
      const float delta = delta_coeff * (qn + gstats[1]);
      const float lb2 = tau + qn - delta;
      const float lb = sqrtf(fmaxf(lb2, 0.f)) * 0.999999f - (qerr[qi] + gstats[0]);
      ok = dk < lb;
    }
    if (!ok) {
      const int slot = atomicAdd(&status[0], 1);
      if (slot < unc_cap) {
        unc_query[slot] = qi;
        unc_bound[slot] = dk;
      } else {
        status[2] = 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ fallback
// Exact recomputation for the (rare) uncertified queries.  Pass 1 collects every gallery row whose exact distance
// is <= the query's upper bound (the k-th exact candidate distance); pass 2 sorts the collected rows.
constexpr int FB_CAP = 2048;  // collected rows per uncertified query

__global__ void knn_fallback_collect_kernel(const float* __restrict__ Q, const float* __restrict__ G, int64_t ng, int D,
                                            int exclude_self, int64_t self_offset, const int* __restrict__ status,
                                            const int* __restrict__ unc_query, const float* __restrict__ unc_bound,
                                            int unc_cap, int* __restrict__ fb_count, float* __restrict__ fb_dist,
                                            int* __restrict__ fb_idx, int* __restrict__ status_w) {
  extern __shared__ float fb_q[];
  const int n_unc = min(status[0], unc_cap);
  const int chunks = gridDim.x;
  for (int slot = blockIdx.y; slot < n_unc; slot += gridDim.y) {
    const int qi = unc_query[slot];
    const float bound = unc_bound[slot];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) fb_q[c] = Q[size_t(qi) * D + c];
    __syncthreads();
    const int self = exclude_self ? int(self_offset + qi) : -1;
    const int64_t per = (ng + chunks - 1) / chunks;
    const int64_t g0 = int64_t(blockIdx.x) * per, g1 = min(ng, g0 + per);
    for (int64_t g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
      if (int(g) == self) continue;
      const float d = exact_l2(fb_q, G + g * D, D);
      if (d <= bound) {
        const int pos = atomicAdd(&fb_count[slot], 1);
        if (pos < FB_CAP) {
          fb_dist[size_t(slot) * FB_CAP + pos] = d;
          fb_idx[size_t(slot) * FB_CAP + pos] = int(g);
        } else {
          status_w[1] = 1;  // overflow: more than FB_CAP rows within the bound (massive ties)
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
knn_fallback_select_kernel(const int* __restrict__ status, const int* __restrict__ unc_query, int unc_cap,
                           const int* __restrict__ fb_count, const float* __restrict__ fb_dist,
                           const int* __restrict__ fb_idx, int k, float* __restrict__ out_dist, int* __restrict__ out_idx) {
  __shared__ float sk[FB_CAP];
  __shared__ int sv[FB_CAP];
  const int n_unc = min(status[0], unc_cap);
  for (int slot = blockIdx.x; slot < n_unc; slot += gridDim.x) {
    const int n = min(fb_count[slot], FB_CAP);
    __syncthreads();
    for (int i = threadIdx.x; i < FB_CAP; i += blockDim.x) {
      sk[i] = i < n ? fb_dist[size_t(slot) * FB_CAP + i] : kInf;
      sv[i] = i < n ? fb_idx[size_t(slot) * FB_CAP + i] : 0x7fffffff;
    }
    __syncthreads();
    for (int size = 2; size <= FB_CAP; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < FB_CAP / 2; t += blockDim.x) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool asc = (lo & size) == 0;
          const float a = sk[lo], b = sk[hi];
          const int ia = sv[lo], ib = sv[hi];
          const bool gt = (a > b) || (a == b && ia > ib);
          if (gt == asc) {
            sk[lo] = b; sk[hi] = a;
            sv[lo] = ib; sv[hi] = ia;
          }
        }
        __syncthreads();
      }
    }
    const int qi = unc_query[slot];
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
      out_dist[size_t(qi) * k + r] = sk[r];
      out_idx[size_t(qi) * k + r] = sk[r] < kInf ? sv[r] : -1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// fp16 row-major [rows, Dp] matrix, boxes of 64 columns (128 bytes, SWIZZLE_128B) x box_rows rows; OOB rows read 0.
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int Dp, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  MMSIM_REQUIRE(fn != nullptr, MMSIM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {cuuint64_t(Dp), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(Dp) * 2};
  cuuint32_t box[2] = {cuuint32_t(KATOM), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMSIM_REQUIRE(r == CUDA_SUCCESS, MMSIM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return MMSIM_OK;
}

Plan make_plan(int64_t nq, int64_t ng, int64_t D, int k, int num_sms) {
  Plan p{};
  p.Dp = int(align_up(size_t(D), KATOM));
  p.katoms = p.Dp / KATOM;
  p.n_qblocks = int((nq + BM - 1) / BM);
  p.n_tiles = int((ng + BN - 1) / BN);
  // gallery splits: fill the persistent grid in whole waves without making sweeps too short
  int best_s = 1;
  double best_eff = 0;
  for (int s = 1; s <= 8; ++s) {
    const int tps = (p.n_tiles + s - 1) / s;
    if (s > 1 && tps < 64) break;
    const int s_eff = (p.n_tiles + tps - 1) / tps;
    const int64_t items = int64_t(p.n_qblocks) * s_eff;
    const int64_t waves = (items + num_sms - 1) / num_sms;
    // cost model: waves * tiles per sweep (+ a small per-split rerank/merge overhead)
    const double eff = double(p.n_qblocks) * p.n_tiles / (double(waves) * num_sms * tps) - 0.01 * (s - 1);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_s = s_eff;
    }
  }
  p.n_splits = best_s;
  p.tiles_per_split = (p.n_tiles + p.n_splits - 1) / p.n_splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.grid = int(std::min<int64_t>(num_sms, int64_t(p.n_qblocks) * p.n_splits));
  p.unc_cap = int(std::min<int64_t>(nq, 1024));

  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  const size_t q_rows = size_t(p.n_qblocks) * BM;
  p.off_qh = take(q_rows * p.Dp * 2);
  p.off_gh = take(size_t(ng) * p.Dp * 2);
  p.off_gnorm = take(size_t(p.n_tiles) * BN * 4);
  p.off_qnorm = take(q_rows * 4);
  p.off_qerr = take(q_rows * 4);
  p.off_stats = take(64);
  p.off_cand_key = take(q_rows * p.n_splits * KP * 4);
  p.off_cand_idx = take(q_rows * p.n_splits * KP * 4);
  p.off_unc_query = take(size_t(p.unc_cap) * 4);
  p.off_unc_bound = take(size_t(p.unc_cap) * 4);
  p.off_fb_count = take(size_t(p.unc_cap) * 4);
  p.off_fb_dist = take(size_t(p.unc_cap) * FB_CAP * 4);
  p.off_fb_idx = take(size_t(p.unc_cap) * FB_CAP * 4);
  p.total_bytes = off;
  return p;
}

template <int KATOMS>
static int launch_tc(const Plan& p, const CUtensorMap& tq, const CUtensorMap& tg, const float* gnorm, int nq,
                     float* cand_key, int* cand_idx, cudaStream_t stream) {
  using S = Smem<KATOMS>;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(knn_tc_kernel<KATOMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES));
  knn_tc_kernel<KATOMS><<<p.grid, NUM_THREADS, S::DYN_BYTES, stream>>>(tq, tg, gnorm, nq, p.n_qblocks, p.n_splits,
                                                                       p.tiles_per_split, p.n_tiles, cand_key, cand_idx);
  MMSIM_CUDA_CHECK(cudaGetLastError());
  return MMSIM_OK;
}

int run(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self, int64_t self_offset,
        float* out_dist, int* out_idx, int* status, void* ws, size_t ws_bytes, cudaStream_t stream, int phases) {
  MMSIM_REQUIRE(Q && G && out_dist && out_idx && status && ws, MMSIM_ERR_ARG, "knn: null pointer argument");
  MMSIM_REQUIRE(nq > 0 && ng > 0 && D > 0, MMSIM_ERR_ARG, "knn: empty input (nq=%lld ng=%lld D=%lld)", (long long)nq,
                (long long)ng, (long long)D);
  MMSIM_REQUIRE(D <= 4 * KATOM, MMSIM_ERR_UNSUPPORTED, "knn: D=%lld > 256 is not supported by the tcgen05 path", (long long)D);
  MMSIM_REQUIRE(k >= 1 && k + (exclude_self ? 1 : 0) <= KP - 16, MMSIM_ERR_UNSUPPORTED,
                "knn: k=%d unsupported (1 <= k <= %d)", k, KP - 16 - (exclude_self ? 1 : 0));
  MMSIM_REQUIRE(ng < (int64_t(1) << 31) - BN && nq < (int64_t(1) << 31) - BM, MMSIM_ERR_ARG, "knn: a shard must hold < 2^31 rows");
  MMSIM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MMSIM_ERR_WORKSPACE, "knn: workspace must be 256-byte aligned");

  int dev = 0, num_sms = 0;
  MMSIM_CUDA_CHECK(cudaGetDevice(&dev));
  MMSIM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const Plan p = make_plan(nq, ng, D, k, num_sms);
  MMSIM_REQUIRE(ws_bytes >= p.total_bytes, MMSIM_ERR_WORKSPACE, "knn: workspace too small (%zu < %zu)", ws_bytes, p.total_bytes);

  uint8_t* w = static_cast<uint8_t*>(ws);
  __half* qh = reinterpret_cast<__half*>(w + p.off_qh);
  __half* gh = reinterpret_cast<__half*>(w + p.off_gh);
  float* gnorm = reinterpret_cast<float*>(w + p.off_gnorm);
  float* qnorm = reinterpret_cast<float*>(w + p.off_qnorm);
  float* qerr = reinterpret_cast<float*>(w + p.off_qerr);
  float* gstats = reinterpret_cast<float*>(w + p.off_stats);
  float* cand_key = reinterpret_cast<float*>(w + p.off_cand_key);
  int* cand_idx = reinterpret_cast<int*>(w + p.off_cand_idx);
  int* unc_query = reinterpret_cast<int*>(w + p.off_unc_query);
  float* unc_bound = reinterpret_cast<float*>(w + p.off_unc_bound);
  int* fb_count = reinterpret_cast<int*>(w + p.off_fb_count);
  float* fb_dist = reinterpret_cast<float*>(w + p.off_fb_dist);
  int* fb_idx = reinterpret_cast<int*>(w + p.off_fb_idx);

  if (phases & kPhaseRerank) {
    MMSIM_CUDA_CHECK(cudaMemsetAsync(status, 0, 8 * sizeof(int), stream));
    MMSIM_CUDA_CHECK(cudaMemsetAsync(fb_count, 0, size_t(p.unc_cap) * 4, stream));
  }

  // 1. operand copies
  if (phases & kPhasePrep) {
    MMSIM_CUDA_CHECK(cudaMemsetAsync(gstats, 0, 64, stream));
    const int threads = 256;
    const int64_t g_pad = int64_t(p.n_tiles) * BN;
    const int64_t gb = (g_pad * 32 + threads - 1) / threads;
    prep_rows_kernel<<<unsigned(gb), threads, 0, stream>>>(G, ng, g_pad, int(D), p.Dp, 1.0f, gh, gnorm, nullptr,
                                                           reinterpret_cast<unsigned int*>(gstats));
    MMSIM_CUDA_CHECK(cudaGetLastError());
    const int64_t qb = (nq * 32 + threads - 1) / threads;
    prep_rows_kernel<<<unsigned(qb), threads, 0, stream>>>(Q, nq, nq, int(D), p.Dp, -2.0f, qh, qnorm, qerr, nullptr);
    MMSIM_CUDA_CHECK(cudaGetLastError());
  }

  // 2. fused distance + candidate selection
  int rc = MMSIM_OK;
  if (phases & kPhaseTensor) {
  CUtensorMap tq, tg;
  rc = make_tmap(&tq, qh, nq, p.Dp, BM);
  if (rc) return rc;
  rc = make_tmap(&tg, gh, ng, p.Dp, BN);
  if (rc) return rc;
  switch (p.katoms) {
    case 1: rc = launch_tc<1>(p, tq, tg, gnorm, int(nq), cand_key, cand_idx, stream); break;
    case 2: rc = launch_tc<2>(p, tq, tg, gnorm, int(nq), cand_key, cand_idx, stream); break;
    case 3: rc = launch_tc<3>(p, tq, tg, gnorm, int(nq), cand_key, cand_idx, stream); break;
    default: rc = launch_tc<4>(p, tq, tg, gnorm, int(nq), cand_key, cand_idx, stream); break;
  }
  if (rc) return rc;
  }

  // 3. exact re-rank + certificate.  delta bounds the fp32 accumulation error of key = |g|^2 - 2 q.g:
  //    (Dp + 8) roundings of relative size 2^-24, on terms bounded by (|q|^2 + |g|^2), with a 4x safety factor.
  if (phases & kPhaseRerank) {
    const float delta_coeff = 4.0f * float(p.Dp + 8) * 5.9604645e-8f;
    const int blocks = int((nq + RR_WARPS - 1) / RR_WARPS);
    const size_t smem = size_t(RR_WARPS) * (size_t(D) + 2 * KP) * 4;
    knn_rerank_kernel<<<blocks, RR_WARPS * 32, smem, stream>>>(Q, G, int(nq), ng, int(D), cand_key, cand_idx, p.n_splits,
                                                               qnorm, qerr, gstats, delta_coeff, k, exclude_self,
                                                               self_offset, out_dist, out_idx, status, unc_query,
                                                               unc_bound, p.unc_cap);
    MMSIM_CUDA_CHECK(cudaGetLastError());
  }

  // 4. exact fallback for uncertified queries (no-op when status[0] == 0; the count lives on the device)
  if (phases & kPhaseFallback) {
    dim3 grid(64, 16);
    knn_fallback_collect_kernel<<<grid, 256, size_t(D) * 4, stream>>>(Q, G, ng, int(D), exclude_self, self_offset, status,
                                                                      unc_query, unc_bound, p.unc_cap, fb_count, fb_dist,
                                                                      fb_idx, status);
    MMSIM_CUDA_CHECK(cudaGetLastError());
    knn_fallback_select_kernel<<<32, 256, 0, stream>>>(status, unc_query, p.unc_cap, fb_count, fb_dist, fb_idx, k, out_dist,
                                                       out_idx);
    MMSIM_CUDA_CHECK(cudaGetLastError());
  }
  return MMSIM_OK;
}

}  // namespace knn
}  // namespace mmsim
