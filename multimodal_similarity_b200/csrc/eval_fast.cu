// K5c: leave-one-out retrieval evaluation without a per-query re-read of the embeddings and without a comparison sort --
// the loop body of utils.evaluate / utils.evaluate_simple (src/utils.py:83-229) in three kernels:
//   1. eval_tile_dist_kernel     exact fp32 distances (NumPy summation order, exact.cuh) of a batch of query rows to ALL rows,
//                                register-tiled: a CTA keeps 32 query rows in shared memory and streams 32-row gallery tiles
//                                past them; 8 adjacent lanes own NumPy's 8 strided accumulators of a pair, a thread holds an
//                                8 x 4 block of pairs (12 shared-memory loads per 96 arithmetic instructions).  Output: the
//                                float bits [B][ldk] in the ORIGINAL row numbering, the query's own row as 0xffffffff.
//   2. eval_sort_metrics_kernel  one CTA per query: its N keys go through a stable 4-pass LSD radix sort (8-bit digits) in
//                                shared memory -- warp-private digit counters, `match.any` ranks inside a warp step, the
//                                scatter runs in place through registers -- which orders rows by (distance, index) exactly
//                                like the reference's argsort does outside ties; passes whose digit is already in order
//                                (the exponent byte of distances in [0.5, 2), say) are skipped.  Then the streaming metrics
//                                pass (eval_metrics.cuh) reads the ranking from shared memory.          N <= 24,576
//   3. eval_confusion_kernel     the reference's confusion-matrix accumulation `cm[row] += frac.astype(float32)` in query
//                                order (src/utils.py:214-220), one block per class row, float32 adds in the same sequence.
// Non-negative float32 bit patterns order like the floats; NaN distances sort after +inf, the query's own row last.
#include <cuda_runtime.h>

#include "common.cuh"
#include "eval.h"
#include "eval_metrics.cuh"
#include "exact.cuh"

namespace mmsim {
namespace eval {

// ------------------------------------------------------------------------------------------------ 1. tiled distances
constexpr int DT_TQ = 32, DT_TR = 32, DT_THREADS = 256;

bool leaf_plan(int D, LeafPlan* lp) {
  if (D < 8) return false;
  if (D <= 128) {
    lp->n = 1; lp->lo[0] = 0; lp->len[0] = D; lp->lo[1] = 0; lp->len[1] = 0;
    return true;
  }
  int n2 = D / 2;
  n2 -= n2 % 8;
  if (n2 > 128 || D - n2 > 128) return false;
  lp->n = 2; lp->lo[0] = 0; lp->len[0] = n2; lp->lo[1] = n2; lp->len[1] = D - n2;
  return true;
}

__device__ __forceinline__ float sq_term(float a, float b) {
  const float d = __fsub_rn(a, b);
  return __fmul_rn(d, d);
}

template <int NLEAF>
__global__ void __launch_bounds__(DT_THREADS)
eval_tile_dist_kernel(const float* __restrict__ E, int N, int D, const int* __restrict__ queries, int nqb, LeafPlan lp,
                      int tiles_per_cta, uint32_t* __restrict__ keys, int64_t ldk, int* __restrict__ vals) {
  extern __shared__ float dsm[];
  __shared__ int qrow[DT_TQ];
  const int P = (D + 7) / 8 * 8 + 2;            // 4 P = 8 (mod 32): the four row groups of a warp hit disjoint banks
  float* sa = dsm;
  float* sb = dsm + DT_TQ * P;
  const int t = threadIdx.x, lane = t & 31, sub = lane & 7;
  const int grp = (t >> 5) * 4 + (lane >> 3);   // 32 groups of 8 lanes: 4 (queries) x 8 (rows)
  const int gq = grp >> 3, gr = grp & 7;
  const int q0 = blockIdx.y * DT_TQ;
  const bool vec4 = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(E) & 15) == 0;

  if (t < DT_TQ) qrow[t] = q0 + t < nqb ? queries[q0 + t] : -1;
  __syncthreads();
  // stage 32 rows into shared memory (128-bit global loads when the rows allow it; the +2 pitch keeps the stores scalar)
  auto stage = [&](float* dst, auto row_of) {
    if (vec4) {
      const int D4 = D >> 2;
      for (int x = t; x < 32 * D4; x += DT_THREADS) {
        const int r = x / D4, c = (x - r * D4) * 4;
        const int64_t row = row_of(r);
        const float4 v = row >= 0 ? *reinterpret_cast<const float4*>(E + size_t(row) * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        float* d = dst + r * P + c;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      for (int x = t; x < 32 * D; x += DT_THREADS) {
        const int r = x / D, c = x - r * D;
        const int64_t row = row_of(r);
        dst[r * P + c] = row >= 0 ? E[size_t(row) * D + c] : 0.f;
      }
    }
  };
  stage(sa, [&](int r) { return int64_t(qrow[r]); });
  const float* a0 = sa + (gq * 8) * P + sub;
  const float* b0 = sb + (gr * 4) * P + sub;

  const int tile0 = blockIdx.x * tiles_per_cta;
  for (int tile = tile0; tile < tile0 + tiles_per_cta; ++tile) {
    const int j0 = tile * DT_TR;
    if (j0 >= N) break;
    __syncthreads();                              // the previous tile's rows are consumed (and sa is staged)
    stage(sb, [&](int r) { return j0 + r < N ? int64_t(j0 + r) : int64_t(-1); });
    __syncthreads();

    float res[8][4];
#pragma unroll
    for (int leaf = 0; leaf < NLEAF; ++leaf) {
      const int lo = lp.lo[leaf], len = lp.len[leaf];
      const float* a = a0 + lo;
      const float* b = b0 + lo;
      float acc[8][4];
      {
        float av[8], bv[4];
#pragma unroll
        for (int q = 0; q < 8; ++q) av[q] = a[q * P];
#pragma unroll
        for (int r = 0; r < 4; ++r) bv[r] = b[r * P];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[q][r] = sq_term(av[q], bv[r]);
      }
      int c = 8;
      for (; c + 8 <= len; c += 8) {
        float av[8], bv[4];
#pragma unroll
        for (int q = 0; q < 8; ++q) av[q] = a[q * P + c];
#pragma unroll
        for (int r = 0; r < 4; ++r) bv[r] = b[r * P + c];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[q][r] = __fadd_rn(acc[q][r], sq_term(av[q], bv[r]));
      }
      // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)): lane `sub` holds r[sub]
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float x = acc[q][r];
          x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, 1));
          x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, 2));
          x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, 4));
          acc[q][r] = x;
        }
      // the len % 8 tail, sequentially (every lane of the group computes the same value)
      for (; c < len; ++c) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[q][r] = __fadd_rn(acc[q][r], sq_term(a[q * P + c - sub], b[r * P + c - sub]));
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) res[q][r] = leaf == 0 ? acc[q][r] : __fadd_rn(res[q][r], acc[q][r]);
    }

    // lane `sub` of a group stores the group's query `sub`
    const int ql = gq * 8 + sub;
    const int qi = qrow[ql];
    if (qi >= 0) {
      uint32_t* out = keys + size_t(q0 + ql) * ldk;
      int* vout = vals ? vals + size_t(q0 + ql) * ldk : nullptr;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q == sub) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int j = j0 + gr * 4 + r;
            if (j < N) {
              out[j] = j == qi ? 0xffffffffu : __float_as_uint(__fsqrt_rn(res[q][r]));
              if (vout) vout[j] = j;
            }
          }
        }
      }
    }
  }
}

// any D (feature widths whose NumPy summation tree is deeper than two leaves, or D < 8): one thread per pair
__global__ void __launch_bounds__(256)
eval_dist_generic_kernel(const float* __restrict__ E, int N, int D, const int* __restrict__ queries, uint32_t* __restrict__ keys,
                         int64_t ldk, int* __restrict__ vals) {
  extern __shared__ float dq[];
  const int i = queries[blockIdx.y];
  for (int c = threadIdx.x; c < D; c += blockDim.x) dq[c] = E[size_t(i) * D + c];
  __syncthreads();
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    keys[size_t(blockIdx.y) * ldk + j] = j == i ? 0xffffffffu : __float_as_uint(exact_l2(dq, E + size_t(j) * D, D));
    if (vals) vals[size_t(blockIdx.y) * ldk + j] = j;
  }
}

int launch_distances(const float* E, int64_t N, int64_t D, const int* queries, int nqb, uint32_t* keys, int64_t ldk, int* vals,
                     cudaStream_t s) {
  LeafPlan lp;
  if (leaf_plan(int(D), &lp)) {
    const int P = (int(D) + 7) / 8 * 8 + 2;
    const size_t smem = size_t(DT_TQ + DT_TR) * P * 4;
    const int row_tiles = int((N + DT_TR - 1) / DT_TR);
    const int qtiles = (nqb + DT_TQ - 1) / DT_TQ;
    int tiles_per_cta = 16;                                   // amortises staging the query rows; keep >= ~4 CTAs per SM in flight
    while (tiles_per_cta > 1 && int64_t((row_tiles + tiles_per_cta - 1) / tiles_per_cta) * qtiles < 148 * 8) tiles_per_cta >>= 1;
    const dim3 grid(unsigned((row_tiles + tiles_per_cta - 1) / tiles_per_cta), unsigned(qtiles));
    if (lp.n == 1) {
      MMSIM_CUDA_CHECK(cudaFuncSetAttribute(eval_tile_dist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      eval_tile_dist_kernel<1><<<grid, DT_THREADS, smem, s>>>(E, int(N), int(D), queries, nqb, lp, tiles_per_cta, keys, ldk, vals);
    } else {
      MMSIM_CUDA_CHECK(cudaFuncSetAttribute(eval_tile_dist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      eval_tile_dist_kernel<2><<<grid, DT_THREADS, smem, s>>>(E, int(N), int(D), queries, nqb, lp, tiles_per_cta, keys, ldk, vals);
    }
  } else {
    const dim3 grid(unsigned(std::min<int64_t>((N + 255) / 256, 4096)), unsigned(nqb));
    eval_dist_generic_kernel<<<grid, 256, size_t(D) * 4, s>>>(E, int(N), int(D), queries, keys, ldk, vals);
  }
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

// ------------------------------------------------------------------------- 2. shared-memory radix sort + metrics
template <int NT, int IPT>
__global__ void __launch_bounds__(NT, NT <= 256 ? 4 : (NT <= 512 ? 2 : 1))
eval_sort_metrics_kernel(const uint32_t* __restrict__ keys, int64_t ldk, const int* __restrict__ labels, const int* __restrict__ cls,
                         int N, int C, const int* __restrict__ queries, double alpha, int aligned, double* __restrict__ out_ap,
                         int* __restrict__ out_npos, int* __restrict__ out_first, int* __restrict__ out_depth,
                         int* __restrict__ out_hist, int* __restrict__ out_rank) {
  constexpr int NW = NT / 32, CAP = NT * IPT, CP = 257;
  static_assert(IPT % 2 == 0 && CAP <= 65536 && NW % 8 == 0, "two 16-bit row indices per register");
  extern __shared__ __align__(16) unsigned char esm[];
  uint32_t* sk = reinterpret_cast<uint32_t*>(esm);                  // [CAP] key bits
  uint16_t* si = reinterpret_cast<uint16_t*>(sk + CAP);             // [CAP] original row
  int* cnt = reinterpret_cast<int*>(si + CAP);                      // [NW][CP] digit counters, one row per warp (odd pitch:
  int* lhist = cnt + CP * NW;                                       //   lanes with different digits hit different banks); [C]
  __shared__ int wtot[32];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int qn = blockIdx.x;
  const uint32_t* krow = keys + size_t(qn) * ldk;
  const int base = warp * (IPT * 32);
  const unsigned lt_mask = (1u << lane) - 1u;

  uint32_t key[IPT];
  uint32_t ip[IPT / 2];

  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    // ---- A: load my items (warp `warp` owns the contiguous items [base, base + 32 IPT)), count digits per warp
    for (int x = t; x < CP * NW; x += NT) cnt[x] = 0;
    int unsorted = pass == 0;
    uint32_t prev_digit = 0;
    if (pass > 0) prev_digit = base > 0 ? (sk[base - 1] >> shift) & 255u : 0u;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
      const int item = base + it * 32 + lane;
      uint32_t k, id;
      if (pass == 0) {
        k = item < N ? krow[item] : 0xffffffffu;
        id = uint32_t(item) & 0xffffu;
      } else {
        k = sk[item];
        id = si[item];
      }
      key[it] = k;
      if (it & 1) ip[it / 2] |= id << 16; else ip[it / 2] = id;
      const uint32_t digit = (k >> shift) & 255u;
      uint32_t before = __shfl_up_sync(0xffffffffu, digit, 1);
      if (lane == 0) before = prev_digit;
      unsorted |= digit < before;
      prev_digit = __shfl_sync(0xffffffffu, digit, 31);
      atomicAdd(&cnt[warp * CP + digit], 1);      // warp-private row: counting needs no order (the scatter below does)
    }
    if (!__syncthreads_or(unsorted)) continue;      // this digit is already in order everywhere: the pass is the identity

    // ---- B: exclusive scan of the counters in (digit, warp) order
    {
      // thread t owns entries 8 t .. 8 t + 7 of the (digit, warp) order: digit (8 t) / NW, warps (8 t) % NW ..
      int* c8 = cnt + ((8 * t) % NW) * CP + (8 * t) / NW;
      int v[8], s = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = c8[e * CP];
#pragma unroll
      for (int e = 0; e < 8; ++e) { const int x = v[e]; v[e] = s; s += x; }
      int incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) wtot[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        int w = lane < NW ? wtot[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += y;
        }
        wtot[lane] = w;
      }
      __syncthreads();
      const int off = incl - s + (warp ? wtot[warp - 1] : 0);
#pragma unroll
      for (int e = 0; e < 8; ++e) c8[e * CP] = v[e] + off;
    }
    __syncthreads();

    // ---- C: stable scatter, in place (every item is in a register by now)
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
      const uint32_t k = key[it];
      const uint32_t id = (it & 1) ? ip[it / 2] >> 16 : ip[it / 2] & 0xffffu;
      const uint32_t digit = (k >> shift) & 255u;
      const unsigned same = __match_any_sync(0xffffffffu, digit);
      const int b = cnt[warp * CP + digit];
      __syncwarp();
      if (lane == __ffs(same) - 1) cnt[warp * CP + digit] = b + __popc(same);
      __syncwarp();
      const int pos = b + __popc(same & lt_mask);
      sk[pos] = k;
      si[pos] = uint16_t(id);
    }
    __syncthreads();
  }
  __syncthreads();
  loo_metrics<NT, uint16_t>(sk, si, labels, cls, N, C, queries[qn], qn, alpha, aligned, lhist, out_ap, out_npos, out_first,
                            out_depth, out_hist, out_rank);
}

static size_t sort_smem_bytes(int nt, int ipt, int C) { return size_t(nt) * ipt * 6 + size_t(257) * (nt / 32) * 4 + size_t(C) * 4; }

template <int NT, int IPT>
static int launch_sort(const uint32_t* keys, int64_t ldk, const int* labels, const int* cls, int N, int C, const int* queries, int nqb,
                       double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank,
                       cudaStream_t s) {
  const size_t smem = sort_smem_bytes(NT, IPT, C);
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(eval_sort_metrics_kernel<NT, IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  eval_sort_metrics_kernel<NT, IPT><<<nqb, NT, smem, s>>>(keys, ldk, labels, cls, N, C, queries, alpha, aligned, ap, npos, first,
                                                          depth, hist, rank);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

bool sort_path_fits(int64_t N, int C) {
  if (N > kSortMaxN) return false;
  const int ipt = int((N + 1023) / 1024);
  return sort_smem_bytes(1024, ipt + (ipt & 1), C) <= 220 * 1024;
}

int launch_sort_metrics(const uint32_t* keys, int64_t ldk, const int* labels, const int* cls, int64_t N, int C, const int* queries,
                        int nqb, double alpha, int aligned, double* ap, int* npos, int* first, int* depth, int* hist, int* rank,
                        cudaStream_t s) {
#define MMSIM_SORT_CASE(NT, IPT) \
  if (N <= int64_t(NT) * IPT)    \
    return launch_sort<NT, IPT>(keys, ldk, labels, cls, int(N), C, queries, nqb, alpha, aligned, ap, npos, first, depth, hist, rank, s);
  MMSIM_SORT_CASE(256, 2)        // small rankings: small CTAs, several per SM
  MMSIM_SORT_CASE(256, 8)
  MMSIM_SORT_CASE(512, 8)
  MMSIM_SORT_CASE(512, 12)
  MMSIM_SORT_CASE(512, 16)
  MMSIM_SORT_CASE(1024, 12)
  MMSIM_SORT_CASE(1024, 16)
  MMSIM_SORT_CASE(1024, 20)
  MMSIM_SORT_CASE(1024, 24)
#undef MMSIM_SORT_CASE
  set_error("evaluate: N=%lld exceeds the shared-memory sort (N <= %d)", (long long)N, kSortMaxN);
  return MMSIM_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------ 3. confusion accumulation
__global__ void __launch_bounds__(256)
eval_confusion_kernel(const int* __restrict__ hist, const int* __restrict__ depth, const int* __restrict__ npos,
                      const int* __restrict__ qcls, int nq, int C, float* __restrict__ cm, int* __restrict__ count,
                      int* __restrict__ lists /* [nq] scratch: the queries of each class, in query order */,
                      const int* __restrict__ list_off /* [C] start of each class's list (all its queries would fit) */) {
  __shared__ int wcnt[8];
  __shared__ int s_total;
  const int row = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int* mine = lists + list_off[row];
  // the queries of this class with at least one positive (src/utils.py:175-180 skips the others), in query order
  int total = 0;
  for (int base = 0; base < nq; base += 256) {
    const int n = base + t;
    const bool ok = n < nq && qcls[n] == row && npos[n] > 0;
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) wcnt[warp] = __popc(b);
    __syncthreads();
    int off = total;
    for (int w = 0; w < warp; ++w) off += wcnt[w];
    if (ok) mine[off + __popc(b & ((1u << lane) - 1u))] = n;
    for (int w = 0; w < 8; ++w) total += wcnt[w];
    __syncthreads();
  }
  if (t == 0) { count[row] = total; s_total = total; }
  __syncthreads();                                   // (also orders the list writes before the reads below: same block)
  for (int j = t; j < C; j += 256) {
    float acc = 0.f;
    for (int e = 0; e < total; ++e) {
      const int n = mine[e];
      acc = __fadd_rn(acc, float(double(hist[size_t(n) * C + j]) / double(depth[n])));   // int / int -> float64 -> float32 (:252,218)
    }
    cm[size_t(row) * C + j] = acc;
  }
}

int confusion(const int* hist, const int* depth, const int* npos, const int* qcls, int64_t nq, int C, float* cm, int* count,
              int* lists, const int* list_off, cudaStream_t s) {
  MMSIM_REQUIRE(hist && depth && npos && qcls && cm && count && lists && list_off, MMSIM_ERR_ARG, "evaluate_confusion: null pointer argument");
  MMSIM_REQUIRE(nq >= 0 && nq < (int64_t(1) << 31) && C >= 1, MMSIM_ERR_ARG, "evaluate_confusion: bad sizes nq=%lld C=%d", (long long)nq, C);
  eval_confusion_kernel<<<C, 256, 0, s>>>(hist, depth, npos, qcls, int(nq), C, cm, count, lists, list_off);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace eval
}  // namespace mmsim
