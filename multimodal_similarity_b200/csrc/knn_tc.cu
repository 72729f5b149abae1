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
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "exact.cuh"
#include "knn.h"
#include "knn_fallback.cuh"
#include "ptx.cuh"

// the ablation instantiations of the sweep (ABL != 0) return early from the scan on purpose
#pragma nv_diag_suppress 128

namespace mmsim {
namespace knn {

// ------------------------------------------------------------------------------------------------ prep
constexpr int BM = 128;                    // queries per CTA tile (UMMA M, one TMEM lane each)
constexpr int BN = 256;                    // gallery rows per tile (UMMA N)
constexpr int KATOM = 64;                  // fp16 elements per 128-byte swizzle atom
constexpr int NPACK = 320;                 // floats per gallery tile in the norm pack: 256 norms | 32 min8 | 8 min32 | pad

// One warp per row, grid-stride.  xh[row, 0..Dp) = fp16(x) * scale (zero padded), norm = sum fp16(x)^2 (unscaled, fp32),
// err[row] = ||x - fp16(x)||_2 (row = OUTPUT row).  With tile_pack != 0 (gallery) the norm goes to the per-tile pack
// norm[(row / 256) * 320 + row % 256] and rows in [n, n_pad) get +inf (masks padded gallery columns).
// HBM-bound: 4 D bytes read + 2 Dp bytes written per row; a lane moves 16 bytes in / 8 bytes out per step when
// D % 4 == 0 (VEC), and the running maxima reach global memory as ONE atomic pair per block.
constexpr int PREP_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(PREP_THREADS)
prep_rows_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int D, int Dp, float scale,
                 __half* __restrict__ xh, float* __restrict__ norm, float* __restrict__ err, int tile_pack,
                 unsigned int* __restrict__ max_stats /* [0]=max err bits, [1]=max norm bits, or null */,
                 const int* __restrict__ gather /* output row r is input row gather[r], or null: identity */) {
  const uint32_t lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float max_e = 0.f, max_s = 0.f;
  for (int64_t row = warp0; row < n_pad; row += nwarps) {
    const int64_t nslot = tile_pack ? (row / BN) * NPACK + (row % BN) : row;
    if (row >= n) {
      if (lane == 0) norm[nslot] = kInf;
      continue;
    }
    const float* xr = x + (gather ? int64_t(gather[row]) : row) * D;
    __half* hr = xh + row * Dp;
    float s = 0.f, e = 0.f;
    auto one = [&](float v) {
      const __half h = __float2half_rn(v);
      const float r = __half2float(h);
      s = fmaf(r, r, s);
      const float d = v - r;
      e = fmaf(d, d, e);
      const __half o = __float2half_rn(r * scale);  // scale is a power of two: exact unless it overflows (then err = inf)
      if (isinf(r * scale) || isinf(r)) e = kInf;
      return o;
    };
    if (VEC) {
      for (int c = lane * 4; c < Dp; c += 128) {
        const float4 v = c < D ? __ldcs(reinterpret_cast<const float4*>(xr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const __half2 lo = __halves2half2(one(v.x), one(v.y)), hi = __halves2half2(one(v.z), one(v.w));
        uint2 w;
        w.x = *reinterpret_cast<const uint32_t*>(&lo);
        w.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(hr + c) = w;
      }
    } else {
      for (int c = lane; c < Dp; c += 32) hr[c] = one(c < D ? xr[c] : 0.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    // inflate by a few ulps so the values are upper bounds despite fp32 summation error
    const float en = sqrtf(e) * 1.0001f;
    if (lane == 0) {
      norm[nslot] = s;
      if (err) err[row] = en;
    }
    max_e = fmaxf(max_e, en);    // NaN-free: e >= 0 or +inf
    max_s = fmaxf(max_s, s);
  }
  if (max_stats) {
    __shared__ unsigned int sm[2];
    if (threadIdx.x < 2) sm[threadIdx.x] = 0u;
    __syncthreads();
    if (lane == 0) {
      atomicMax(&sm[0], __float_as_uint(max_e));   // non-negative floats order like their bit patterns
      atomicMax(&sm[1], __float_as_uint(max_s));
    }
    __syncthreads();
    if (threadIdx.x < 2) atomicMax(&max_stats[threadIdx.x], sm[threadIdx.x]);
  }
}

// One block per gallery tile: minima of the 256 norms over groups of 8 and of 32 columns (the epilogue's early-out bounds).
__global__ void __launch_bounds__(BN) pack_min_kernel(float* __restrict__ pack) {
  float* p = pack + size_t(blockIdx.x) * NPACK;
  float v = p[threadIdx.x];
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  if ((threadIdx.x & 7) == 0) p[BN + (threadIdx.x >> 3)] = v;
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 16));
  if ((threadIdx.x & 31) == 0) p[BN + 32 + (threadIdx.x >> 5)] = v;
}

// ------------------------------------------------------------------------------------------------ fused kernel
constexpr int NT = 4;                      // norm-pack ring slots
constexpr int A_ATOM_BYTES = BM * 128;     // 16 KiB
constexpr int B_STAGE_BYTES = BN * 128;    // 32 KiB
constexpr int NPIV = 16;                   // pivot pre-pass: the 16 smallest sampled keys per query define the ladder
constexpr int NPSUB = 8;                   // ... gathered from per-warp sub-lists of 8 (each warp scans a quarter of the columns)
constexpr int KPT = KP;                    // candidates that must lie below a pivot before it becomes the threshold

// MODE_RESWEEP: the sweep of the exact fallback (knn_fallback.cuh) -- same epilogue as MODE_SWEEP with a fixed threshold per
// row, but the number of queries is only known on the device: the geometry is derived in the kernel from *dyn_count.
enum { MODE_PIVOT = 0, MODE_SWEEP = 1, MODE_ASSIGN = 2, MODE_RESWEEP = 3 };

// Experiment counters (MMSIM_DEBUG_BUILD=1 builds only; scripts/sweep_debug.py): [5] 8-column groups handed to the
// candidate path, [6] epilogue warp cycles (sum over warps), [7] 32-column warp chunks with a hit
#ifdef MMSIM_SWEEP_DEBUG
// Round 2, per-role cycle breakdown of the sweep (lane 0 of the MMA warp; lane 0 of epilogue warp 4): [0] MMA warp total,
// [1] ... waiting for a drained accumulator, [2] ... waiting for a shared-memory stage, [3] epilogue warp waiting for a
// ready accumulator, [4] ... from "ready" to "released" (its TMEM loads), [8] ... total, [9] tiles (MMA warp)
__device__ unsigned long long g_dbg[16];
#define DBG_ADD(i, v) atomicAdd(&g_dbg[i], (unsigned long long)(v))
#define DBG_CLOCK() clock64()
#else
#define DBG_ADD(i, v) ((void)0)
#define DBG_CLOCK() 0ll
#endif

struct SweepArgs {
  const float* gpack;        // [n_tiles][NPACK]
  int nq, n_qblocks, n_tiles;
  int dual;                  // host side only: launch the DUAL instantiation (items = pairs of query blocks)
  // MODE_SWEEP: item = (split, query block); contiguous tile range per split
  int n_splits, tiles_per_split;
  int item_begin, item_end;  // MODE_SWEEP: this launch takes items [item_begin, item_end); item_end == 0: all of them
                             // (host-buffer mode launches one split at a time, as its gallery rows arrive)
  int use_pivots;            // 0: threshold +inf (everything is logged; small galleries)
  int close_rows;            // experiment (MMSIM_SWEEP_FLAGS=8): every row starts closed (threshold -inf): the product
                             // kernel's fast path with no candidate ever found (wrong results)
  uint2* log;                // [(row * n_splits + split) * logcap] {key bits, gallery row}
  int logcap;
  int* log_cnt;              // [row * n_splits + split] entries appended (may exceed logcap: overflow)
  float* log_tau;            // [row * n_splits + split] final threshold of the sweep
  int* split_done;           // [row * n_splits + split] bit 0: log_tau is published (later splits start from it);
                             // bits [1,16) / [16,31): rows this split logged below ladder rung piv1 / piv0
  // MODE_PIVOT: item = query block; the "gallery" is the compact sample block of n_sample_tiles tiles, swept whole
  int n_sample_tiles;
  float* piv16;              // MODE_PIVOT out: [row][16] the smallest sampled keys, ascending (+inf where missing)
  int* assign;               // MODE_ASSIGN out: [row] nearest row of the (small) "gallery" given -- the query's anchor
  const float* ladder;       // MODE_SWEEP in:  [row][4] = ladder pivots (ascending) and the initial threshold
  // MODE_RESWEEP: number of queries (device), its cap, and the inputs of resweep_geometry()
  const int* dyn_count;
  int dyn_cap, dyn_ctas;
  long long dyn_budget;
};

template <int KATOMS, bool DUAL = false>
struct Smem {
  static constexpr int NA = DUAL ? 2 : 1;                      // query tiles resident (DUAL: two 128-query blocks per CTA)
  static constexpr int NS = DUAL ? 4 : (KATOMS <= 2 ? 5 : 4);  // gallery smem stages (one 32 KiB K atom of one tile each)
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + NA * KATOMS * A_ATOM_BYTES;
  static constexpr int NORM_OFF = B_OFF + NS * B_STAGE_BYTES;
  static constexpr int TAU_OFF = NORM_OFF + NT * NPACK * 4;   // float [NA][BM]   current threshold of each row
  static constexpr int CNT_OFF = TAU_OFF + NA * BM * 4;       // u32   [NA][BM]   packed log cursor + ladder counters
  static constexpr int RS_OFF = CNT_OFF + NA * BM * 4;        // float2 [NA][BM]  DUAL: ladder rungs of each row (piv0, piv1)
  static constexpr int PV_OFF = RS_OFF + (DUAL ? NA * BM * 8 : 0);   // float [4][NPSUB][BM] pivot pre-pass: per-warp sorted sub-lists
  static constexpr int BAR_OFF = PV_OFF + (DUAL ? 0 : 4 * NPSUB * BM * 4);
  static constexpr int NUM_BARS = 4 * NS + 2 + 2 + 2 + NT + NT;   // PAIR runs twice the stages (half the bytes each)
  static constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int TOTAL = TMEM_PTR_OFF + 8;
  static constexpr int DYN_BYTES = TOTAL;
  static_assert(BAR_OFF % 8 == 0, "barriers must be 8-byte aligned");
  static_assert(DYN_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void epi_bar_sync(int nthreads) {  // named barrier 1: epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }


__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add32(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}

// Per-row state word (one native 32-bit shared atomic per candidate; a 64-bit shared add would be a CAS loop):
//   bits [ 0,14)  log cursor          bits [14,23)  logged entries below ladder pivot 1 (6th smallest sampled key)
//   bits [23,32)  logged entries below ladder pivot 0 (3rd smallest sampled key)
// A counter only moves while its pivot is still below the row threshold, so it stays below KPT + one tile of columns
// (< 512); the cursor is kept below 2^14 by closing the row (threshold = -inf) once its log is full.
constexpr uint32_t CUR_MASK = 0x3FFFu;
constexpr int CN1_SHIFT = 14, CN0_SHIFT = 23;

// Rare path of the sweep, out of line to keep the hot loop small: the 8 keys of one column group that has at least one
// candidate in the warp.  Every key below the row threshold is appended to the row's log: ONE native 32-bit shared
// atomic claims the slot and bumps the ladder counters, one 8-byte store.
// (Tried and dropped, both slower on B200: a logger warp fed through shared-memory queues -- a single consumer saturates,
// gpurun_out/sweep_debug.log -- and warp-private candidate rings drained in batches, gpurun_out/grouped_ring.log.)
__device__ __noinline__ void sweep_group8(float k0, float k1, float k2, float k3, float k4, float k5, float k6, float k7,
                                          int cbase, float tau, float piv0, float piv1, uint2* mylog, int logcap,
                                          uint32_t cnt_addr, uint32_t tau_addr) {
  const float key[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (key[j] < tau) {
      const uint32_t inc = 1u | (key[j] < piv1 && piv1 < tau ? (1u << CN1_SHIFT) : 0u) |
                           (key[j] < piv0 && piv0 < tau ? (1u << CN0_SHIFT) : 0u);
      const int slot = int(atoms_add32(cnt_addr, inc) & CUR_MASK);
      if (slot < logcap) {
        mylog[slot] = make_uint2(__float_as_uint(key[j]), uint32_t(cbase + j));
      } else {
        sts_f32(tau_addr, -kInf);   // log full: the row is uncertifiable from here on, stop accepting candidates
      }
    }
  }
}

// Rare path of the pivot pre-pass: insert x into this thread's sorted sub-list of the NPSUB smallest keys (lane-private
// column of shared memory, stride BM floats); returns the new NPSUB-th smallest.
__device__ __noinline__ float pivot_insert(float x, uint32_t pv_addr) {
  float p[NPSUB];
#pragma unroll
  for (int i = 0; i < NPSUB; ++i) p[i] = lds_f32(pv_addr + i * BM * 4);
#pragma unroll
  for (int i = 0; i < NPSUB; ++i) {
    const float lo = fminf(p[i], x);
    x = fmaxf(p[i], x);
    p[i] = lo;
  }
#pragma unroll
  for (int i = 0; i < NPSUB; ++i) sts_f32(pv_addr + i * BM * 4, p[i]);
  return p[NPSUB - 1];
}

template <int MODE>
__device__ __forceinline__ int tile_of(const SweepArgs& a, int t0, int i) {
  return t0 + i;
}

// ABL: ablation variants for scripts/sweep_ablate.py (wrong results by construction): 0 = product, 2 = never log a
// candidate (first-level test only), 4 = drain the accumulator without scanning it.
// PAIR: two CTAs of a cluster (adjacent SMs) work as one tcgen05 cta_group::2 unit: each has its own 128 queries and its
// own accumulators, but a 256-row gallery tile is loaded ONCE per pair -- each CTA's TMA brings half of it into its own
// shared memory and the pair's MMAs (issued by the leader CTA, M = 256) read both halves.  Per CTA and tile that halves
// the TMA writes into shared memory and the L2 -> SM traffic; the 1-CTA kernel moves 160 KB through a 128 B/cycle shared
// memory per 1,024-cycle tile (64 KB TMA writes + 96 KB operand reads), the pair 96 KB.
// DUAL (round 2): a CTA keeps TWO 128-query blocks resident and multiplies every gallery tile with both before releasing its
// shared-memory stage -- the sweep is bound by L2 -> SM bytes (782 query blocks each stream the fp16 gallery: 204 GB per
// launch = the ~6,300 B/cycle the L2 delivers chip-wide, profiles/r2h_sweep_roles.txt), and this halves them.  The two
// accumulators of a gallery tile are the two TMEM stages; everything per (query block, tile) is unchanged.  K <= 128 only
// (two 32 KiB query tiles + a four-stage gallery ring fit in shared memory).
template <int KATOMS, int NEPI, int MODE, int ABL, bool PAIR, bool DUAL = false>
__global__ void __launch_bounds__(128 + NEPI * 32, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_g, const SweepArgs a) {
  static_assert(!DUAL || (!PAIR && KATOMS <= 2 && MODE == MODE_SWEEP), "DUAL: product sweep, K <= 128, single CTAs");
  using S = Smem<KATOMS, DUAL>;
  constexpr int NA = DUAL ? 2 : 1;
  constexpr int NS = PAIR ? 2 * S::NS : S::NS;   // PAIR: a stage holds half a tile's K atom (16 KiB), so twice as many fit
  constexpr int NH = NEPI / 4;            // epilogue warps per TMEM lane quarter; each takes every NH-th 32-column chunk
  constexpr int EPI_THREADS = NEPI * 32;
  // No static shared memory in this kernel, so the dynamic window starts at the CTA's shared base and the declared
  // alignment holds (checked below); keeping the pointer un-rounded lets the compiler emit LDS/STS/ATOMS.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();

  uint8_t* smem_a = smem + S::A_OFF;
  uint8_t* smem_b = smem + S::B_OFF;
  float* norm_ring = reinterpret_cast<float*>(smem + S::NORM_OFF);
  float* s_tau = reinterpret_cast<float*>(smem + S::TAU_OFF);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem + S::CNT_OFF);
  [[maybe_unused]] float2* s_rs = reinterpret_cast<float2*>(smem + S::RS_OFF);   // DUAL: (piv0, piv1) of every resident row
  float* s_pv = reinterpret_cast<float*>(smem + S::PV_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* full = bars;                 // [NS]  TMA -> MMA
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* tfull = bars + 2 * NS;       // [2]   MMA -> epilogue (accumulator ready)
  uint64_t* tempty = bars + 2 * NS + 2;  // [2]   epilogue -> MMA (accumulator drained)
  uint64_t* afull = bars + 2 * NS + 4;   //       query tile landed
  uint64_t* aempty = bars + 2 * NS + 5;  //       all MMAs of the item done, query tile may be overwritten
  uint64_t* nempty = bars + 2 * NS + 6;  // [NT]  epilogue finished a tile: its norm-pack slot may be refilled
  uint64_t* nfull = bars + 2 * NS + 6 + NT;  // [NT]  PAIR: this CTA's copy of a tile's norm pack landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::TMEM_PTR_OFF);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;   // rank in the CTA pair; 0 = leader (issues the MMAs)
  constexpr int BROWS = PAIR ? BN / 2 : BN;                    // gallery rows of a tile in THIS CTA's shared memory
  constexpr int B_BYTES = BROWS * 128;                         // ... per K atom

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_g);
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      // PAIR: one elected arrival per epilogue warp of BOTH CTAs (they arrive on the leader's barrier)
      ptx::mbar_init(&tempty[i], PAIR ? 2 * NEPI : EPI_THREADS);
    }
    ptx::mbar_init(afull, 1);
    ptx::mbar_init(aempty, 1);
    for (int i = 0; i < NT; ++i) {
      ptx::mbar_init(&nempty[i], EPI_THREADS);
      ptx::mbar_init(&nfull[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {                                     // 512 columns: two 128x256 fp32 accumulators
    if (PAIR) ptx::tmem_alloc_pair(tmem_ptr, 2 * BN); else ptx::tmem_alloc(tmem_ptr, 2 * BN);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();                       // the peer's barriers are initialised before anyone signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  constexpr bool SWEEP = MODE == MODE_SWEEP || MODE == MODE_RESWEEP;
  int nq = a.nq, n_qblocks = a.n_qblocks, n_splits = a.n_splits, tiles_per_split = a.tiles_per_split;
  if (MODE == MODE_RESWEEP) {   // geometry from the device-side count of uncertified queries (uniform over the grid)
    const DynGeom dg = resweep_geometry(min(*a.dyn_count, a.dyn_cap), a.n_tiles, a.dyn_ctas, a.dyn_budget);
    nq = dg.nq; n_qblocks = dg.n_qblocks; n_splits = dg.n_splits; tiles_per_split = dg.tiles_per_split;
  }

  // items: (split, query block) -- PAIR: (split, PAIR of query blocks), this CTA takes block 2 * pair + rank
  const int n_qb_items = (PAIR || DUAL) ? (n_qblocks + 1) / 2 : n_qblocks;   // DUAL: this CTA takes blocks 2 * item, 2 * item + 1
  const int n_items = SWEEP ? (a.item_end ? a.item_end : n_qb_items * n_splits) : n_qblocks;
  const int item0 = (MODE == MODE_SWEEP ? a.item_begin : 0) + (PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x));
  const int item_step = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);
  auto item_range = [&](int item, int& qb, int& split, int& t0, int& nt) {
    if (SWEEP) {
      split = item / n_qb_items;
      qb = item - split * n_qb_items;
      if (PAIR) qb = 2 * qb + int(crank);              // may be == n_qblocks (odd count): all rows invalid, loads zero-filled
      if (DUAL) qb = 2 * qb;                           // ... and block qb + 1 (same remark)
      t0 = split * tiles_per_split;
      nt = min(a.n_tiles, t0 + tiles_per_split) - t0;
    } else {
      split = 0; qb = item; t0 = 0; nt = a.n_sample_tiles;
    }
  };

  const uint32_t wg = __shfl_sync(0xffffffffu, threadIdx.x >> 7, 0);   // warpgroup, provably uniform
  if (wg == 0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // =============================================================== TMA producer (one thread)
    if (lane == 0) {
      uint32_t it = 0, tc = 0, ic = 0;
      for (int item = item0; item < n_items; item += item_step, ++ic) {
        int qb, split, t0, nt;
        item_range(item, qb, split, t0, nt);
        ptx::mbar_wait(aempty, (ic & 1) ^ 1);
        if (!PAIR) {
          ptx::mbar_expect_tx(afull, NA * KATOMS * A_ATOM_BYTES);
          for (int hb = 0; hb < NA; ++hb)
            for (int ka = 0; ka < KATOMS; ++ka)
              ptx::tma_load_2d(smem_a + (hb * KATOMS + ka) * A_ATOM_BYTES, &tm_q, ka * KATOM, (qb + hb) * BM, afull);
        } else {
          // both CTAs' query tiles are counted on the leader's barrier, which expects the bytes of the pair
          if (crank == 0) ptx::mbar_expect_tx(afull, 2 * KATOMS * A_ATOM_BYTES);
          for (int ka = 0; ka < KATOMS; ++ka)
            ptx::tma_load_2d_pair(smem_a + ka * A_ATOM_BYTES, &tm_q, ka * KATOM, qb * BM, afull);
        }
        for (int i = 0; i < nt; ++i, ++tc) {
          const int t = tile_of<MODE>(a, t0, i);
          for (int ka = 0; ka < KATOMS; ++ka, ++it) {
            const uint32_t stage = it % NS, phase = (it / NS) & 1;
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            if (!PAIR) {
              ptx::mbar_expect_tx(&full[stage], B_STAGE_BYTES + (ka == 0 ? NPACK * 4 : 0));
              ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tm_g, ka * KATOM, t * BN, &full[stage]);
            } else {
              if (crank == 0) ptx::mbar_expect_tx(&full[stage], 2 * B_BYTES);
              ptx::tma_load_2d_pair(smem_b + stage * B_BYTES, &tm_g, ka * KATOM, t * BN + int(crank) * BROWS, &full[stage]);
            }
            if (ka == 0) {
              // the epilogue reads this tile's norms long after it released the accumulator: separate hand-back
              ptx::mbar_wait(&nempty[tc % NT], ((tc / NT) & 1) ^ 1);
              if (!PAIR) {
                ptx::bulk_load_1d(norm_ring + (tc % NT) * NPACK, a.gpack + size_t(t) * NPACK, NPACK * 4, &full[stage]);
              } else {   // every CTA needs the whole tile's pack: its own copy, its own barrier
                ptx::mbar_expect_tx(&nfull[tc % NT], NPACK * 4);
                ptx::bulk_load_1d(norm_ring + (tc % NT) * NPACK, a.gpack + size_t(t) * NPACK, NPACK * 4, &nfull[tc % NT]);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer (one thread issues, warp waits)
    // (Round 2 tried TWO issuer warps, one per accumulator stage, to spread the issue path -- barrier waits, commits and the
    // ~110 cycles every tcgen05.mma costs whatever its shape: 2-3% faster at 128-d when it worked, but two consumers of one
    // mbarrier ring can fall a phase apart -- a parity wait only tells consecutive phases apart -- and it deadlocked at
    // K = 256 and together with DUAL; removed.  profiles/r2i_sweep_dual_mma2.txt)
    // PAIR: only the leader CTA issues (M = 256 over both CTAs' query tiles); the peer's warp 1 just owns its TMEM
    const uint32_t idesc = ptx::umma_idesc_f16(PAIR ? 2 * BM : BM, BN);
    const uint64_t adesc0 = ptx::umma_desc_k128(ptx::smem_u32(smem_a));
    const uint64_t bdesc0 = ptx::umma_desc_k128(ptx::smem_u32(smem_b));
    uint32_t it = 0, tc = 0, ic = 0;
    [[maybe_unused]] const long long t_mma0 = DBG_CLOCK();
    if (!PAIR || crank == 0) {
    for (int item = item0; item < n_items; item += item_step, ++ic) {
      int qb, split, t0, nt;
      item_range(item, qb, split, t0, nt);
      ptx::mbar_wait(afull, ic & 1);
      ptx::tc_fence_after();
      for (int i = 0; i < nt; ++i) {
        const uint32_t it0 = it;
#pragma unroll
        for (int hb = 0; hb < NA; ++hb, ++tc) {          // DUAL: both query blocks against this gallery tile (tc counts MMA tiles)
        const uint32_t as = tc & 1;
        [[maybe_unused]] const long long w0 = DBG_CLOCK();
        ptx::mbar_wait(&tempty[as], ((tc >> 1) & 1) ^ 1);   // (a spinning wait here is slower: 22.3 vs 21.2 ms)
        ptx::tc_fence_after();
        if (MODE == MODE_SWEEP && lane == 0) { DBG_ADD(1, DBG_CLOCK() - w0); DBG_ADD(9, 1); }
        it = it0;
        for (int ka = 0; ka < KATOMS; ++ka, ++it) {
          const uint32_t stage = it % NS, phase = (it / NS) & 1;
          [[maybe_unused]] const long long w1 = DBG_CLOCK();
          if (hb == 0) {                                 // (the second block finds the stages this thread already waited for)
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
          }
          if (MODE == MODE_SWEEP && lane == 0) DBG_ADD(2, DBG_CLOCK() - w1);
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < KATOM / 16; ++k) {
              const uint64_t ad = adesc0 + uint64_t((hb * KATOMS + ka) * (A_ATOM_BYTES >> 4) + k * 2);
              const uint64_t bd = bdesc0 + uint64_t(stage * (B_BYTES >> 4) + k * 2);
              if (PAIR) ptx::umma_f16_pair(tmem_base + as * BN, ad, bd, idesc, (ka | k) != 0);
              else ptx::umma_f16(tmem_base + as * BN, ad, bd, idesc, (ka | k) != 0);
            }
            if (PAIR) {
              ptx::umma_commit_pair(&empty[stage]);                      // both CTAs' producers
              if (ka == KATOMS - 1) ptx::umma_commit_pair(&tfull[as]);   // both CTAs' epilogue warps
            } else {
              if (hb == NA - 1) ptx::umma_commit(&empty[stage]);    // smem stage reusable once these MMAs retire
              if (ka == KATOMS - 1) ptx::umma_commit(&tfull[as]);   // accumulator complete
            }
          }
          __syncwarp();
        }
        }
      }
      if (lane == 0) {
        if (PAIR) ptx::umma_commit_pair(aempty); else ptx::umma_commit(aempty);
      }
      __syncwarp();
    }
    if (MODE == MODE_SWEEP && lane == 0) DBG_ADD(0, DBG_CLOCK() - t_mma0);
    }
  }
  } else {
    // =============================================================== epilogue warps: TMEM -> candidates
    if constexpr (NEPI >= 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");     // 8 warps: four 32-register accumulator buffers each
    const uint32_t q = warp & 3;              // TMEM lane quarter this warp may read
    const uint32_t h = (warp - 4) >> 2;       // which of the NH warps of this quarter
    const int row = q * 32 + lane;            // row of the CTA tile == TMEM lane
    const uint32_t ring_u32 = ptx::smem_u32(norm_ring);
#ifdef MMSIM_SWEEP_DEBUG
    const long long t_epi0 = clock64();
#endif
    uint32_t tc = 0;

    for (int item = item0; item < n_items; item += item_step) {
      int qb, split, t0, nt;
      item_range(item, qb, split, t0, nt);
      const bool qb_ok = !PAIR || qb < n_qblocks;      // PAIR: the second CTA of the last pair may have no block
      const int grow = (qb_ok ? qb : 0) * BM + row;      // global query row (may be >= nq in the last block)
      const bool valid = qb_ok && grow < nq;

      float piv0 = -kInf, piv1 = -kInf;                 // MODE_SWEEP: ladder below the initial threshold
      uint32_t carry = 0;                               // MODE_SWEEP: rows finished splits logged below piv1 | piv0 << 15
      float pv_thr = kInf;                              // MODE_PIVOT: current NPIV-th smallest sampled key of the row
      float as_best = kInf;                             // MODE_ASSIGN: smallest key of the row so far in this thread's chunks,
      int as_col = 0x7fffffff;                          //              and its column
      // (DUAL: these follow the query block whose accumulator is being scanned -- set at the top of every MMA tile below)
      uint32_t tau_addr = ptx::smem_u32(s_tau + row), cnt_addr = ptx::smem_u32(s_cnt + row);
      uint2* mylog = a.log + (size_t(grow) * n_splits + split) * a.logcap;
      [[maybe_unused]] const uint32_t tau_addr0 = tau_addr, cnt_addr0 = cnt_addr, rs_addr0 = ptx::smem_u32(s_rs + row);
      [[maybe_unused]] uint2* const mylog0 = mylog;
      [[maybe_unused]] const size_t log_half = size_t(BM) * n_splits * a.logcap;   // log entries between the two query blocks' rows
      const uint32_t pv_addr = ptx::smem_u32(s_pv + (h * NPSUB) * BM + row);
      if constexpr (DUAL) {
        static_assert(!DUAL || NH >= 2, "DUAL: warp h of a lane quarter prepares / publishes the rows of query block qb + h");
        epi_bar_sync(EPI_THREADS);             // every epilogue warp is done with the previous item
        if (h < 2) {
          const int g = grow + int(h) * BM;
          const bool ok = qb + int(h) < n_qblocks && g < nq;
          float tau0 = kInf, p0 = -kInf, p1 = -kInf;
          uint32_t cy = 0;
          if (ok) {
            if (a.use_pivots) {
              const float4 pp = *reinterpret_cast<const float4*>(a.ladder + size_t(g) * 4);
              p0 = pp.y; p1 = pp.z; tau0 = pp.w;
            }
            for (int s2 = 0; s2 < n_splits; ++s2) {   // thresholds and rung counters of the finished splits (see below)
              if (s2 == split) continue;
              const size_t o = size_t(g) * n_splits + s2;
              uint32_t done;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(a.split_done + o) : "memory");
              if (done) {
                tau0 = fminf(tau0, __ldcg(a.log_tau + o));
                cy += done >> 1;
              }
            }
            cy = min(cy & 0x7fffu, uint32_t(KPT)) | (min(cy >> 15, uint32_t(KPT)) << 15);
          }
          if (a.close_rows) tau0 = -kInf;
          s_tau[h * BM + row] = ok ? tau0 : -kInf;
          s_cnt[h * BM + row] = ((cy & 0x7fffu) << CN1_SHIFT) | ((cy >> 15) << CN0_SHIFT);
          s_rs[h * BM + row] = make_float2(p0, p1);
          carry = cy;                          // this warp publishes the row's counters when the item ends
        }
        epi_bar_sync(EPI_THREADS);
      } else
      if (SWEEP) {
        float tau0 = kInf;
        if (a.use_pivots) {
          const float4 pp = *reinterpret_cast<const float4*>(a.ladder + size_t(grow) * 4);
          piv0 = pp.y; piv1 = pp.z; tau0 = pp.w;        // ladder rungs and initial threshold (make_ladder_kernel)
        }
        // A finished sweep of another gallery split of the same query ended with a threshold that is usually much
        // tighter than the sampled one: start from it.  Items are ordered split-major, so with more query blocks than
        // SMs the earlier splits of this block finished waves ago.  (Thresholds only steer how many candidates are
        // kept; the certificate in the rerank kernel bounds every unlogged row by the smallest threshold in force.)
        // It also publishes how many rows it logged below each rung: the ladder counters continue across splits instead
        // of restarting (a split on its own rarely sees KPT rows below the lower rung).
        for (int s2 = 0; MODE == MODE_SWEEP && s2 < n_splits; ++s2) {   // (the re-sweep's thresholds are fixed)
          if (s2 == split) continue;
          const size_t o = size_t(grow) * n_splits + s2;
          uint32_t done;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(a.split_done + o) : "memory");
          if (done) {
            tau0 = fminf(tau0, __ldcg(a.log_tau + o));
            carry += done >> 1;                // two 15-bit fields, each at most 511 per split and at most 8 splits
          }
        }
        carry = min(carry & 0x7fffu, uint32_t(KPT)) | (min(carry >> 15, uint32_t(KPT)) << 15);
        if (a.close_rows) tau0 = -kInf;
        epi_bar_sync(EPI_THREADS);             // every epilogue warp is done with the previous item
        if (h == 0) {
          s_tau[row] = valid ? tau0 : -kInf;   // rows past the last query never accept a candidate
          s_cnt[row] = ((carry & 0x7fffu) << CN1_SHIFT) | ((carry >> 15) << CN0_SHIFT);
        }
        epi_bar_sync(EPI_THREADS);
      } else if (MODE == MODE_PIVOT) {
        static_assert(MODE == MODE_SWEEP || NH <= 4, "pivot sub-lists are laid out for at most 4 warps per quarter");
        epi_bar_sync(EPI_THREADS);             // the previous item's merge has read every sub-list
        pv_thr = valid ? kInf : -kInf;
#pragma unroll
        for (int i = 0; i < NPSUB; ++i) s_pv[(h * NPSUB + i) * BM + row] = pv_thr;
      } else {
        epi_bar_sync(EPI_THREADS);             // the previous item's merge has read every partial result
      }

      // One 32-column chunk of the accumulator, already in registers.  Level 1 (every chunk): minimum of the raw
      // accumulator over the chunk against tau - (smallest |g|^2 of the chunk): key_j = acc_j + |g_j|^2 >= acc_j + min |g|^2,
      // so no key of the chunk can be below tau when the test fails -- 18 FMNMX3, one LDS, one FADD, one FSETP, one vote
      // per 32 columns.  Level 2 (warp has a hit): the same test per 8-column group.  Level 3 (out of line): the 8 keys.
      auto scan_chunk = [&](float (&v)[32], int c, uint32_t nrm, int col0, float tau) {
        if (MODE == MODE_ASSIGN) {             // running argmin of key = |g|^2 - 2 q.g over this thread's columns
          // chunk-level early out like the sweep's: no key of the chunk is below min(acc) + min |g|^2
          float m = v[0];
#pragma unroll
          for (int j = 1; j < 31; j += 2) m = min3(m, v[j], v[j + 1]);
          m = fminf(m, v[31]);
          if (!(m + lds_f32(nrm + (BN + 32 + c) * 4) < as_best)) return;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 n = lds_f32x4(nrm + (c * 32 + g * 4) * 4);
            const float k4[4] = {v[g * 4] + n.x, v[g * 4 + 1] + n.y, v[g * 4 + 2] + n.z, v[g * 4 + 3] + n.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k4[j] < as_best) { as_best = k4[j]; as_col = col0 + c * 32 + g * 4 + j; }   // ascending columns: ties keep the first
          }
          return;
        }
        if (ABL == 4) return;
        float gm[4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          gm[g] = min3(min3(v[g * 8], v[g * 8 + 1], v[g * 8 + 2]), min3(v[g * 8 + 3], v[g * 8 + 4], v[g * 8 + 5]),
                       fminf(v[g * 8 + 6], v[g * 8 + 7]));
        const float m = fminf(min3(gm[0], gm[1], gm[2]), gm[3]);
        const float nm32 = lds_f32(nrm + (BN + 32 + c) * 4);
        const float bound = tau - nm32;
        if (!__any_sync(0xffffffffu, m < bound) || ABL == 2) return;
        if (SWEEP) {
          // Level 2 (warp has a hit): the same test per 8-column group; level 3 (out of line): the 8 keys of a group.
          // (The query operand was pre-scaled by -2: acc = -2 q.g, key = |g|^2 - 2 q.g.)
          // (Tried in round 2: levels 2 + 3 as ONE out-of-line function fed through a local-memory copy of the chunk, to
          // shrink the hot loop's code from 6.5 KB to 3.3 KB -- 27.4 ms instead of 22.7: the 32-register spill + reload per
          // candidate chunk costs far more than the footprint; profiles/r2_sweep_experiments.txt.)
          const float4 nm8 = lds_f32x4(nrm + (BN + c * 4) * 4);
          const uint32_t mine = (gm[0] < tau - nm8.x ? 1u : 0u) | (gm[1] < tau - nm8.y ? 2u : 0u) |
                                (gm[2] < tau - nm8.z ? 4u : 0u) | (gm[3] < tau - nm8.w ? 8u : 0u);
          const uint32_t groups = __reduce_or_sync(0xffffffffu, mine);
          if (lane == 0) { DBG_ADD(7, 1); DBG_ADD(5, __popc(groups)); }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!(groups & (1u << g))) continue;
            const float4 n0 = lds_f32x4(nrm + (c * 32 + g * 8) * 4);
            const float4 n1 = lds_f32x4(nrm + (c * 32 + g * 8 + 4) * 4);
            sweep_group8(v[g * 8 + 0] + n0.x, v[g * 8 + 1] + n0.y, v[g * 8 + 2] + n0.z, v[g * 8 + 3] + n0.w,
                         v[g * 8 + 4] + n1.x, v[g * 8 + 5] + n1.y, v[g * 8 + 6] + n1.z, v[g * 8 + 7] + n1.w,
                         col0 + c * 32 + g * 8, tau, piv0, piv1, mylog, a.logcap, cnt_addr, tau_addr);
          }
        } else {
          const float4 nm8 = lds_f32x4(nrm + (BN + c * 4) * 4);
          const uint32_t mine = (gm[0] < tau - nm8.x ? 1u : 0u) | (gm[1] < tau - nm8.y ? 2u : 0u) |
                                (gm[2] < tau - nm8.z ? 4u : 0u) | (gm[3] < tau - nm8.w ? 8u : 0u);
          const uint32_t groups = __reduce_or_sync(0xffffffffu, mine);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!(groups & (1u << g))) continue;
            const float4 n0 = lds_f32x4(nrm + (c * 32 + g * 8) * 4);
            const float4 n1 = lds_f32x4(nrm + (c * 32 + g * 8 + 4) * 4);
            const float key[8] = {v[g * 8 + 0] + n0.x, v[g * 8 + 1] + n0.y, v[g * 8 + 2] + n0.z, v[g * 8 + 3] + n0.w,
                                  v[g * 8 + 4] + n1.x, v[g * 8 + 5] + n1.y, v[g * 8 + 6] + n1.z, v[g * 8 + 7] + n1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (key[j] < pv_thr) pv_thr = pivot_insert(key[j], pv_addr);
          }
        }
      };

      const uint32_t taddr0 = tmem_base + ((q * 32) << 16) + h * 32;   // this warp's first chunk of accumulator stage 0
      constexpr int CPW = (BN / 32) / NH;    // chunks per warp and tile (even)
      for (int i = 0; i < nt; ++i) {
#pragma unroll 1
      for (int hb = 0; hb < NA; ++hb, ++tc) {   // DUAL: the accumulators of query blocks qb, qb + 1 for this gallery tile
        if constexpr (DUAL) {
          tau_addr = tau_addr0 + hb * BM * 4;
          cnt_addr = cnt_addr0 + hb * BM * 4;
          mylog = mylog0 + size_t(hb) * log_half;
          const float2 pp = lds_f32x2(rs_addr0 + hb * BM * 8);
          piv0 = pp.x; piv1 = pp.y;
        }
        const uint32_t gtc = DUAL ? tc >> 1 : tc;   // gallery tiles seen (norm-pack ring, tightening turns)
        const uint32_t as = tc & 1;
        [[maybe_unused]] const long long e0 = DBG_CLOCK();
        ptx::mbar_wait(&tfull[as], (tc >> 1) & 1);
        ptx::tc_fence_after();
        [[maybe_unused]] const long long e1 = DBG_CLOCK();
        if (MODE == MODE_SWEEP && warp == 4 && lane == 0) DBG_ADD(3, e1 - e0);
        const uint32_t nrm = ring_u32 + (gtc % NT) * NPACK * 4;
        const uint32_t taddr = taddr0 + as * BN;
        const int col0 = tile_of<MODE>(a, t0, i) * BN;
        // Two register buffers.  Both loads of a pair are issued before either chunk is scanned, and the TMEM stage is
        // handed back to the MMA warp as soon as this warp's LAST load has landed -- before anything is scanned: the
        // MMA warp and the epilogue warps wait on each other once per tile, and whatever sits between the accumulator
        // becoming ready and its release (a scan, worse: a scan that finds candidates) is on the critical path of the
        // whole SM (profiles/r1_e_sweep.txt: the MMA warp spends 39% of its time waiting for this release).
        // (Tried and dropped: accumulating a tile as two 128 x 128 halves with their own ready / drained barriers, four
        // half-accumulators in flight -- 29 ms instead of 20.6, gpurun_out/ablate_v6.log: N = 128 MMAs and twice the
        // barrier traffic cost more than the finer hand-off saves.)
        // (Round 2, scripts/ubench/tmem_mma.cu: the hand-off itself costs ~130 cycles per tile whatever the waits look like;
        // on top of it 16 warps with two loads in flight each run at ~1,350 cycles per tile, 8 warps with four at ~1,200.
        // NB register buffers == loads in flight: four when a warp owns four or more chunks of a tile.)
        constexpr int NB = CPW >= 4 ? 4 : 2;
        float v[NB][32];
#pragma unroll
        for (int b = 0; b < NB; ++b) ptx::tmem_ld32(taddr + b * NH * 32, v[b]);
#pragma unroll 1
        for (int cp = 0; cp < CPW; cp += NB) {
          const bool last = cp + NB >= CPW;
#pragma unroll
          for (int b = 0; b < NB; ++b) ptx::tmem_ld_wait(v[b]);
          if (last) {
            ptx::tc_fence_before();
            if (PAIR) {
              __syncwarp();                                   // every lane's loads have landed
              if (lane == 0) ptx::mbar_arrive_leader(&tempty[as]);
              ptx::mbar_wait(&nfull[tc % NT], (tc / NT) & 1); // this CTA's copy of the tile's norm pack
            } else {
              ptx::mbar_arrive(&tempty[as]);
            }
            if (MODE == MODE_SWEEP && warp == 4 && lane == 0) DBG_ADD(4, DBG_CLOCK() - e1);
          }
          const float tau = SWEEP ? lds_f32(tau_addr) : pv_thr;
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            scan_chunk(v[b], h + (cp + b) * NH, nrm, col0, SWEEP ? tau : pv_thr);
            if (!last) ptx::tmem_ld32(taddr + (cp + NB + b) * NH * 32, v[b]);
          }
        }
        if (MODE == MODE_SWEEP && h == (gtc & (NH - 1))) {
          // tighten (the NH warps of a row take turns): once KPT logged entries lie below a pivot, the KPT smallest
          // keys all lie below it
          const int r_ = (DUAL ? hb * BM : 0) + row;
          const uint32_t cn = s_cnt[r_];
          const float ot = s_tau[r_];
          float nt_ = ot;
          if (((cn >> CN1_SHIFT) & 511u) >= KPT) nt_ = fminf(nt_, piv1);
          if (((cn >> CN0_SHIFT) & 511u) >= KPT) nt_ = fminf(nt_, piv0);
          if (nt_ < ot) s_tau[r_] = nt_;
        }
        if (!DUAL || hb == NA - 1) ptx::mbar_arrive(&nempty[gtc % NT]);
      }
      }

      if constexpr (DUAL) {
        epi_bar_sync(EPI_THREADS);             // all appends of this item are done
        if (h < 2 && qb + int(h) < n_qblocks) {
          const size_t o = size_t(grow + int(h) * BM) * n_splits + split;
          const uint32_t cn = s_cnt[h * BM + row];
          a.log_cnt[o] = int(cn & CUR_MASK);
          a.log_tau[o] = s_tau[h * BM + row];
          if (n_splits > 1) {
            const uint32_t own1 = ((cn >> CN1_SHIFT) & 511u) - (carry & 0x7fffu), own0 = ((cn >> CN0_SHIFT) & 511u) - (carry >> 15);
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.split_done + o), "r"(1u | (own1 << 1) | (own0 << 16))
                         : "memory");
          }
        }
      } else if (SWEEP) {
        epi_bar_sync(EPI_THREADS);             // all appends of this item are done
        if (h == 0 && qb_ok) {
          const size_t o = size_t(grow) * n_splits + split;
          const uint32_t cn = s_cnt[row];
          a.log_cnt[o] = int(cn & CUR_MASK);
          a.log_tau[o] = s_tau[row];
          if (MODE == MODE_SWEEP && n_splits > 1) {
            // Thresholds only steer how many candidates are kept: the final certificate (rerank kernel) bounds every
            // unlogged row by the smallest threshold in force, so a shared threshold can cost a fallback, never exactness.
            const uint32_t own1 = ((cn >> CN1_SHIFT) & 511u) - (carry & 0x7fffu), own0 = ((cn >> CN0_SHIFT) & 511u) - (carry >> 15);
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.split_done + o), "r"(1u | (own1 << 1) | (own0 << 16))
                         : "memory");
          }
        }
      } else if (MODE == MODE_ASSIGN) {
        // merge the NH partial results of a row: smallest (key, column)
        s_pv[(2 * h) * BM + row] = as_best;
        reinterpret_cast<int*>(s_pv)[(2 * h + 1) * BM + row] = as_col;
        epi_bar_sync(EPI_THREADS);
        if (h == 0 && valid) {
          float b = kInf;
          int bc = 0x7fffffff;
          for (int e = 0; e < NH; ++e) {
            const float x = s_pv[(2 * e) * BM + row];
            const int xc = reinterpret_cast<const int*>(s_pv)[(2 * e + 1) * BM + row];
            if (x < b || (x == b && xc < bc)) { b = x; bc = xc; }
          }
          a.assign[grow] = bc == 0x7fffffff ? 0 : bc;
        }
      } else {
        // merge the NH sub-lists: the ladder is the 2nd / 4th / 8th / 16th smallest of their union.  (A sub-list keeps
        // only 8 keys, so the union can miss a few of the true 16 smallest: the pivots are steering values, not bounds.)
        epi_bar_sync(EPI_THREADS);
        if (h == 0) {
          float m[NPIV];
#pragma unroll
          for (int i = 0; i < NPIV; ++i) m[i] = kInf;
          for (int e = 0; e < NH * NPSUB; ++e) {
            float x = s_pv[e * BM + row];
#pragma unroll
            for (int i = 0; i < NPIV; ++i) {
              const float lo = fminf(m[i], x);
              x = fmaxf(m[i], x);
              m[i] = lo;
            }
          }
          float4* dst = reinterpret_cast<float4*>(a.piv16 + size_t(grow) * NPIV);
#pragma unroll
          for (int i = 0; i < NPIV / 4; ++i) dst[i] = make_float4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]);
        }
      }
    }
#ifdef MMSIM_SWEEP_DEBUG
    if (lane == 0) DBG_ADD(6, clock64() - t_epi0);
    if (MODE == MODE_SWEEP && warp == 4 && lane == 0) DBG_ADD(8, clock64() - t_epi0);
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();       // the peer may still be reading accumulators the pair's allocation covers
  if (warp == 1) {
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 2 * BN); else ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace knn
}  // namespace mmsim
#include "knn_sweepq.cuh"
namespace mmsim {
namespace knn {

// ------------------------------------------------------------------------------------------------ query grouping
// The sweep tests 32 query rows x 32 gallery columns per warp step, and a step that finds a candidate costs several times
// a step that finds none.  With queries in arbitrary order about 40% of the steps find one (every query has its own ~700
// candidates); with similar queries side by side their candidates are the same gallery rows and the share drops to 4%
// (gpurun_out/sweep_debug2.log) -- 27.0 -> 21.8 ms on the benchmark (gpurun_out/ladder_v4.log).  So the queries are
// sorted by their nearest ANCHOR before the sweep: anchors = n_anchor evenly spaced query rows, nearest anchor found by
// the tensor-core kernel in MODE_ASSIGN (the anchors are its "gallery": a few tiles), then a stable counting sort.  The
// permutation depends on the queries only (deterministic: every gallery shard derives the same one), results are
// scattered back to the caller's order by the re-rank kernel, and nothing about exactness changes.
constexpr int GROUP_BLOCK = 1024;          // queries per block of the counting sort

__global__ void anchor_index_kernel(int* __restrict__ idx, int n_anchor, int64_t nq) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_anchor) idx[j] = int((int64_t(j) * nq) / n_anchor);
}

// blockhist[bin * gridDim.x + block] = queries of this block's chunk assigned to anchor `bin`
__global__ void __launch_bounds__(GROUP_BLOCK)
group_hist_kernel(const int* __restrict__ assign, int nq, int n_anchor, int* __restrict__ blockhist) {
  extern __shared__ int gh_smem[];
  for (int b = threadIdx.x; b < n_anchor; b += blockDim.x) gh_smem[b] = 0;
  __syncthreads();
  const int i = blockIdx.x * GROUP_BLOCK + threadIdx.x;
  if (i < nq) atomicAdd(&gh_smem[assign[i]], 1);
  __syncthreads();
  for (int b = threadIdx.x; b < n_anchor; b += blockDim.x) blockhist[size_t(b) * gridDim.x + blockIdx.x] = gh_smem[b];
}

// one warp per anchor: in-place exclusive scan of its row of blockhist (over the sort's blocks) + the anchor's total
__global__ void __launch_bounds__(256)
group_binscan_kernel(int* __restrict__ blockhist, int nblk, int n_anchor, int* __restrict__ bin_total) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= n_anchor) return;
  int* row = blockhist + size_t(b) * nblk;
  int run = 0;
  for (int base = 0; base < nblk; base += 32) {
    const int i = base + lane;
    const int x = i < nblk ? row[i] : 0;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (i < nblk) row[i] = run + incl - x;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) bin_total[b] = run;
}

// in-place exclusive scan of n ints, one block
__global__ void __launch_bounds__(1024) group_scan_kernel(int* __restrict__ v, int n) {
  __shared__ int part[1024];
  const int per = (n + 1023) / 1024, lo = min(n, int(threadIdx.x) * per), hi = min(n, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += v[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int x = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += x;
    __syncthreads();
  }
  int run = part[threadIdx.x] - sum;
  for (int i = lo; i < hi; ++i) {
    const int x = v[i];
    v[i] = run;
    run += x;
  }
}

// perm[position] = query; stable: queries of one anchor keep their order (warps of a block take turns)
__global__ void __launch_bounds__(GROUP_BLOCK)
group_scatter_kernel(const int* __restrict__ assign, int nq, int n_anchor, const int* __restrict__ blockhist,
                     const int* __restrict__ bin_start, int* __restrict__ perm) {
  extern __shared__ int gs_smem[];
  for (int b = threadIdx.x; b < n_anchor; b += blockDim.x) gs_smem[b] = bin_start[b] + blockhist[size_t(b) * gridDim.x + blockIdx.x];
  __syncthreads();
  const int i = blockIdx.x * GROUP_BLOCK + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bin = i < nq ? assign[i] : -1;
  for (int w = 0; w < GROUP_BLOCK / 32; ++w) {
    if (warp == w) {
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int base = 0;
      if (bin >= 0) base = gs_smem[bin];
      __syncwarp();
      if (bin >= 0) {
        perm[base + rank] = i;
        if (rank == 0) gs_smem[bin] = base + __popc(peers);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ threshold ladder
// piv16 -> ladder.  The sample is about 1/64 of the gallery, so the j-th smallest sampled key has about 64 j gallery rows
// below it (Gamma(j) spread).  Initial threshold = 12th smallest (768 expected, P(fewer than KPT = 128) ~ 1e-6; or the
// largest finite key when the sample was tiny), ladder rungs = 6th and 3rd smallest (384 / 192 expected): the sweep moves
// the threshold down a rung once KPT logged rows lie below it.  MMSIM_LADDER="init,mid,low" overrides the ranks
// (scripts/ladder_ablate.py: starting at the 8th saves 1.3 ms of a 28 ms sweep but sends ~23 of 100k queries to the exact
// fallback, which costs more than that; gpurun_out/ladder_ablate.log).  A separate step so
// that the gallery-sharded path can first merge the shards' lists into the list of the whole gallery's sample (all shards
// then use the same thresholds).
struct LadderRanks { int init, mid, low; };
static LadderRanks ladder_ranks() {
  LadderRanks r{12, 6, 3};
  if (const char* e = getenv("MMSIM_LADDER")) {
    int a = 0, b = 0, c = 0;
    if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a >= b && b >= c && c >= 1 && a <= NPIV) r = LadderRanks{a, b, c};
  }
  return r;
}
__global__ void make_ladder_kernel(const float* __restrict__ piv16, int rows, float* __restrict__ ladder, LadderRanks rk) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* m = piv16 + size_t(r) * NPIV;
  float top = m[0];
  for (int i = 1; i < rk.init; ++i) top = m[i] < kInf ? m[i] : top;
  const float low = m[rk.low - 1], mid = m[rk.mid - 1];
  *reinterpret_cast<float4*>(ladder + size_t(r) * 4) = make_float4(-kInf, low < top ? low : -kInf, mid < top ? mid : -kInf, top);
}

__global__ void fill_f32_kernel(float* __restrict__ p, int64_t n, float v) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// out[row] = the NPIV smallest of the union of parts[p][row] (p < nparts), ascending.
__global__ void merge_pivots_kernel(const float* __restrict__ parts, int nparts, int64_t part_stride, int rows,
                                    float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m[NPIV];
#pragma unroll
  for (int i = 0; i < NPIV; ++i) m[i] = kInf;
  for (int p = 0; p < nparts; ++p) {
    const float* src = parts + size_t(p) * part_stride + size_t(r) * NPIV;
    for (int e = 0; e < NPIV; ++e) {
      float x = src[e];
#pragma unroll
      for (int i = 0; i < NPIV; ++i) {
        const float lo = fminf(m[i], x);
        x = fmaxf(m[i], x);
        m[i] = lo;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NPIV; ++i) out[size_t(r) * NPIV + i] = m[i];
}

// ------------------------------------------------------------------------------------------------ rerank
// One warp per query: pick the KP smallest approximate keys from the sweep's candidate log, recompute those
// candidates' distances exactly, sort by (distance, index), emit top-k, certify.
constexpr int RR_WARPS = 4;
constexpr int RR_STAGE = 1024;             // logged keys per query staged in shared memory for the selection (at least)

__device__ __forceinline__ uint32_t sortable(uint32_t fbits) { return (fbits & 0x80000000u) ? ~fbits : (fbits | 0x80000000u); }
__device__ __forceinline__ float unsortable(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

template <bool VEC>
__global__ void __launch_bounds__(RR_WARPS * 32)
knn_rerank_kernel(const float* __restrict__ Q, const float* __restrict__ G, int nq, int64_t ng, int D,
                  const uint2* __restrict__ log, int logcap, const int* __restrict__ log_cnt, const float* __restrict__ log_tau,
                  int n_splits, const float* __restrict__ qnorm, const float* __restrict__ qerr,
                  const float* __restrict__ gstats, float delta_coeff, int k, int kp, int exclude_self, int64_t self_offset,
                  float* __restrict__ out_dist, int* __restrict__ out_idx, float* __restrict__ out_lb,
                  int* __restrict__ status, int* __restrict__ unc_query, float* __restrict__ unc_bound, int unc_cap,
                  const int* __restrict__ perm /* sweep position -> query (query grouping), or null: identity */,
                  int force_mod /* test hook (MMSIM_KNN_FORCE_FALLBACK=m): every m-th query counts as uncertified */,
                  int scratch_words /* per-warp scratch: the staged keys of the selection */,
                  int slice_rows, int64_t slice_stride /* gallery-shard mode, slice_rows > 0: query qo's outputs go to block
                  qo / slice_rows (blocks slice_stride words apart) at row qo % slice_rows -- the layout the all-to-all of the
                  candidate lists sends, written directly */) {
  // VEC: 8 | D <= 256 and 16-byte aligned gallery rows -- 128-bit gathers, two lanes per candidate (below)
  // kp <= KP candidates are re-ranked.  out_lb == nullptr: emit the top-k and certify locally (k <= kp).
  // out_lb != nullptr (gallery-shard mode): emit all kp re-ranked candidates (k == kp) plus the lower bound on the
  // true distance of every row of this shard that is NOT among them; the certificate is evaluated after the merge.
  extern __shared__ __align__(16) float rr_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * RR_WARPS + warp;       // position in the sweep's order: logs, norms
  if (qi >= nq) return;
  const int qo = perm ? perm[qi] : qi;                // the caller's query index: Q rows, outputs
  float* qs = rr_smem + warp * (D + 2 * KP);      // query vector, then candidate / sort scratch
  float* sk = qs + D;
  int* sv = reinterpret_cast<int*>(sk + KP);

  for (int c = lane; c < D; c += 32) qs[c] = Q[size_t(qo) * D + c];
  for (int c = lane; c < KP; c += 32) { sk[c] = kInf; sv[c] = -1; }

  // ---- gather the KP smallest logged keys (all of them when there are fewer)
  const size_t l0 = size_t(qi) * n_splits;
  int total = 0;
  bool overflow = false;
  float tau_min = kInf;
  for (int s = 0; s < n_splits; ++s) {
    const int c = log_cnt[l0 + s];
    overflow |= c > logcap;
    total += min(c, logcap);
    tau_min = fminf(tau_min, log_tau[l0 + s]);
  }
  // ---- selection: the candidates are the logged keys below a bound `ub` chosen so that between kp - 8 and kp of them
  // qualify.  Any bound works for the certificate (every unselected logged row has key >= ub), it only has to be as large
  // as the kp re-rank slots allow -- so instead of the exact kp-th smallest key (round 1: a 32-pass radix descent + two
  // gather passes, a third of this kernel's instructions) the bound comes from a bisection over the key range that stops
  // as soon as the count lands in that window: 1 min/max pass + about 8 counting passes + 1 gather pass.
  // The keys are staged in shared memory once when they fit (they do unless a log is nearly full).
  uint32_t ub = 0xffffffffu;      // exclusive bound in sortable-uint form
  uint32_t* ek = reinterpret_cast<uint32_t*>(rr_smem + RR_WARPS * (D + 2 * KP)) + size_t(warp) * scratch_words;
  const bool staged = total > kp && total <= scratch_words;
  if (total > kp) {
    uint32_t mn = 0xffffffffu, mx = 0u;
    {
      int base = 0;
      for (int s = 0; s < n_splits; ++s) {
        const uint2* ls = log + (l0 + s) * logcap;
        const int c = min(log_cnt[l0 + s], logcap);
        for (int e = lane; e < c; e += 32) {
          const uint32_t u = sortable(__ldg(&ls[e].x));
          if (staged) ek[base + e] = u;
          mn = min(mn, u);
          mx = max(mx, u);
        }
        base += c;
      }
      mn = __reduce_min_sync(0xffffffffu, mn);
      mx = __reduce_max_sync(0xffffffffu, mx);
      __syncwarp();
    }
    auto count_below = [&](uint32_t x) {
      int cnt = 0;
      if (staged) {
        for (int e = lane; e < total; e += 32) cnt += ek[e] < x ? 1 : 0;
      } else {
        for (int s = 0; s < n_splits; ++s) {
          const uint2* ls = log + (l0 + s) * logcap;
          const int c = min(log_cnt[l0 + s], logcap);
          for (int e = lane; e < c; e += 32) cnt += sortable(__ldg(&ls[e].x)) < x ? 1 : 0;
        }
      }
      return __reduce_add_sync(0xffffffffu, cnt);
    };
    // invariant: count_below(lo) <= kp < count_below(hi)   (lo = smallest key: 0 below it; hi = past the largest: all)
    uint32_t lo = mn, hi = mx == 0xffffffffu ? mx : mx + 1u;
    const int want = max(kp - 8, 1);
    while (hi - lo > 1u) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      const int c = count_below(mid);
      if (c <= kp) {
        lo = mid;
        if (c >= want) break;
      } else {
        hi = mid;
      }
    }
    ub = lo;
  }
  __syncwarp();
  {
    int filled = 0, base = 0;
    for (int s = 0; s < n_splits; ++s) {
      const uint2* ls = log + (l0 + s) * logcap;
      const int c = min(log_cnt[l0 + s], logcap);
      for (int e0 = 0; e0 < c; e0 += 32) {
        const int e = e0 + lane;
        uint32_t u = 0;
        bool take = false;
        if (e < c) {
          u = staged ? ek[base + e] : sortable(__ldg(&ls[e].x));
          take = total <= kp ? true : u < ub;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, take);
        const int pos = filled + __popc(bal & ((1u << lane) - 1u));
        if (take && pos < kp) {
          sk[pos] = unsortable(u);
          sv[pos] = int(__ldg(&ls[e].y));
        }
        filled += __popc(bal);
      }
      base += c;
    }
  }
  __syncwarp();
  // every gallery row that is not a candidate has an approximate key >= tau (never logged: >= the sweep's final
  // threshold; logged but not selected: >= the selection bound)
  const float tau = total > kp ? fminf(tau_min, unsortable(ub)) : tau_min;

  // ---- exact distances (reference arithmetic): 8 lanes per candidate, 4 candidates per round
  const int self = exclude_self ? int(self_offset + qo) : -1;
  if constexpr (VEC) {
    // 128-bit form (8 | D <= 256): TWO lanes per candidate, lane h = lane & 1 owns NumPy's accumulators r[4h .. 4h + 3] -- one
    // float4 of every 8-element step -- so a row goes through a quarter of the load instructions of the 8-lane form below
    // and 32 candidates (two per lane pair) are in flight per round.  Same operations in the same order per accumulator,
    // the same combine ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) (the second operand comes from the partner lane; fp32 addition
    // commutes), the same two-leaf split above 128 elements (exact.cuh): every bit as before.
    const int h = lane & 1, pr = lane >> 1;
    const int n2 = D <= 128 ? D : (D / 2 - (D / 2) % 8);      // first leaf [0, n2), second leaf [n2, D): multiples of 8, <= 128
    const float4* q4 = reinterpret_cast<const float4*>(qs) + h;
#pragma unroll 1
    for (int r0 = 0; r0 < KP; r0 += 32) {
      if (r0 >= kp) break;                      // slots >= kp hold the (+inf, -1) padding
      const int ia = sv[r0 + pr], ib = sv[r0 + 16 + pr];
      const bool la = ia >= 0 && ia != self, lb = ib >= 0 && ib != self;
      const float4* ga = reinterpret_cast<const float4*>(G + size_t(la ? ia : 0) * D) + h;
      const float4* gb = reinterpret_cast<const float4*>(G + size_t(lb ? ib : 0) * D) + h;
      float sa = 0.f, sb = 0.f;
      for (int e0 = 0; e0 < D;) {
        const int e1 = e0 == 0 ? n2 : D;
        const int j0 = e0 >> 2, j1 = e1 >> 2;   // in float4 units; this lane's float4s are j0, j0 + 2, ... (+ h in the pointers)
        float4 x = __ldg(ga + j0), y = __ldg(gb + j0), q = q4[j0];
        float a0 = exact_term<kSquaredEuclidean>(q.x, x.x), a1 = exact_term<kSquaredEuclidean>(q.y, x.y);
        float a2 = exact_term<kSquaredEuclidean>(q.z, x.z), a3 = exact_term<kSquaredEuclidean>(q.w, x.w);
        float b0 = exact_term<kSquaredEuclidean>(q.x, y.x), b1 = exact_term<kSquaredEuclidean>(q.y, y.y);
        float b2 = exact_term<kSquaredEuclidean>(q.z, y.z), b3 = exact_term<kSquaredEuclidean>(q.w, y.w);
#pragma unroll 4
        for (int j = j0 + 2; j < j1; j += 2) {
          x = __ldg(ga + j); y = __ldg(gb + j); q = q4[j];
          a0 = __fadd_rn(a0, exact_term<kSquaredEuclidean>(q.x, x.x));
          a1 = __fadd_rn(a1, exact_term<kSquaredEuclidean>(q.y, x.y));
          a2 = __fadd_rn(a2, exact_term<kSquaredEuclidean>(q.z, x.z));
          a3 = __fadd_rn(a3, exact_term<kSquaredEuclidean>(q.w, x.w));
          b0 = __fadd_rn(b0, exact_term<kSquaredEuclidean>(q.x, y.x));
          b1 = __fadd_rn(b1, exact_term<kSquaredEuclidean>(q.y, y.y));
          b2 = __fadd_rn(b2, exact_term<kSquaredEuclidean>(q.z, y.z));
          b3 = __fadd_rn(b3, exact_term<kSquaredEuclidean>(q.w, y.w));
        }
        float ta = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
        float tb = __fadd_rn(__fadd_rn(b0, b1), __fadd_rn(b2, b3));
        ta = __fadd_rn(ta, __shfl_xor_sync(0xffffffffu, ta, 1));
        tb = __fadd_rn(tb, __shfl_xor_sync(0xffffffffu, tb, 1));
        sa = e0 == 0 ? ta : __fadd_rn(sa, ta);
        sb = e0 == 0 ? tb : __fadd_rn(sb, tb);
        e0 = e1;
      }
      __syncwarp();
      if (h == 0) {
        sk[r0 + pr] = la ? __fsqrt_rn(sa) : kInf;
        sv[r0 + pr] = la ? ia : 0x7fffffff;
        sk[r0 + 16 + pr] = lb ? __fsqrt_rn(sb) : kInf;
        sv[r0 + 16 + pr] = lb ? ib : 0x7fffffff;
      }
    }
  } else {
    // (Round 2 tried gathering the rows with 16-byte cp.async into shared memory, eight rows per batch, double-buffered: a
    // whole 512-byte row per instruction and far more bytes in flight per warp -- and it was SLOWER, 2.9 vs 2.25 ms: the row
    // buffers halve the resident warps, and reducing from shared memory adds an LDS per element.  gpurun_out/r2_s_bench*.json)
    const int sub = lane & 7, grp = lane >> 3;
    // two independent candidates per lane group and round: twice the loads in flight (the gather is latency-bound:
    // 61% of the stall samples were on the gallery-row loads, profiles/r1_g_rerank.txt)
#pragma unroll 1
    for (int r0 = 0; r0 < KP; r0 += 8) {
      if (r0 >= kp) break;                      // slots >= kp hold the (+inf, -1) padding
      const int ia = sv[r0 + grp], ib = sv[r0 + 4 + grp];
      const bool la = ia >= 0 && ia != self, lb = ib >= 0 && ib != self;
      const float* ga = G + size_t(la ? ia : 0) * D;
      const float* gb = G + size_t(lb ? ib : 0) * D;
      float sa, sb;
      if (D <= 128 && D >= 8) {
        // interleaved form of exact_leaf_8 (same operations per candidate in the same order)
        sa = exact_term<kSquaredEuclidean>(qs[sub], ga[sub]);
        sb = exact_term<kSquaredEuclidean>(qs[sub], gb[sub]);
        int i = 8;
#pragma unroll 8
        for (; i + 8 <= D; i += 8) {
          const float q = qs[i + sub];
          sa = __fadd_rn(sa, exact_term<kSquaredEuclidean>(q, ga[i + sub]));
          sb = __fadd_rn(sb, exact_term<kSquaredEuclidean>(q, gb[i + sub]));
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          sa = __fadd_rn(sa, __shfl_xor_sync(0xffffffffu, sa, o));
          sb = __fadd_rn(sb, __shfl_xor_sync(0xffffffffu, sb, o));
        }
        for (; i < D; ++i) {
          sa = __fadd_rn(sa, exact_term<kSquaredEuclidean>(qs[i], ga[i]));
          sb = __fadd_rn(sb, exact_term<kSquaredEuclidean>(qs[i], gb[i]));
        }
      } else {
        sa = exact_reduce_8<kSquaredEuclidean>(qs, ga, D, sub);
        sb = exact_reduce_8<kSquaredEuclidean>(qs, gb, D, sub);
      }
      __syncwarp();
      if (sub == 0) {
        sk[r0 + grp] = la ? __fsqrt_rn(sa) : kInf;
        sv[r0 + grp] = la ? ia : 0x7fffffff;
        sk[r0 + 4 + grp] = lb ? __fsqrt_rn(sb) : kInf;
        sv[r0 + 4 + grp] = lb ? ib : 0x7fffffff;
      }
    }
  }
  __syncwarp();

  // ---- bitonic sort of the candidate slots (padded to a power of two >= kp) in shared memory by (distance, index)
  const int SP = kp <= 32 ? 32 : (kp <= 64 ? 64 : KP);
  // VEC: the slots are packed first -- (distance bits << 32 | index) in one 64-bit word over the sk | sv arrays -- so that a
  // compare-exchange is two 64-bit loads, one unsigned comparison and two stores instead of four loads, a two-level comparison
  // and four stores (the sort was a sixth of this kernel's instructions).  Distances are non-negative floats or +inf and
  // indices non-negative, so the unsigned order of the words is the (distance, index) order.
  unsigned long long* pk = reinterpret_cast<unsigned long long*>(sk);
  if constexpr (VEC) {
    unsigned long long e[KP / 32];
#pragma unroll
    for (int r = 0; r < KP / 32; ++r)
      e[r] = (static_cast<unsigned long long>(__float_as_uint(sk[lane + 32 * r])) << 32) | static_cast<uint32_t>(sv[lane + 32 * r]);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < KP / 32; ++r) pk[lane + 32 * r] = e[r];
    __syncwarp();
    for (int size = 2; size <= SP; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < SP / 2; t += 32) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool asc = (lo & size) == 0;
          const unsigned long long a = pk[lo], b = pk[hi];
          if ((a > b) == asc) {
            pk[lo] = b;
            pk[hi] = a;
          }
        }
        __syncwarp();
      }
    }
  } else
  for (int size = 2; size <= SP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < SP / 2; t += 32) {
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

  const size_t o_row = slice_rows > 0 ? size_t(qo / slice_rows) * size_t(slice_stride) + size_t(qo % slice_rows) * k : size_t(qo) * k;
  for (int r = lane; r < k; r += 32) {
    const float d = VEC ? __uint_as_float(uint32_t(pk[r] >> 32)) : sk[r];
    out_dist[o_row + r] = d;
    out_idx[o_row + r] = (d < kInf) ? (VEC ? int(uint32_t(pk[r])) : sv[r]) : -1;
  }

  // ---- certificate: lower bound on the true distance of every row that is not a candidate
  if (lane == 0) {
    float lb;
    if (overflow) {
      lb = -kInf;                     // the log dropped entries: the candidate set is incomplete
    } else if (!(tau < kInf)) {
      // no threshold was ever applied and every logged row is a candidate: exact by construction
      lb = ((gstats[0] < kInf) && (gstats[1] < kInf) && (qerr[qi] < kInf)) ? kInf : -kInf;
    } else {
      const float qn = qnorm[qi];
      const float delta = delta_coeff * (qn + gstats[1]);
      const float lb2 = tau + qn - delta;
      lb = sqrtf(fmaxf(lb2, 0.f)) * 0.999999f - (qerr[qi] + gstats[0]);
      if (!(lb == lb)) lb = -kInf;    // NaN inputs certify nothing
    }
    if (force_mod > 0 && qo % force_mod == 0) lb = -kInf;
    if (out_lb) {
      out_lb[slice_rows > 0 ? size_t(qo / slice_rows) * size_t(slice_stride) + size_t(qo % slice_rows) : size_t(qo)] = lb;
    } else {
      const float dk = VEC ? __uint_as_float(uint32_t(pk[k - 1] >> 32)) : sk[k - 1];
      if (!(dk < lb)) {
        const int slot = atomicAdd(&status[0], 1);     // unc_cap == nq: every query has a slot
        if (slot < unc_cap) {
          unc_query[slot] = qo;
          unc_bound[slot] = dk;
        }
      }
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

static bool sweep_pairs();

Plan make_plan(int64_t nq, int64_t ng, int64_t D, int k, int num_sms, bool host_mode) {
  Plan p{};
  p.Dp = int(align_up(size_t(D), KATOM));
  p.katoms = p.Dp / KATOM;
  p.n_qblocks = int((nq + BM - 1) / BM);
  p.n_tiles = int((ng + BN - 1) / BN);
  // MMSIM_KNN_SWEEP=q: the query-streaming sweep (knn_sweepq.cuh; opt-in: faster pipeline, slower candidate path, no net gain)
  {
    const char* e = getenv("MMSIM_KNN_SWEEP");
    p.sweepq = e && e[0] == 'q';
  }
  // DUAL sweep (two query blocks resident per CTA, half the L2 -> SM gallery bytes): K <= 128, at least two query blocks.
  // Opt-in (MMSIM_KNN_DUAL=1): results identical, but halving the L2 traffic bought 0-2% at 128-d and lost 4% at 64-d
  // (profiles/r2i_sweep_dual_mma2.txt) -- the L2 -> SM stream is not what binds the sweep.
  {
    const char* e = getenv("MMSIM_KNN_DUAL");
    p.dual = p.katoms <= 2 && p.n_qblocks >= 2 && !p.sweepq && !sweep_pairs() && e && atoi(e) == 1;
  }
  const int q_items = p.dual ? (p.n_qblocks + 1) / 2 : p.n_qblocks;     // work items per gallery split
  // gallery splits: fill the persistent grid in whole waves without making sweeps too short
  int best_s = 1;
  double best_eff = 0;
  for (int s = 1; s <= 8; ++s) {
    const int tps = (p.n_tiles + s - 1) / s;
    if (s > 1 && tps < 64) break;
    const int s_eff = (p.n_tiles + tps - 1) / tps;
    const int64_t items = int64_t(q_items) * s_eff;
    const int64_t waves = (items + num_sms - 1) / num_sms;
    // cost model: waves * tiles per sweep, minus a penalty per extra split (every split restarts its threshold ladder)
    const double eff = double(q_items) * p.n_tiles / (double(waves) * num_sms * tps) - 0.02 * (s - 1);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_s = s_eff;
    }
  }
  p.host_splits = int(std::max<int64_t>(1, std::min<int64_t>(8, p.n_tiles / 64)));
  if (host_mode) {
    // Host-buffer mode launches one sweep per split as its rows arrive, two in flight (no wave quantisation to balance):
    // splits are the unit of the copy / sweep pipeline.  8 exposes only the queries' transfer -- 29.4 / 28.2 / 28.0 / 27.1 /
    // 27.1 / 27.9 ms end to end with 3 / 4 / 6 / 8 / 12 / 16 splits against 26.1 ms device-resident (gpurun_out/host_splits*.log).
    best_s = int(std::max<int64_t>(1, std::min<int64_t>(8, p.n_tiles / 64)));
  }
  if (const char* e = getenv("MMSIM_KNN_SPLITS")) {   // experiment switch
    const int v = atoi(e);
    if (v >= 1 && v <= 16) best_s = std::min(v, p.n_tiles);
  }
  if (p.sweepq) {
    if (host_mode) p.host_splits = std::max(1, std::min(best_s, 16));
    best_s = 1;                     // one log per query; the gallery tiles are the work items
  }
  p.n_splits = best_s;
  p.tiles_per_split = (p.n_tiles + p.n_splits - 1) / p.n_splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.grid = int(std::min<int64_t>(num_sms, int64_t(q_items) * p.n_splits));
  p.unc_cap = int(nq);

  // candidate log + pivot pre-pass.  Small galleries are logged whole (threshold +inf); otherwise a systematic sample of
  // about 1/61 of the gallery rows gives each query its 16 smallest sampled keys: the 12th is the initial threshold, the
  // 6th / 3rd are the ladder the sweep tightens along (make_ladder_kernel).  The sample takes every 61st row (a prime
  // stride; the offset inside the stride changes from segment to segment), NOT contiguous tiles: with a class-sorted
  // gallery a contiguous sample holds whole classes or nothing of them, and a query whose class was sampled got a threshold
  // near its 12th neighbour (round-1 advisor finding).  The rows are gathered into a compact block the pre-pass sweeps whole.
  p.logcap = p.n_splits == 1 ? 2048 : 1024;
  p.use_pivots = ng > p.logcap ? 1 : 0;
  p.n_sample = 0; p.n_sample_tiles = 0; p.sample_div = 61; p.sample_seg = 1;
  if (p.use_pivots) {
    if (const char* e = getenv("MMSIM_PIVOT_DIV")) {   // experiment switch: sample 1/div of the gallery (with MMSIM_LADDER)
      const int v = atoi(e);
      if (v >= 8 && v <= 1024) p.sample_div = v;
    }
    p.n_sample = int(ng / p.sample_div);               // >= 33 rows (ng > 2048)
    p.sample_seg = (p.n_sample + 15) / 16;             // 16 segments, each with its own offset inside the stride
    p.n_sample_tiles = (p.n_sample + BN - 1) / BN;
    p.pivot_grid = std::min(num_sms, p.n_qblocks);
  }

  // query grouping (see "query grouping" above): anchors for enough queries to form groups of a warp's size; on by
  // default when the sweep is long enough to repay the 0.3 ms, and ALWAYS in gallery-shard mode (every shard must derive
  // the same sweep order from the queries alone: the pivot lists they exchange are indexed by sweep position)
  p.n_anchor = nq >= 65536 ? 4096 : nq >= 16384 ? 2048 : nq >= 4096 ? 1024 : 0;
  p.group_default = p.use_pivots && p.n_tiles >= 256;
  if (const char* e = getenv("MMSIM_KNN_GROUP")) {     // experiment switch: 0 = off
    if (atoi(e) == 0) p.n_anchor = 0;
  }
  p.group_blocks = int((nq + GROUP_BLOCK - 1) / GROUP_BLOCK);

  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  const size_t q_rows = size_t(p.n_qblocks) * BM;
  p.off_qh = take(q_rows * p.Dp * 2);
  p.off_gh = take(size_t(ng) * p.Dp * 2);
  p.off_gpack = take(size_t(p.n_tiles) * NPACK * 4);
  p.off_qnorm = take(q_rows * 4);
  p.off_qerr = take(q_rows * 4);
  p.off_stats = take(64);
  p.off_piv16 = take(q_rows * NPIV * 4);
  p.off_ladder = take(q_rows * 16);
  p.off_log = take(q_rows * p.n_splits * p.logcap * 8);
  p.off_log_cnt = take(q_rows * p.n_splits * 4);
  p.off_log_tau = take(q_rows * p.n_splits * 4);
  p.off_split_done = take(q_rows * p.n_splits * 4);
  p.off_tau = take(q_rows * 4);
  p.off_state = take(q_rows * 4);
  p.off_unc_query = take(size_t(p.unc_cap) * 4);
  p.off_unc_bound = take(size_t(p.unc_cap) * 4);
  p.off_fb2_list = take(size_t(p.unc_cap) * 4);
  p.off_fb_qh = take(q_rows * p.Dp * 2);
  p.off_fb_ladder = take(q_rows * 16);
  p.off_fb2_dist = take(size_t(FB2_WAVE) * FB2_CHUNKS * KP * 4);
  p.off_fb2_idx = take(size_t(FB2_WAVE) * FB2_CHUNKS * KP * 4);
  p.off_ah = take(size_t(p.n_anchor) * p.Dp * 2);
  p.off_apack = take(size_t(p.n_anchor / BN + 1) * NPACK * 4);
  p.off_aidx = take(size_t(p.n_anchor) * 4);
  p.off_assign = take(p.n_anchor ? q_rows * 4 : 0);
  p.off_perm = take(p.n_anchor ? q_rows * 4 : 0);
  p.off_ghist = take(size_t(p.n_anchor) * (p.group_blocks + 1) * 4);   // per (anchor, block) counts + per anchor totals
  // the gallery sample of the pivot pre-pass: row indices, fp16 rows + norm pack; host-buffer mode copies the fp32 rows
  // ahead of the gallery into s32
  const size_t s_rows = size_t(p.n_sample_tiles) * BN;
  p.off_sidx = take(s_rows * 4);
  p.off_s32 = take(host_mode ? s_rows * size_t(D) * 4 : 0);
  p.off_sh = take(s_rows * p.Dp * 2);
  p.off_spack = take(size_t(p.n_sample_tiles) * NPACK * 4);
  p.total_bytes = off;
  return p;
}

// Epilogue warps (NEPI / 4 per TMEM lane quarter): 16 measured best on B200 (60.6% of the dense bf16 peak vs 47.7% with 8,
// profiles/r1_b_log_epilogue_nepi8.txt, r1_c_nepi16.txt).
constexpr int NEPI_SWEEP = 16;

// MMSIM_KNN_PAIR=1: run the sweep as CTA pairs (tcgen05 cta_group::2), see the PAIR template parameter.  Opt-in: results
// are identical (tests/test_gpu_knn.py::test_cta_pair_sweep_matches), but on B200 it is not faster -- 28.1 vs 21.2 ms at
// 128-d, 43.0 vs 42.8 ms at 256-d (gpurun_out/pair_ab.log): the MMA warp and the epilogue warps hand each 256-column
// accumulator back and forth once per tile, the pair adds a cross-SM hop to that loop (remote mbarrier arrive, multicast
// commit), and at K = 128 the loop, not shared-memory bandwidth, sets the tile rate.  (Pairs with FOUR 128-column
// accumulators -- M = 256, N = 128 MMAs, three half-tiles of slack for the hand-off -- were also built and verified
// bit-identical, and were slower still: 39.9 ms.)  Read on every call so that a test can switch it.
static bool sweep_pairs() {
  const char* e = getenv("MMSIM_KNN_PAIR");
  return e && atoi(e) == 1;
}

// MMSIM_SWEEP_FLAGS (scripts/sweep_ablate.py): 2 = never log a candidate, 4 = TMEM drain only.  Ablations, wrong results.
static int sweep_ablation() {
  const char* e = getenv("MMSIM_SWEEP_FLAGS");
  return e ? atoi(e) : 0;
}

template <int KATOMS, int NEPI, int MODE, int ABL>
static int launch_tc(int grid, const CUtensorMap& tq, const CUtensorMap& tg, const SweepArgs& args, cudaStream_t stream) {
  using S = Smem<KATOMS>;
  auto kern = knn_tc_kernel<KATOMS, NEPI, MODE, ABL, false>;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES));
  kern<<<grid, 128 + NEPI * 32, S::DYN_BYTES, stream>>>(tq, tg, args);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

// the sweep as CTA pairs (clusters of 2): grid = an even number of CTAs, tg = tensor map with 128-row boxes
template <int KATOMS>
static int launch_pair(int grid, const CUtensorMap& tq, const CUtensorMap& tg, const SweepArgs& args, cudaStream_t stream) {
  using S = Smem<KATOMS>;
  auto kern = knn_tc_kernel<KATOMS, NEPI_SWEEP, MODE_SWEEP, 0, true>;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(128 + NEPI_SWEEP * 32);
  cfg.dynamicSmemBytes = S::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMSIM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tq, tg, args));
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

// MMSIM_KNN_NEPI=8 (experiment switch, round 2): the product sweep with 8 epilogue warps and four TMEM loads in flight each
static int sweep_nepi() {
  const char* e = getenv("MMSIM_KNN_NEPI");
  return e && atoi(e) == 8 ? 8 : NEPI_SWEEP;
}

template <int KATOMS>
static int launch_dual(int grid, const CUtensorMap& tq, const CUtensorMap& tg, const SweepArgs& args, cudaStream_t stream) {
  using S = Smem<KATOMS, true>;
  auto kern = knn_tc_kernel<KATOMS, NEPI_SWEEP, MODE_SWEEP, 0, false, true>;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES));
  kern<<<grid, 128 + NEPI_SWEEP * 32, S::DYN_BYTES, stream>>>(tq, tg, args);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

template <int MODE, int ABL>
static int launch_mode(int katoms, int grid, const CUtensorMap& tq, const CUtensorMap& tg, const SweepArgs& args, cudaStream_t s) {
  if constexpr (MODE == MODE_SWEEP && ABL == 0) {
    if (args.dual && katoms <= 2) return katoms == 1 ? launch_dual<1>(grid, tq, tg, args, s) : launch_dual<2>(grid, tq, tg, args, s);
    if (sweep_nepi() == 8) {
      switch (katoms) {
        case 1: return launch_tc<1, 8, MODE, ABL>(grid, tq, tg, args, s);
        case 2: return launch_tc<2, 8, MODE, ABL>(grid, tq, tg, args, s);
        case 3: return launch_tc<3, 8, MODE, ABL>(grid, tq, tg, args, s);
        default: return launch_tc<4, 8, MODE, ABL>(grid, tq, tg, args, s);
      }
    }
  }
  switch (katoms) {
    case 1: return launch_tc<1, NEPI_SWEEP, MODE, ABL>(grid, tq, tg, args, s);
    case 2: return launch_tc<2, NEPI_SWEEP, MODE, ABL>(grid, tq, tg, args, s);
    case 3: return launch_tc<3, NEPI_SWEEP, MODE, ABL>(grid, tq, tg, args, s);
    default: return launch_tc<4, NEPI_SWEEP, MODE, ABL>(grid, tq, tg, args, s);
  }
}

// query chunks of the query-streaming sweep: items = tiles x chunks should fill the persistent grid in whole waves
static void sweepq_chunks(int n_tiles, int n_qblocks, int num_sms, int* n_qchunks, int* qpc) {
  int best = 1;
  double best_eff = 0;
  for (int c = 1; c <= 16; ++c) {
    const int per = (n_qblocks + c - 1) / c;
    if (c > 1 && per < 24) break;                       // an item should amortise loading its gallery tile
    const int c_eff = (n_qblocks + per - 1) / per;
    const int64_t items = int64_t(n_tiles) * c_eff;
    const int64_t waves = (items + num_sms - 1) / num_sms;
    const double eff = double(n_tiles) * n_qblocks / (double(waves) * num_sms * per) - 0.005 * (c_eff - 1);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = c_eff;
    }
  }
  if (const char* e = getenv("MMSIM_KNN_QCHUNKS")) {   // experiment switch
    const int v = atoi(e);
    if (v >= 1 && v <= n_qblocks) best = v;
  }
  *qpc = (n_qblocks + best - 1) / best;
  *n_qchunks = (n_qblocks + *qpc - 1) / *qpc;
}

template <int KATOMS>
static int launch_sweepq_k(int grid, const CUtensorMap& tq, const CUtensorMap& tg, const SweepQArgs& args, cudaStream_t stream) {
  using S = SmemQ<KATOMS>;
  auto kern = knn_sweepq_kernel<KATOMS, NEPI_SWEEP>;
  MMSIM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES));
  kern<<<grid, 128 + NEPI_SWEEP * 32, S::DYN_BYTES, stream>>>(tq, tg, args);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}
static int launch_sweepq(int katoms, int grid, const CUtensorMap& tq, const CUtensorMap& tg, const SweepQArgs& args, cudaStream_t s) {
  switch (katoms) {
    case 1: return launch_sweepq_k<1>(grid, tq, tg, args, s);
    case 2: return launch_sweepq_k<2>(grid, tq, tg, args, s);
    case 3: return launch_sweepq_k<3>(grid, tq, tg, args, s);
    default: return launch_sweepq_k<4>(grid, tq, tg, args, s);
  }
}

// Host-buffer mode: three internal streams per device (copies, gallery prep, every second sweep) and the events that order
// them against the caller's stream.  Created once, never destroyed; calls on one device enqueue one at a time.
constexpr int kMaxSplits = 16;
struct PipeStreams {
  cudaStream_t copy = nullptr, prep = nullptr, sweep2 = nullptr;
  cudaEvent_t start = nullptr, q_in = nullptr, sample_in = nullptr, ladder = nullptr, sweep2_done = nullptr;
  cudaEvent_t chunk_in[kMaxSplits] = {}, chunk_ready[kMaxSplits] = {};
  bool ready = false;
};
static std::mutex g_pipe_mutex;
static PipeStreams g_pipe[64];

static int pipe_streams(int dev, PipeStreams** out) {
  MMSIM_REQUIRE(dev >= 0 && dev < 64, MMSIM_ERR_ARG, "knn_host: device ordinal %d out of range", dev);
  PipeStreams& ps = g_pipe[dev];
  if (!ps.ready) {
    MMSIM_CUDA_CHECK(cudaStreamCreateWithFlags(&ps.copy, cudaStreamNonBlocking));
    MMSIM_CUDA_CHECK(cudaStreamCreateWithFlags(&ps.prep, cudaStreamNonBlocking));
    MMSIM_CUDA_CHECK(cudaStreamCreateWithFlags(&ps.sweep2, cudaStreamNonBlocking));
    cudaEvent_t* single[] = {&ps.start, &ps.q_in, &ps.sample_in, &ps.ladder, &ps.sweep2_done};
    for (cudaEvent_t* e : single) MMSIM_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int i = 0; i < kMaxSplits; ++i) {
      MMSIM_CUDA_CHECK(cudaEventCreateWithFlags(&ps.chunk_in[i], cudaEventDisableTiming));
      MMSIM_CUDA_CHECK(cudaEventCreateWithFlags(&ps.chunk_ready[i], cudaEventDisableTiming));
    }
    ps.ready = true;
  }
  *out = &ps;
  return MMSIM_OK;
}

__global__ void sample_index_kernel(int* __restrict__ idx, int n_sample, int div, int seg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_sample) idx[j] = int(sample_row(j, div, seg));
}

// delta bounds the fp32 accumulation error of key = |g|^2 - 2 q.g: (Dp + 8) roundings of relative size 2^-24, on terms
// bounded by (|q|^2 + |g|^2), 4x safety (which also covers the one-ulp slack of the sweep's reordered chunk test
// m < tau - min|g|^2; tests/test_gpu_certificate.py measures the real key error against it).
static int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

static float delta_coeff_of(int Dp) { return 4.0f * float(Dp + 8) * 5.9604645e-8f; }

// Exact fallback (knn_fallback.cuh) for the queries listed in L (count on the device).  first_wave: tier 1 + the first
// tier-2 wave (part of every call); otherwise one further tier-2 wave (mmsim_knn_finish_f32).
static int run_fallback(const Plan& p, uint8_t* w, const float* Q, const float* G, int64_t ng, int D, int k, int exclude_self,
                        int64_t self_offset, const FbLists& L, const FbOut& out, int num_sms, bool first_wave, cudaStream_t stream) {
  const __half* gh = reinterpret_cast<const __half*>(w + p.off_gh);
  const float* gpack = reinterpret_cast<const float*>(w + p.off_gpack);
  const float* gstats = reinterpret_cast<const float*>(w + p.off_stats);
  __half* fb_qh = reinterpret_cast<__half*>(w + p.off_fb_qh);
  float* fb_ladder = reinterpret_cast<float*>(w + p.off_fb_ladder);
  uint2* log = reinterpret_cast<uint2*>(w + p.off_log);
  int* log_cnt = reinterpret_cast<int*>(w + p.off_log_cnt);
  float* log_tau = reinterpret_cast<float*>(w + p.off_log_tau);
  float* part_dist = reinterpret_cast<float*>(w + p.off_fb2_dist);
  int* part_idx = reinterpret_cast<int*>(w + p.off_fb2_idx);
  const size_t q_rows = size_t(p.n_qblocks) * BM;
  const long long budget = (long long)(q_rows) * p.n_splits;     // (query row, split) slots of the main plan's log
  if (first_wave) {
    fb_prepare_kernel<<<2 * num_sms, FB_THREADS, 0, stream>>>(Q, D, p.Dp, L, gstats, delta_coeff_of(p.Dp), fb_qh, fb_ladder,
                                                              env_int("MMSIM_KNN_FORCE_TIER2"));
    MMSIM_CUDA_CHECK(::mmsim::launched());
    CUtensorMap tq, tg;
    int rc = make_tmap(&tq, fb_qh, int64_t(q_rows), p.Dp, BM);
    if (rc) return rc;
    rc = make_tmap(&tg, gh, ng, p.Dp, BN);
    if (rc) return rc;
    SweepArgs a{};
    a.gpack = gpack;
    a.n_tiles = p.n_tiles;
    a.use_pivots = 1;
    a.log = log; a.logcap = p.logcap; a.log_cnt = log_cnt; a.log_tau = log_tau;
    a.ladder = fb_ladder;
    a.dyn_count = L.count; a.dyn_cap = L.cap; a.dyn_ctas = num_sms; a.dyn_budget = budget;
    rc = launch_mode<MODE_RESWEEP, 0>(p.katoms, num_sms, tq, tg, a, stream);
    if (rc) return rc;
    fb_select_kernel<<<4 * num_sms, FB_THREADS, size_t(D) * 4, stream>>>(Q, G, D, log, p.logcap, log_cnt, p.n_tiles, num_sms, budget, L,
                                                                         k, exclude_self, self_offset, out);
    MMSIM_CUDA_CHECK(::mmsim::launched());
  }
  fb2_scan_kernel<<<dim3(FB2_CHUNKS, FB2_WAVE), FB_THREADS, size_t(D) * 4, stream>>>(Q, G, ng, D, L, k, exclude_self, self_offset,
                                                                                      part_dist, part_idx);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  fb2_merge_kernel<<<FB2_WAVE, FB_THREADS, 0, stream>>>(L, k, part_dist, part_idx, out);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  fb2_advance_kernel<<<1, 1, 0, stream>>>(L.status, L.cap);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

int shard_fallback(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self, int64_t self_offset,
                   const float* flag, int cap, float* out_dist, int* out_idx, int* out_query, int* status, void* ws,
                   size_t ws_bytes, cudaStream_t stream, bool host_layout) {
  MMSIM_REQUIRE(Q && G && flag && out_dist && out_idx && out_query && status && ws, MMSIM_ERR_ARG, "knn_shard_fallback: null pointer argument");
  MMSIM_REQUIRE(nq > 0 && ng > 0 && D > 0 && D <= 4 * KATOM && k >= 1 && k <= KP && cap >= 1, MMSIM_ERR_ARG,
                "knn_shard_fallback: bad sizes");
  int dev = 0, num_sms = 0;
  MMSIM_CUDA_CHECK(cudaGetDevice(&dev));
  MMSIM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const Plan p = make_plan(nq, ng, D, k, num_sms, host_layout);
  MMSIM_REQUIRE(ws_bytes >= p.total_bytes, MMSIM_ERR_WORKSPACE, "knn_shard_fallback: workspace too small (%zu < %zu)", ws_bytes,
                p.total_bytes);
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* bound = reinterpret_cast<float*>(w + p.off_unc_bound);
  fb_list_from_flags_kernel<<<1, 1024, 0, stream>>>(flag, int(nq), cap, out_query, bound, status);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  FbLists L{status, out_query, bound, reinterpret_cast<int*>(w + p.off_fb2_list), status, cap};
  const FbOut out{out_dist, out_idx, 1};
  return run_fallback(p, w, Q, G, ng, int(D), k, exclude_self, self_offset, L, out, num_sms, true, stream);
}

int run(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self, int64_t self_offset,
        float* out_dist, int* out_idx, int* status, void* ws, size_t ws_bytes, cudaStream_t stream, int phases, int shard_kp,
        float* out_lb, const HostPipe* host, int64_t slice_rows, int64_t slice_stride) {
  MMSIM_REQUIRE(slice_rows == 0 || (shard_kp != 0 && slice_rows > 0 && slice_stride >= slice_rows * (2 * shard_kp + 1)), MMSIM_ERR_ARG,
                "knn: the slice layout belongs to gallery-shard mode (stride >= slice_rows * (2 kp + 1))");
  // host-buffer mode: the gallery always comes from the host; the queries too (unsharded call: all phases at once), or they
  // are on the device already (gallery-shard mode: phases as the sharded protocol needs them, kPhasePrepG first)
  MMSIM_REQUIRE(!host || (host->g_host && (host->q_host ? (phases == kPhaseAll && shard_kp == 0) : shard_kp != 0)), MMSIM_ERR_ARG,
                "knn_host: host-buffer mode needs the host gallery, and either host queries (all phases, unsharded) or shard mode");
  MMSIM_REQUIRE(shard_kp == 0 || (out_lb && shard_kp >= 1 && shard_kp <= KP), MMSIM_ERR_ARG,
                "knn: shard mode needs out_lb and 1 <= kp <= %d", KP);
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
  const Plan p = make_plan(nq, ng, D, k, num_sms, host != nullptr);
  MMSIM_REQUIRE(ws_bytes >= p.total_bytes, MMSIM_ERR_WORKSPACE, "knn: workspace too small (%zu < %zu)", ws_bytes, p.total_bytes);

  uint8_t* w = static_cast<uint8_t*>(ws);
  __half* qh = reinterpret_cast<__half*>(w + p.off_qh);
  __half* gh = reinterpret_cast<__half*>(w + p.off_gh);
  float* gpack = reinterpret_cast<float*>(w + p.off_gpack);
  float* qnorm = reinterpret_cast<float*>(w + p.off_qnorm);
  float* qerr = reinterpret_cast<float*>(w + p.off_qerr);
  float* gstats = reinterpret_cast<float*>(w + p.off_stats);
  float* piv16 = reinterpret_cast<float*>(w + p.off_piv16);
  float* ladder = reinterpret_cast<float*>(w + p.off_ladder);
  uint2* log = reinterpret_cast<uint2*>(w + p.off_log);
  int* log_cnt = reinterpret_cast<int*>(w + p.off_log_cnt);
  float* log_tau = reinterpret_cast<float*>(w + p.off_log_tau);
  int* split_done = reinterpret_cast<int*>(w + p.off_split_done);
  int* unc_query = reinterpret_cast<int*>(w + p.off_unc_query);
  float* unc_bound = reinterpret_cast<float*>(w + p.off_unc_bound);
  __half* ah = reinterpret_cast<__half*>(w + p.off_ah);
  float* apack = reinterpret_cast<float*>(w + p.off_apack);
  int* aidx = reinterpret_cast<int*>(w + p.off_aidx);
  int* assign = reinterpret_cast<int*>(w + p.off_assign);
  const bool group = p.n_anchor != 0 && (shard_kp != 0 || p.group_default);
  int* perm = group ? reinterpret_cast<int*>(w + p.off_perm) : nullptr;
  int* ghist = reinterpret_cast<int*>(w + p.off_ghist);
  int* sidx = reinterpret_cast<int*>(w + p.off_sidx);
  float* s32 = reinterpret_cast<float*>(w + p.off_s32);
  __half* sh = reinterpret_cast<__half*>(w + p.off_sh);
  float* spack = reinterpret_cast<float*>(w + p.off_spack);

  PipeStreams* ps = nullptr;
  std::unique_lock<std::mutex> pipe_lock;
  if (host) {
    MMSIM_REQUIRE(p.n_splits <= kMaxSplits, MMSIM_ERR_UNSUPPORTED, "knn_host: more than %d gallery splits", kMaxSplits);
    pipe_lock = std::unique_lock<std::mutex>(g_pipe_mutex);
    if (int rc0 = pipe_streams(dev, &ps)) return rc0;
  }
  const bool vec = D % 4 == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0;
  auto prep = vec ? prep_rows_kernel<true> : prep_rows_kernel<false>;
  const int warps_per_block = PREP_THREADS / 32;
  const int64_t cap = int64_t(num_sms) * 16;     // grid-stride: a few resident waves, one atomic pair per block

  if (phases & kPhaseRerank) {
    MMSIM_CUDA_CHECK(cudaMemsetAsync(status, 0, 8 * sizeof(int), stream));
  }

  // 1. operand copies: fp16 rows, norm pack (+ per-8 / per-32 column minima), rounding-error norms
  // (kPhasePrepG / kPhasePrepQ: the gallery half and the query half on their own -- the sharded end-to-end path prepares the
  // queries while the gallery shard is still on its way from the host)
  const bool prep_g = (phases & (kPhasePrep | kPhasePrepG)) != 0, prep_q = (phases & (kPhasePrep | kPhasePrepQ)) != 0;
  if (prep_g) {
    MMSIM_RANGE("knn: gallery operand copies");
    MMSIM_CUDA_CHECK(cudaMemsetAsync(gstats, 0, 64, stream));
    if (host) {
      // host-buffer mode: EVERY host -> device copy of the call is queued here, on the copy stream, in the order the
      // kernels need the data: the queries (unsharded call), the gallery sample of the pivot pre-pass (strided 2-D copies:
      // the pre-pass, and with it the first sweep, does not wait for the gallery), then the gallery split by split, each
      // split converted on the prep stream as it lands.  The phases below only wait for the events.  Everything queued
      // earlier on the caller's stream -- a previous call still reading the staging buffers or this workspace -- precedes
      // the copies.
      MMSIM_CUDA_CHECK(cudaEventRecord(ps->start, stream));
      MMSIM_CUDA_CHECK(cudaStreamWaitEvent(ps->copy, ps->start, 0));
      if (host->q_host) {
        MMSIM_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(Q), host->q_host, size_t(nq) * D * 4, cudaMemcpyHostToDevice, ps->copy));
        MMSIM_CUDA_CHECK(cudaEventRecord(ps->q_in, ps->copy));
        MMSIM_CUDA_CHECK(cudaStreamWaitEvent(stream, ps->q_in, 0));
      }
      if (p.use_pivots) {
        for (int64_t j0 = 0; j0 < p.n_sample; j0 += p.sample_seg) {
          const int64_t j1 = std::min<int64_t>(p.n_sample, j0 + p.sample_seg);
          MMSIM_CUDA_CHECK(cudaMemcpy2DAsync(s32 + size_t(j0) * D, size_t(D) * 4, host->g_host + size_t(sample_row(j0, p.sample_div, p.sample_seg)) * D,
                                             size_t(p.sample_div) * D * 4, size_t(D) * 4, size_t(j1 - j0), cudaMemcpyHostToDevice,
                                             ps->copy));
        }
        MMSIM_CUDA_CHECK(cudaEventRecord(ps->sample_in, ps->copy));
      }
      const int n_chunks = p.sweepq ? p.host_splits : p.n_splits;
      const int tpc = p.sweepq ? (p.n_tiles + n_chunks - 1) / n_chunks : p.tiles_per_split;
      for (int c = 0; c < n_chunks; ++c) {
        const int64_t t0 = int64_t(c) * tpc, t1 = std::min<int64_t>(p.n_tiles, t0 + tpc);
        if (t0 >= t1) break;
        const int64_t r0 = t0 * BN, r1 = std::min<int64_t>(ng, t1 * BN), r_pad = t1 * BN - r0;
        MMSIM_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(G) + size_t(r0) * D, host->g_host + size_t(r0) * D,
                                         size_t(r1 - r0) * D * 4, cudaMemcpyHostToDevice, ps->copy));
        MMSIM_CUDA_CHECK(cudaEventRecord(ps->chunk_in[c], ps->copy));
        MMSIM_CUDA_CHECK(cudaStreamWaitEvent(ps->prep, ps->chunk_in[c], 0));
        const unsigned gb = unsigned(std::min<int64_t>((r_pad + warps_per_block - 1) / warps_per_block, cap));
        prep<<<gb, PREP_THREADS, 0, ps->prep>>>(G + size_t(r0) * D, r1 - r0, r_pad, int(D), p.Dp, 1.0f, gh + size_t(r0) * p.Dp,
                                                gpack + size_t(t0) * NPACK, nullptr, 1,
                                                reinterpret_cast<unsigned int*>(gstats), nullptr);
        MMSIM_CUDA_CHECK(::mmsim::launched());
        pack_min_kernel<<<unsigned(t1 - t0), BN, 0, ps->prep>>>(gpack + size_t(t0) * NPACK);
        MMSIM_CUDA_CHECK(::mmsim::launched());
        MMSIM_CUDA_CHECK(cudaEventRecord(ps->chunk_ready[c], ps->prep));
      }
    } else {
      const int64_t g_pad = int64_t(p.n_tiles) * BN;
      const unsigned gb = unsigned(std::min<int64_t>((g_pad + warps_per_block - 1) / warps_per_block, cap));
      prep<<<gb, PREP_THREADS, 0, stream>>>(G, ng, g_pad, int(D), p.Dp, 1.0f, gh, gpack, nullptr, 1,
                                            reinterpret_cast<unsigned int*>(gstats), nullptr);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      pack_min_kernel<<<unsigned(p.n_tiles), BN, 0, stream>>>(gpack);
      MMSIM_CUDA_CHECK(::mmsim::launched());
    }
  }
  if (prep_q) {
    MMSIM_RANGE("knn: query operand copies + grouping");
    const unsigned qb = unsigned(std::min<int64_t>((nq + warps_per_block - 1) / warps_per_block, cap));
    prep<<<qb, PREP_THREADS, 0, stream>>>(Q, nq, nq, int(D), p.Dp, -2.0f, qh, qnorm, qerr, 0, nullptr, nullptr);
    MMSIM_CUDA_CHECK(::mmsim::launched());
    if (group) {
      // query grouping: anchors = evenly spaced query rows -> nearest anchor of every query (tensor cores) -> stable
      // counting sort -> operand copies of the queries again, in the sweep's order
      const int a_tiles = p.n_anchor / BN;
      anchor_index_kernel<<<(p.n_anchor + 255) / 256, 256, 0, stream>>>(aidx, p.n_anchor, nq);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      prep<<<unsigned((p.n_anchor + warps_per_block - 1) / warps_per_block), PREP_THREADS, 0, stream>>>(
          Q, p.n_anchor, p.n_anchor, int(D), p.Dp, 1.0f, ah, apack, nullptr, 1, nullptr, aidx);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      // the assign pass reads the per-32-column minima of the pack for its chunk-level early out.  (Round 1 left them
      // uninitialised: the assignment then depended on whatever the workspace held, so two shards with their own
      // workspaces could derive DIFFERENT sweep orders from the same queries and exchange misaligned pivot lists --
      // the uncertified query of tests/test_gpu_merge.py::test_reduced_protocol_with_grouped_queries on a fresh box.)
      pack_min_kernel<<<unsigned(a_tiles), BN, 0, stream>>>(apack);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      CUtensorMap tq0, ta;
      int rc0 = make_tmap(&tq0, qh, nq, p.Dp, BM);
      if (rc0) return rc0;
      rc0 = make_tmap(&ta, ah, p.n_anchor, p.Dp, BN);
      if (rc0) return rc0;
      SweepArgs aa{};
      aa.gpack = apack;
      aa.nq = int(nq); aa.n_qblocks = p.n_qblocks; aa.n_tiles = a_tiles;
      aa.n_splits = 1; aa.tiles_per_split = a_tiles;
      aa.n_sample_tiles = a_tiles;
      aa.assign = assign;
      rc0 = launch_mode<MODE_ASSIGN, 0>(p.katoms, std::min(num_sms, p.n_qblocks), tq0, ta, aa, stream);
      if (rc0) return rc0;
      const size_t hsm = size_t(p.n_anchor) * 4;
      group_hist_kernel<<<p.group_blocks, GROUP_BLOCK, hsm, stream>>>(assign, int(nq), p.n_anchor, ghist);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      int* bin_start = ghist + size_t(p.n_anchor) * p.group_blocks;
      group_binscan_kernel<<<(p.n_anchor * 32 + 255) / 256, 256, 0, stream>>>(ghist, p.group_blocks, p.n_anchor, bin_start);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      group_scan_kernel<<<1, 1024, 0, stream>>>(bin_start, p.n_anchor);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      group_scatter_kernel<<<p.group_blocks, GROUP_BLOCK, hsm, stream>>>(assign, int(nq), p.n_anchor, ghist, bin_start, perm);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      prep<<<qb, PREP_THREADS, 0, stream>>>(Q, nq, nq, int(D), p.Dp, -2.0f, qh, qnorm, qerr, 0, nullptr, perm);
      MMSIM_CUDA_CHECK(::mmsim::launched());
    }
  }

  // 2. pivot pre-pass (sampled gallery tiles) + fused distance / candidate sweep
  int rc = MMSIM_OK;
  if (phases & (kPhaseTensor | kPhasePivot | kPhaseLadder)) {
    CUtensorMap tq, tg;
    rc = make_tmap(&tq, qh, nq, p.Dp, BM);
    if (rc) return rc;
    rc = make_tmap(&tg, gh, ng, p.Dp, BN);
    if (rc) return rc;
    SweepArgs args{};
    args.gpack = gpack;
    args.nq = int(nq); args.n_qblocks = p.n_qblocks; args.n_tiles = p.n_tiles;
    args.dual = (p.dual && sweep_ablation() != 2 && sweep_ablation() != 4) ? 1 : 0;   // (the two ablation instantiations are single-block kernels)
    args.n_splits = p.n_splits; args.tiles_per_split = p.tiles_per_split;
    args.use_pivots = p.use_pivots;
    args.log = log; args.logcap = p.logcap; args.log_cnt = log_cnt; args.log_tau = log_tau; args.split_done = split_done;
    args.n_sample_tiles = p.n_sample_tiles;
    args.piv16 = piv16; args.ladder = ladder;
    if ((phases & kPhasePivot) && !p.use_pivots) {   // shard small enough to be logged whole: an empty (+inf) pivot list
      const int64_t n = int64_t(p.n_qblocks) * BM * NPIV;
      fill_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(piv16, n, kInf);
      MMSIM_CUDA_CHECK(::mmsim::launched());
    }
    if ((phases & kPhasePivot) && p.use_pivots) {
      MMSIM_RANGE("knn: pivot pre-pass");
      // The pre-pass sweeps a compact block holding the sampled gallery rows (sample_row(): every sample_div-th row).
      // Device-resident call: the operand-copy kernel gathers them from G.  Host-buffer call: they come over first, as
      // one strided 2-D copy per segment, so the pre-pass (and with it the first sweep) does not wait for the gallery.
      // Same rows either way, so both calls derive the same pivot lists.
      const int64_t s_pad = int64_t(p.n_sample_tiles) * BN;
      const unsigned sb = unsigned(std::min<int64_t>((s_pad + warps_per_block - 1) / warps_per_block, cap));
      if (host) {
        MMSIM_CUDA_CHECK(cudaStreamWaitEvent(stream, ps->sample_in, 0));       // (copied by the kPhasePrep block)
        prep<<<sb, PREP_THREADS, 0, stream>>>(s32, p.n_sample, s_pad, int(D), p.Dp, 1.0f, sh, spack, nullptr, 1, nullptr, nullptr);
      } else {
        sample_index_kernel<<<unsigned((p.n_sample + 255) / 256), 256, 0, stream>>>(sidx, p.n_sample, p.sample_div, p.sample_seg);
        MMSIM_CUDA_CHECK(::mmsim::launched());
        prep<<<sb, PREP_THREADS, 0, stream>>>(G, p.n_sample, s_pad, int(D), p.Dp, 1.0f, sh, spack, nullptr, 1, nullptr, sidx);
      }
      MMSIM_CUDA_CHECK(::mmsim::launched());
      pack_min_kernel<<<unsigned(p.n_sample_tiles), BN, 0, stream>>>(spack);
      MMSIM_CUDA_CHECK(::mmsim::launched());
      CUtensorMap ts;
      rc = make_tmap(&ts, sh, p.n_sample, p.Dp, BN);
      if (rc) return rc;
      SweepArgs sa = args;
      sa.gpack = spack;
      sa.n_tiles = p.n_sample_tiles;           // every tile of the compact block is swept whole
      rc = launch_mode<MODE_PIVOT, 0>(p.katoms, p.pivot_grid, tq, ts, sa, stream);
      if (rc) return rc;
    }
    if ((phases & kPhaseLadder) && p.use_pivots) {
      const int rows = p.n_qblocks * BM;
      make_ladder_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(piv16, rows, ladder, ladder_ranks());
      MMSIM_CUDA_CHECK(::mmsim::launched());
    }
    if (phases & kPhaseTensor) {
      MMSIM_RANGE("knn: tcgen05 sweep");
      if (p.n_splits > 1)
        MMSIM_CUDA_CHECK(cudaMemsetAsync(split_done, 0, size_t(p.n_qblocks) * BM * p.n_splits * 4, stream));
      const int abl = sweep_ablation();
      args.close_rows = abl == 8;
      const int q_rows = p.n_qblocks * BM;
      float* tau = reinterpret_cast<float*>(w + p.off_tau);
      unsigned int* state = reinterpret_cast<unsigned int*>(w + p.off_state);
      SweepQArgs qa{};
      if (p.sweepq) {
        // query-streaming sweep (knn_sweepq.cuh): thresholds and cursors are global words, initialised here
        sweepq_init_kernel<<<(q_rows + 255) / 256, 256, 0, stream>>>(ladder, p.use_pivots, int(nq), q_rows, tau, state, abl == 8);
        MMSIM_CUDA_CHECK(::mmsim::launched());
        qa.gpack = gpack; qa.nq = int(nq); qa.n_qblocks = p.n_qblocks;
        qa.tau = tau; qa.state = state; qa.ladder = ladder; qa.use_pivots = p.use_pivots;
        qa.log = log; qa.logcap = p.logcap; qa.drop = (abl == 16 || abl == 32) ? abl : 0;
      }
      // one launch of the query-streaming sweep over gallery tiles [t0, t1)
      auto launch_q = [&](int t0, int t1, cudaStream_t x) -> int {
        SweepQArgs sa = qa;
        sa.tile_begin = t0; sa.tile_end = t1;
        sweepq_chunks(t1 - t0, p.n_qblocks, num_sms, &sa.n_qchunks, &sa.qpc);
        return launch_sweepq(p.katoms, std::min(num_sms, (t1 - t0) * sa.n_qchunks), tq, tg, sa, x);
      };
      if (host) {
        // host-buffer mode: one launch per gallery split.  Split c's rows were queued for copy (copy stream) and conversion
        // (prep stream) by the kPhasePrep block; it is swept while the copy of split c + 1 is in flight.  The sweeps alternate
        // between the caller's stream and a second one so that the CTAs of the next sweep fill the SMs the last wave of the
        // previous one leaves idle.
        MMSIM_CUDA_CHECK(cudaEventRecord(ps->ladder, stream));
        bool used2 = false;
        const char* e1 = getenv("MMSIM_HOST_ONE_STREAM");   // experiment switch: every sweep on the caller's stream
        const bool one_stream = e1 && atoi(e1) == 1;
        const int n_chunks = p.sweepq ? p.host_splits : p.n_splits;
        const int tpc = p.sweepq ? (p.n_tiles + n_chunks - 1) / n_chunks : p.tiles_per_split;
        for (int c = 0; c < n_chunks; ++c) {
          const int64_t t0 = int64_t(c) * tpc, t1 = std::min<int64_t>(p.n_tiles, t0 + tpc);
          if (t0 >= t1) break;
          cudaStream_t x = (c & 1) && !one_stream ? ps->sweep2 : stream;
          if (x != stream && !used2) {
            MMSIM_CUDA_CHECK(cudaStreamWaitEvent(x, ps->ladder, 0));
            used2 = true;
          }
          MMSIM_CUDA_CHECK(cudaStreamWaitEvent(x, ps->chunk_ready[c], 0));
          if (p.sweepq) {
            rc = launch_q(int(t0), int(t1), x);
          } else {
            SweepArgs sa = args;
            const int q_items = sa.dual ? (p.n_qblocks + 1) / 2 : p.n_qblocks;
            sa.item_begin = c * q_items;
            sa.item_end = (c + 1) * q_items;
            rc = launch_mode<MODE_SWEEP, 0>(p.katoms, std::min(num_sms, q_items), tq, tg, sa, x);
          }
          if (rc) return rc;
        }
        if (used2) {
          MMSIM_CUDA_CHECK(cudaEventRecord(ps->sweep2_done, ps->sweep2));
          MMSIM_CUDA_CHECK(cudaStreamWaitEvent(stream, ps->sweep2_done, 0));
        }
      } else if (p.sweepq) {
        rc = launch_q(0, p.n_tiles, stream);
        if (rc) return rc;
      } else if (sweep_pairs() && (abl == 0 || abl == 8)) {
        CUtensorMap tgh;
        rc = make_tmap(&tgh, gh, ng, p.Dp, BN / 2);
        if (rc) return rc;
        const int grid = std::min(num_sms & ~1, 2 * ((p.n_qblocks + 1) / 2) * p.n_splits);
        switch (p.katoms) {
          case 1: rc = launch_pair<1>(grid, tq, tgh, args, stream); break;
          case 2: rc = launch_pair<2>(grid, tq, tgh, args, stream); break;
          case 3: rc = launch_pair<3>(grid, tq, tgh, args, stream); break;
          default: rc = launch_pair<4>(grid, tq, tgh, args, stream); break;
        }
        if (rc) return rc;
      } else
      rc = abl == 2   ? launch_mode<MODE_SWEEP, 2>(p.katoms, p.grid, tq, tg, args, stream)
           : abl == 4 ? launch_mode<MODE_SWEEP, 4>(p.katoms, p.grid, tq, tg, args, stream)
                      : launch_mode<MODE_SWEEP, 0>(p.katoms, p.grid, tq, tg, args, stream);
      if (rc) return rc;
      if (p.sweepq) {
        sweepq_finish_kernel<<<(q_rows + 255) / 256, 256, 0, stream>>>(state, tau, q_rows, log_cnt, log_tau);
        MMSIM_CUDA_CHECK(::mmsim::launched());
      }
    }
  }

  // 3. candidate selection + exact re-rank + certificate.  delta bounds the fp32 accumulation error of
  //    key = |g|^2 - 2 q.g: (Dp + 8) roundings of relative size 2^-24, on terms bounded by (|q|^2 + |g|^2), 4x safety.
  if (phases & kPhaseRerank) {
    const float delta_coeff = delta_coeff_of(p.Dp);
    MMSIM_RANGE("knn: select + exact re-rank + certificate");
    const int blocks = int((nq + RR_WARPS - 1) / RR_WARPS);
    const int scratch_words = RR_STAGE;               // per-warp scratch: the staged keys of the selection
    const size_t smem = size_t(RR_WARPS) * (size_t(D) + 2 * KP + scratch_words) * 4;
    // MMSIM_RR_VEC=1 (opt-in, experiment): 128-bit gathers + packed 64-bit sort where the width and the alignment allow.
    // Bit-identical (tests/test_gpu_zz_graphed.py) and SLOWER on B200 -- re-rank phase 3.12 against 2.35 ms at 100k x 1M x 128
    // (profiles/r2x_rerank_vec.txt): 16 rows per load instruction instead of 4 and 36 instead of 40 resident warps lose more
    // than the 30% fewer instructions win, like the cp.async row gather before it.  The default stays the 8-lane form.
    const char* rr_env = getenv("MMSIM_RR_VEC");
    const bool rr_vec = (D % 8 == 0 && D >= 8 && D <= 256 && (reinterpret_cast<uintptr_t>(G) & 15) == 0 &&
                        rr_env && rr_env[0] == '1');
    auto rerank = rr_vec ? knn_rerank_kernel<true> : knn_rerank_kernel<false>;
    MMSIM_CUDA_CHECK(cudaFuncSetAttribute(rerank, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    rerank<<<blocks, RR_WARPS * 32, smem, stream>>>(Q, G, int(nq), ng, int(D), log, p.logcap, log_cnt, log_tau,
                                                               p.n_splits, qnorm, qerr, gstats, delta_coeff, shard_kp ? shard_kp : k,
                                                               shard_kp ? shard_kp : KP, exclude_self, self_offset, out_dist,
                                                               out_idx, shard_kp ? out_lb : nullptr, status, unc_query,
                                                               unc_bound, p.unc_cap, perm, env_int("MMSIM_KNN_FORCE_FALLBACK"),
                                                               scratch_words, int(slice_rows), slice_stride);
    MMSIM_CUDA_CHECK(::mmsim::launched());
  }

  // 4. exact fallback for the uncertified queries (knn_fallback.cuh; the count lives on the device: when it is zero the
  //    kernels below find nothing to do)
  if (((phases & kPhaseFallback) && !shard_kp) || (phases & kPhaseFinish)) {
    MMSIM_RANGE("knn: exact fallback of uncertified queries");
    if ((phases & kPhaseFallback) && !(phases & kPhaseRerank))      // phase-by-phase timing: restart the tier-2 queue
      MMSIM_CUDA_CHECK(cudaMemsetAsync(status + 1, 0, 2 * sizeof(int), stream));
    FbLists L{status, unc_query, unc_bound, reinterpret_cast<int*>(w + p.off_fb2_list), status, p.unc_cap};
    const FbOut out{out_dist, out_idx, 0};
    rc = run_fallback(p, w, Q, G, ng, int(D), k, exclude_self, self_offset, L, out, num_sms, (phases & kPhaseFallback) != 0, stream);
    if (rc) return rc;
  }
  return MMSIM_OK;
}

int merge_pivots(const float* parts, int nparts, int64_t part_stride, int64_t rows, float* out, cudaStream_t stream) {
  MMSIM_REQUIRE(parts && out && nparts >= 1 && rows >= 0 && part_stride >= rows * NPIV, MMSIM_ERR_ARG, "merge_pivots: bad arguments");
  if (rows == 0) return MMSIM_OK;
  merge_pivots_kernel<<<unsigned((rows + 127) / 128), 128, 0, stream>>>(parts, nparts, part_stride, int(rows), out);
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace knn
}  // namespace mmsim

#ifdef MMSIM_SWEEP_DEBUG
// experiment builds only (MMSIM_DEBUG_BUILD=1 python -m multimodal_similarity_b200.build --force); not part of the C-ABI
extern "C" __attribute__((visibility("default"))) int mmsim_debug_counters(unsigned long long* out, int n, int reset) {  // n <= 16
  unsigned long long h[16] = {};
  if (cudaMemcpyFromSymbol(h, mmsim::knn::g_dbg, sizeof(h)) != cudaSuccess) return -1;
  for (int i = 0; i < n && i < 16; ++i) out[i] = h[i];
  if (reset) {
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(mmsim::knn::g_dbg, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif
