// The main sweep of the kNN retrieval, query-streaming form (included by knn_tc.cu after the legacy kernel).
//
// knn_tc_kernel<MODE_SWEEP> keeps a 128-query block resident and streams the gallery: per 128 x 256 x K tile the SM's
// shared memory takes K/64 * 32 KB of TMA writes (the gallery tile) and serves K/64 * 48 KB of operand reads -- at K = 128
// that is 160 KB per tile against a 128 B/cycle shared memory, 1,250 cycles for MMAs that need 1,024.
// This kernel swaps the roles: a 256-row GALLERY tile stays resident (B operand, K/64 * 32 KB, plus its norm pack) and ALL
// query blocks of the item stream through a ring of 16 KB K atoms (A operand).  Same MMAs, same accumulator layout (TMEM
// lane = query row, column = gallery row), same three-level epilogue, but the TMA writes per tile halve and so does the
// L2 -> SM traffic of the whole sweep (every CTA streams the 25 MB of fp16 queries per gallery tile, not the 256 MB gallery
// per query block).  With every row closed (no candidate ever found; scripts/sweep_ablate.py, one B200, 100k x 1M x 128)
// the pipeline runs at 19.7 ms against 22.0 ms for the gallery-streaming kernel (gpurun_out/r2_f_ablate_*.log).
//
// A query row is now visited by every CTA, so its state is global (L2) words:
//   tau[row]    float, only ever lowered (atomic min): every CTA filters with the current or an older = larger value, so
//               "every unlogged gallery row has key >= the FINAL tau[row]" holds by monotonicity -- the certificate's premise
//   state[row]  u32: log cursor | entries below ladder rung 1 | below rung 0
//   log[row]    ONE candidate log per query (no per-split logs, no threshold inheritance between splits)
// Appending a candidate therefore costs two dependent L2 round trips (rungs, cursor) -- inside the scan that is fatal: the
// scan of a tile sits between "accumulator ready" and the next tile's drain (24.3 ms, gpurun_out/r2_e_bench.json).  So the
// epilogue warps only PUSH the candidate {key, row, column, threshold used} into a shared-memory ring (one shared-memory
// atomic per warp call for the tickets + a 16-byte store per candidate), and the two otherwise idle warps of warpgroup 0
// drain it: 64 lanes, two entries in flight each, do the L2 work with all the latency tolerance they need.  The lane that
// makes a rung counter reach KPT lowers tau to that rung.
//
// STATUS: opt-in (MMSIM_KNN_SWEEP=q); results are identical to the default kernel's
// (tests/test_gpu_knn.py::test_query_streaming_sweep_matches).  Measured on one B200, 100k x 1M x 128, sweep alone
// (scripts/sweep_ablate.py; gpurun_out/r2_f .. r2_p_ablate_*.log, summarised in profiles/r2_sweep_experiments.txt):
//   every row closed (pipeline alone)   this kernel 19.6-20.2 ms   gallery-streaming 21.3-22.2 ms
//   product                             this kernel 22.2-23.6 ms   gallery-streaming 22.3-23.2 ms
// The role swap buys 10% of pipeline, and the candidate path gives it back: entering it costs 1.2 ms, the ticket 0.4 ms,
// storing the entries 2.2 ms (each store waits for its slot's flag), against 0.6-1.2 ms for the gallery-streaming kernel's
// per-row shared-memory state -- at 128-d the epilogue warps have no slack, so every cycle of the candidate path shows.  At
// 256-d (twice the MMA time per tile) pushing is free but the drain's L2 traffic costs 4 ms.  Variants measured and dropped:
// thresholds / cursors updated from inside the scan (24.3 ms); query chunks with shared-memory state and one or two
// resident gallery tiles (the tile switch every 16 query blocks, or the shallower A ring two tiles leave room for, cost
// what the swap gains: 22.1 / 23.4 ms closed-rows); st.release / ld.acquire ring flags (MEMBAR.ALL.CTA per entry); one
// ring per epilogue warp with register tails (bursts of one warp meet four drain lanes: 26.5 ms).
// Work items are gallery tiles (x query chunks when there are fewer tiles than SMs); CTAs start their pass over the
// query blocks at different offsets.
#pragma once

namespace mmsim {
namespace knn {

struct SweepQArgs {
  const float* gpack;        // [n_tiles][NPACK]
  int nq, n_qblocks;
  int tile_begin, tile_end;  // gallery tiles of this launch
  int n_qchunks, qpc;        // query blocks are cut into n_qchunks chunks of qpc blocks: item = (tile, chunk)
  float* tau;                // [q_rows]  in: initial threshold (-inf for rows past nq), out: final threshold
  unsigned int* state;       // [q_rows]  packed cursor + ladder counters (zero on entry)
  const float* ladder;       // [q_rows][4] = (-, rung0, rung1, tau0)
  int use_pivots;
  uint2* log;                // [row][logcap]
  int logcap;
  int drop;                  // ablation (MMSIM_SWEEP_FLAGS=16): the drain lanes discard the candidates (wrong results)
};

constexpr int SQ_QCAP = 1024;                  // candidate ring entries (16 bytes each)
constexpr uint32_t SQ_POISON = 0xffffffffu;    // row id of the entries that tell the drain lanes to stop

template <int KATOMS>
struct SmemQ {
  static constexpr int NSA = KATOMS <= 2 ? 6 : 4;                 // query K-atom stages (16 KiB each)
  static constexpr int B_OFF = 0;                                  // resident gallery tile: KATOMS x 32 KiB
  static constexpr int A_OFF = B_OFF + KATOMS * B_STAGE_BYTES;
  static constexpr int NORM_OFF = A_OFF + NSA * A_ATOM_BYTES;      // two norm-pack slots (items alternate)
  static constexpr int QENT_OFF = NORM_OFF + 2 * NPACK * 4;        // uint4 [SQ_QCAP] candidate ring
  static constexpr int QFLAG_OFF = QENT_OFF + SQ_QCAP * 16;        // u32   [SQ_QCAP] 2g: free for generation g, 2g+1: full
  static constexpr int QCTL_OFF = QFLAG_OFF + SQ_QCAP * 4;         // u32 tail, u32 head
  static constexpr int BAR_OFF = QCTL_OFF + 16;
  static constexpr int NUM_BARS = 2 * NSA + 2 + 2 + 2 + 2;         // full_a, empty_a | tfull[2] | tempty[2] | bfull, bempty | nempty[2]
  static constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int DYN_BYTES = TMEM_PTR_OFF + 8;
  static_assert(QENT_OFF % 16 == 0 && BAR_OFF % 8 == 0, "alignment");
  static_assert(DYN_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

// tau only ever decreases: atomic min on a float that may be negative (CAS loop; drain lanes only)
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
  unsigned int* a = reinterpret_cast<unsigned int*>(addr);
  unsigned int old = *reinterpret_cast<volatile unsigned int*>(a);
  while (__uint_as_float(old) > v) {
    const unsigned int seen = atomicCAS(a, old, __float_as_uint(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ float ldcg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// Ring flags: volatile shared-memory accesses, NOT st.release / ld.acquire -- those compile to MEMBAR.ALL.CTA, which makes the
// pushing epilogue thread wait for everything it has in flight.  An SM executes one thread's shared-memory accesses in
// program order, and every access of the protocol is volatile (entry before flag on the producer side, flag before entry on
// the consumer side), which is all the ring needs.
__device__ __forceinline__ void sts_u32_volatile(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// one candidate into ring slot `t` (a ticket the caller owns): wait for the slot's generation (normally free already),
// entry, publish
__device__ __forceinline__ void sq_put(uint32_t qbase, uint32_t t, uint32_t key_bits, uint32_t grow, uint32_t col, uint32_t tau_bits) {
  constexpr uint32_t ENT = 0, FLAG = SQ_QCAP * 16;
  const uint32_t slot = t & (SQ_QCAP - 1), gen2 = (t / SQ_QCAP) * 2;
  const uint32_t faddr = qbase + FLAG + slot * 4;
  while (lds_u32_volatile(faddr) != gen2) __nanosleep(32);
  asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(qbase + ENT + slot * 16), "r"(key_bits), "r"(grow), "r"(col),
               "r"(tau_bits)
               : "memory");
  sts_u32_volatile(faddr, gen2 + 1);
}
__device__ __forceinline__ void sq_push(uint32_t qbase, uint32_t key_bits, uint32_t grow, uint32_t col, uint32_t tau_bits) {
  sq_put(qbase, atoms_add32(qbase + SQ_QCAP * 20, 1u), key_bits, grow, col, tau_bits);
}

// Rare path of the scan, out of line: the 8 keys of one column group with at least one candidate somewhere in the warp
// (called by all 32 lanes).  ONE shared-memory atomic per call claims the tickets of all the warp's candidates.
__device__ __noinline__ void sweepq_group8(float k0, float k1, float k2, float k3, float k4, float k5, float k6, float k7,
                                           int cbase, float tau, int grow, uint32_t qbase, int abl) {
  const float key[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
  const uint32_t lane = threadIdx.x & 31;
  if (abl == 32) return;            // ablation: the cost of entering the candidate path at all
  uint32_t hits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) hits |= (key[j] < tau ? 1u : 0u) << j;
  const uint32_t n = __popc(hits);
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= uint32_t(o)) incl += y;
  }
  uint32_t base = 0;
  if (lane == 31) base = atoms_add32(qbase + SQ_QCAP * 20, incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t t = base + incl - n;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (hits & (1u << j)) sq_put(qbase, t++, __float_as_uint(key[j]), uint32_t(grow), uint32_t(cbase + j), __float_as_uint(tau));
}

// A drain lane: pops entries until it meets a poison entry; each entry is appended to its query's log in global memory.
// Two entries are in flight per lane, and within an entry the cursor atomic does not wait for the ladder rungs: the lanes
// are latency-bound (L2 round trips).
struct SqEntry {
  uint32_t kb, grow, col, tb;
};
__device__ __forceinline__ bool sq_pop(uint32_t qbase, SqEntry& e) {
  constexpr uint32_t ENT = 0, FLAG = SQ_QCAP * 16, HEAD = SQ_QCAP * 20 + 4;
  const uint32_t h = atoms_add32(qbase + HEAD, 1u);
  const uint32_t slot = h & (SQ_QCAP - 1), gen2 = (h / SQ_QCAP) * 2;
  const uint32_t faddr = qbase + FLAG + slot * 4;
  while (lds_u32_volatile(faddr) != gen2 + 1) __nanosleep(64);
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.kb), "=r"(e.grow), "=r"(e.col), "=r"(e.tb)
               : "r"(qbase + ENT + slot * 16)
               : "memory");
  sts_u32_volatile(faddr, gen2 + 2);                   // the slot is free for the next generation
  return e.grow != SQ_POISON;
}
__device__ __forceinline__ void sq_drain(uint32_t qbase, const SweepQArgs& a) {
  for (;;) {
    SqEntry e0, e1;
    const bool v0 = sq_pop(qbase, e0);
    if (!v0) return;
    const bool v1 = sq_pop(qbase, e1);                   // (a poison entry here ends the lane after e0 is done)
    if (a.drop) {
      if (!v1) return;
      continue;
    }
    // cursor first (the log slot needs nothing else), the rungs travel meanwhile
    const uint32_t c0 = atomicAdd(a.state + e0.grow, 1u);
    const uint32_t c1 = v1 ? atomicAdd(a.state + e1.grow, 1u) : 0u;
    float4 p0 = make_float4(-kInf, -kInf, -kInf, -kInf), p1 = p0;
    if (a.use_pivots) {
      p0 = __ldg(reinterpret_cast<const float4*>(a.ladder + size_t(e0.grow) * 4));
      if (v1) p1 = __ldg(reinterpret_cast<const float4*>(a.ladder + size_t(e1.grow) * 4));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const SqEntry& e = i ? e1 : e0;
      if (i && !v1) break;
      const uint32_t cur = (i ? c1 : c0) & CUR_MASK;
      const float4 pp = i ? p1 : p0;
      if (int(cur) < a.logcap) {
        a.log[size_t(e.grow) * a.logcap + cur] = make_uint2(e.kb, e.col);
      } else {
        atomic_min_f32(a.tau + e.grow, -kInf);   // log full: the row is uncertifiable from here on, stop accepting candidates
      }
      // rung counters: the entry that makes a rung's count reach KPT lowers the threshold to that rung (the KPT smallest
      // keys all lie below it).  A counter stops moving once its rung is no longer below the threshold the scanning CTA
      // filtered with; CTAs with an older threshold may push it a little further, never past 511 before the cursor closes
      // the row.
      const float key = __uint_as_float(e.kb), tau = __uint_as_float(e.tb);
      const bool b1 = key < pp.z && pp.z < tau, b0 = key < pp.y && pp.y < tau;
      if (b1 | b0) {
        const uint32_t old = atomicAdd(a.state + e.grow, (b1 ? (1u << CN1_SHIFT) : 0u) | (b0 ? (1u << CN0_SHIFT) : 0u));
        if (b1 && ((old >> CN1_SHIFT) & 511u) == uint32_t(KPT - 1)) atomic_min_f32(a.tau + e.grow, pp.z);
        if (b0 && ((old >> CN0_SHIFT) & 511u) == uint32_t(KPT - 1)) atomic_min_f32(a.tau + e.grow, pp.y);
      }
    }
    if (!v1) return;
  }
}

template <int KATOMS, int NEPI>
__global__ void __launch_bounds__(128 + NEPI * 32, 1)
knn_sweepq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_g,
                  const __grid_constant__ SweepQArgs a) {
  using S = SmemQ<KATOMS>;
  constexpr int NSA = S::NSA;
  constexpr int NH = NEPI / 4;
  constexpr int EPI_THREADS = NEPI * 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();

  uint8_t* smem_b = smem + S::B_OFF;
  uint8_t* smem_a = smem + S::A_OFF;
  float* norm_slots = reinterpret_cast<float*>(smem + S::NORM_OFF);
  uint32_t* q_flag = reinterpret_cast<uint32_t*>(smem + S::QFLAG_OFF);
  uint32_t* q_ctl = reinterpret_cast<uint32_t*>(smem + S::QCTL_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* full_a = bars;                    // [NSA] TMA -> MMA
  uint64_t* empty_a = bars + NSA;             // [NSA] MMA -> TMA
  uint64_t* tfull = bars + 2 * NSA;           // [2]   MMA -> epilogue
  uint64_t* tempty = bars + 2 * NSA + 2;      // [2]   epilogue -> MMA
  uint64_t* bfull = bars + 2 * NSA + 4;       //       gallery tile + norm pack landed
  uint64_t* bempty = bars + 2 * NSA + 5;      //       all MMAs of the item done: the gallery tile may be overwritten
  uint64_t* nempty = bars + 2 * NSA + 6;      // [2]   epilogue finished an item: its norm-pack slot may be refilled
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::TMEM_PTR_OFF);
  const uint32_t qbase = ptx::smem_u32(smem + S::QENT_OFF);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < SQ_QCAP; i += blockDim.x) q_flag[i] = 0u;
  if (threadIdx.x < 4) q_ctl[threadIdx.x] = 0u;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_g);
    for (int i = 0; i < NSA; ++i) {
      ptx::mbar_init(&full_a[i], 1);
      ptx::mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], EPI_THREADS);
      ptx::mbar_init(&nempty[i], EPI_THREADS);
    }
    ptx::mbar_init(bfull, 1);
    ptx::mbar_init(bempty, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 2 * BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_items = (a.tile_end - a.tile_begin) * a.n_qchunks;
  // item -> gallery tile, query block range [qb_lo, qb_lo + nqb), starting offset of the pass
  auto item_range = [&](int item, int& tile, int& qb_lo, int& nqb, int& rot) {
    const int t = item / a.n_qchunks, c = item - t * a.n_qchunks;
    tile = a.tile_begin + t;
    qb_lo = c * a.qpc;
    nqb = min(a.n_qblocks, qb_lo + a.qpc) - qb_lo;
    rot = int((uint32_t(item) * 2654435761u >> 8) % uint32_t(nqb));
  };

  const uint32_t wg = __shfl_sync(0xffffffffu, threadIdx.x >> 7, 0);
  if (wg == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // =============================================================== TMA producer (one thread)
      if (lane == 0) {
        uint32_t sa = 0, ic = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
          int tile, qb_lo, nqb, rot;
          item_range(item, tile, qb_lo, nqb, rot);
          ptx::mbar_wait(bempty, (ic & 1) ^ 1);
          ptx::mbar_wait(&nempty[ic & 1], ((ic >> 1) & 1) ^ 1);
          ptx::mbar_expect_tx(bfull, KATOMS * B_STAGE_BYTES + NPACK * 4);
          for (int ka = 0; ka < KATOMS; ++ka)
            ptx::tma_load_2d(smem_b + ka * B_STAGE_BYTES, &tm_g, ka * KATOM, tile * BN, bfull);
          ptx::bulk_load_1d(norm_slots + (ic & 1) * NPACK, a.gpack + size_t(tile) * NPACK, NPACK * 4, bfull);
          for (int j = 0; j < nqb; ++j) {
            int qo = j + rot;
            if (qo >= nqb) qo -= nqb;
            const int qb = qb_lo + qo;
            for (int ka = 0; ka < KATOMS; ++ka, ++sa) {
              const uint32_t stage = sa % NSA, phase = (sa / NSA) & 1;
              ptx::mbar_wait(&empty_a[stage], phase ^ 1);
              ptx::mbar_expect_tx(&full_a[stage], A_ATOM_BYTES);
              ptx::tma_load_2d(smem_a + stage * A_ATOM_BYTES, &tm_q, ka * KATOM, qb * BM, &full_a[stage]);
            }
          }
        }
      }
    } else if (warp == 1) {
      // =============================================================== MMA issuer
      const uint32_t idesc = ptx::umma_idesc_f16(BM, BN);
      const uint64_t adesc0 = ptx::umma_desc_k128(ptx::smem_u32(smem_a));
      const uint64_t bdesc0 = ptx::umma_desc_k128(ptx::smem_u32(smem_b));
      uint32_t sa = 0, tc = 0, ic = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
        int tile, qb_lo, nqb, rot;
        item_range(item, tile, qb_lo, nqb, rot);
        ptx::mbar_wait(bfull, ic & 1);
        ptx::tc_fence_after();
        for (int j = 0; j < nqb; ++j, ++tc) {
          const uint32_t as = tc & 1;
          ptx::mbar_wait(&tempty[as], ((tc >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          for (int ka = 0; ka < KATOMS; ++ka, ++sa) {
            const uint32_t stage = sa % NSA, phase = (sa / NSA) & 1;
            ptx::mbar_wait(&full_a[stage], phase);
            ptx::tc_fence_after();
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < KATOM / 16; ++k) {
                const uint64_t ad = adesc0 + uint64_t(stage * (A_ATOM_BYTES >> 4) + k * 2);
                const uint64_t bd = bdesc0 + uint64_t(ka * (B_STAGE_BYTES >> 4) + k * 2);
                ptx::umma_f16(tmem_base + as * BN, ad, bd, idesc, (ka | k) != 0);
              }
              ptx::umma_commit(&empty_a[stage]);
              if (ka == KATOMS - 1) ptx::umma_commit(&tfull[as]);
            }
            __syncwarp();
          }
        }
        if (lane == 0) ptx::umma_commit(bempty);
        __syncwarp();
      }
    } else {
      // =============================================================== candidate drain (warps 2 and 3: 64 independent lanes)
      sq_drain(qbase, a);
    }
  } else {
    // =============================================================== epilogue warps: TMEM -> candidate ring
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const uint32_t q = warp & 3;
    const uint32_t h = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t norm_u32 = ptx::smem_u32(norm_slots);
    const uint32_t taddr0 = tmem_base + ((q * 32) << 16) + h * 32;
    constexpr int CPW = (BN / 32) / NH;
    uint32_t tc = 0, ic = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
      int tile, qb_lo, nqb, rot;
      item_range(item, tile, qb_lo, nqb, rot);
      // The item's norm pack landed with its gallery tile (bfull): the MMA warp waited for that before it issued the MMAs
      // whose completion (tfull) the epilogue waits for below -- the chain the legacy kernel relies on for its packs.
      const uint32_t nrm = norm_u32 + (ic & 1) * NPACK * 4;
      const int col0 = tile * BN;
      float tau_next = ldcg_f32(a.tau + size_t(qb_lo + rot) * BM + row);

      for (int j = 0; j < nqb; ++j, ++tc) {
        int qo = j + rot;
        if (qo >= nqb) qo -= nqb;
        const int grow = (qb_lo + qo) * BM + row;
        const float tau = tau_next;
        {   // the next block's thresholds travel from L2 while this tile is scanned
          int qn = qo + 1;
          if (qn >= nqb) qn -= nqb;
          tau_next = ldcg_f32(a.tau + size_t(qb_lo + qn) * BM + row);
        }
        const uint32_t as = tc & 1;
        ptx::mbar_wait(&tfull[as], (tc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = taddr0 + as * BN;

        auto scan_chunk = [&](float (&v)[32], int c) {
          float gm[4];
#pragma unroll
          for (int g = 0; g < 4; ++g)
            gm[g] = min3(min3(v[g * 8], v[g * 8 + 1], v[g * 8 + 2]), min3(v[g * 8 + 3], v[g * 8 + 4], v[g * 8 + 5]),
                         fminf(v[g * 8 + 6], v[g * 8 + 7]));
          const float m = fminf(min3(gm[0], gm[1], gm[2]), gm[3]);
          const float nm32 = lds_f32(nrm + (BN + 32 + c) * 4);
          if (!__any_sync(0xffffffffu, m < tau - nm32)) return;
          const float4 nm8 = lds_f32x4(nrm + (BN + c * 4) * 4);
          const uint32_t mine = (gm[0] < tau - nm8.x ? 1u : 0u) | (gm[1] < tau - nm8.y ? 2u : 0u) |
                                (gm[2] < tau - nm8.z ? 4u : 0u) | (gm[3] < tau - nm8.w ? 8u : 0u);
          const uint32_t groups = __reduce_or_sync(0xffffffffu, mine);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!(groups & (1u << g))) continue;
            const float4 n0 = lds_f32x4(nrm + (c * 32 + g * 8) * 4);
            const float4 n1 = lds_f32x4(nrm + (c * 32 + g * 8 + 4) * 4);
            sweepq_group8(v[g * 8 + 0] + n0.x, v[g * 8 + 1] + n0.y, v[g * 8 + 2] + n0.z, v[g * 8 + 3] + n0.w,
                          v[g * 8 + 4] + n1.x, v[g * 8 + 5] + n1.y, v[g * 8 + 6] + n1.z, v[g * 8 + 7] + n1.w,
                          col0 + c * 32 + g * 8, tau, grow, qbase, a.drop);
          }
        };

        float va[32], vb[32];
        ptx::tmem_ld32(taddr, va);
        ptx::tmem_ld32(taddr + NH * 32, vb);
#pragma unroll 1
        for (int cp = 0; cp < CPW; cp += 2) {
          const int c0 = h + cp * NH, c1 = c0 + NH;
          const bool last = cp + 2 >= CPW;
          ptx::tmem_ld_wait(va);
          ptx::tmem_ld_wait(vb);
          if (last) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty[as]);
          }
          scan_chunk(va, c0);
          if (!last) ptx::tmem_ld32(taddr + (cp + 2) * NH * 32, va);
          scan_chunk(vb, c1);
          if (!last) ptx::tmem_ld32(taddr + (cp + 3) * NH * 32, vb);
        }
      }
      ptx::mbar_arrive(&nempty[ic & 1]);
    }
    // every candidate of this CTA is in the ring: one poison entry per drain lane (a lane stops at its first one)
    epi_bar_sync(EPI_THREADS);
    if (threadIdx.x - 128 < 64) sq_push(qbase, 0u, SQ_POISON, 0u, 0u);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

// tau[row] = initial threshold (the ladder's, or +inf when everything is logged), -inf for the padding rows; state = 0
__global__ void sweepq_init_kernel(const float* __restrict__ ladder, int use_pivots, int nq, int rows, float* __restrict__ tau,
                                   unsigned int* __restrict__ state, int close_rows /* ablation: no row accepts anything */) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  tau[r] = r < nq && !close_rows ? (use_pivots ? ladder[size_t(r) * 4 + 3] : kInf) : -kInf;
  state[r] = 0u;
}

// state / tau -> the (count, final threshold) arrays the re-rank kernel reads (one "split" per query)
__global__ void sweepq_finish_kernel(const unsigned int* __restrict__ state, const float* __restrict__ tau, int rows,
                                     int* __restrict__ log_cnt, float* __restrict__ log_tau) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  log_cnt[r] = int(state[r] & CUR_MASK);
  log_tau[r] = tau[r];
}

}  // namespace knn
}  // namespace mmsim
