// Round-2 microbenchmark: the query tile as the TMEM-resident A operand of tcgen05.mma (A from tensor memory, B from
// shared memory) and a finer accumulator ring -- the structural option of DESIGN.md section 9.1.
//   part 1 (numeric): D = A . B^T for one 128 x 256 x 128 tile, once with A from shared memory (the product kernel's form)
//                     and once with A written into TMEM by tcgen05.st (lane = row, one 32-bit column per pair of K elements),
//                     as FOUR N = 64 sub-tile MMAs; the two accumulators must agree bit for bit.
//   part 2 (speed)  : the MMA <-> epilogue ring with S accumulator stages of N columns each (S * N <= 384), A from TMEM,
//                     B resident in shared memory, 8 epilogue warps draining every column -- cycles per 256 gallery rows.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I multimodal_similarity_b200/csrc \
//        scripts/ubench/tmem_a_operand.cu -o scripts/ubench/tmem_a_operand
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mmsim;

constexpr int BM = 128, BN = 256, KATOMS = 2;
constexpr int A_ATOM = BM * 128, B_ATOM = BN * 128;          // bytes of one 64-wide K atom
constexpr int A_COL = 384;                                   // TMEM columns [384, 448): the A operand (K = 128)

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline float a_val(int m, int k) { return float((m * 7 + k * 3) % 13 - 6) * 0.25f; }
__host__ __device__ inline float b_val(int n, int k) { return float((n * 5 + k * 11) % 17 - 8) * 0.125f; }
// byte offset of logical element (row, k) of a K-major SWIZZLE_128B operand whose K atoms are `atom_bytes` apart
__device__ __forceinline__ uint32_t swz(int row, int k, int atom_bytes) {
  const int ka = k >> 6, kk = k & 63;
  return ka * atom_bytes + row * 128 + (((kk >> 3) ^ (row & 7)) << 4) + (kk & 7) * 2;
}

// warps 0..3: write A into TMEM (their lane quarters); whole CTA fills the shared-memory operands
__device__ void fill_operands(uint8_t* smem, uint32_t tmem_base, uint32_t warp, uint32_t lane) {
  for (int i = threadIdx.x; i < BM * 128; i += blockDim.x) {
    const int m = i >> 7, k = i & 127;
    *reinterpret_cast<__half*>(smem + swz(m, k, A_ATOM)) = __float2half_rn(a_val(m, k));
  }
  for (int i = threadIdx.x; i < BN * 128; i += blockDim.x) {
    const int n = i >> 7, k = i & 127;
    *reinterpret_cast<__half*>(smem + KATOMS * A_ATOM + swz(n, k, B_ATOM)) = __float2half_rn(b_val(n, k));
  }
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 h = __halves2half2(__float2half_rn(a_val(row, 2 * (c0 + j))), __float2half_rn(a_val(row, 2 * (c0 + j) + 1)));
        r[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st16(tmem_base + ((warp * 32) << 16) + A_COL + c0, r);
    }
    tmem_st_wait();
  }
}

__global__ void __launch_bounds__(64 + 8 * 32, 1)
numeric_kernel(int* mismatches, float* sample) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_ptr;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 1) ptx::tmem_alloc(&tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  fill_operands(smem, tmem_base, warp, lane);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) {
    if (lane == 0) {
      const uint64_t ad0 = ptx::umma_desc_k128(ptx::smem_u32(smem));
      const uint64_t bd0 = ptx::umma_desc_k128(ptx::smem_u32(smem + KATOMS * A_ATOM));
      // reference form: A and B from shared memory, one N = 256 MMA per K step -> columns [0, 256)
      for (int ka = 0; ka < KATOMS; ++ka)
        for (int k = 0; k < 4; ++k)
          ptx::umma_f16(tmem_base, ad0 + uint64_t(ka * (A_ATOM >> 4) + k * 2), bd0 + uint64_t(ka * (B_ATOM >> 4) + k * 2),
                        ptx::umma_idesc_f16(BM, BN), (ka | k) != 0);
      // A from TMEM, four N = 64 sub-tiles -> columns [256, 320), but 4 x 64 = 256 columns do not fit beside the reference
      // and A: the sub-tiles go to [256, 320) one after the other and are compared one at a time (sub-tile j in pass j)
    }
    __syncwarp();
  }
  int bad = 0;
  for (int j = 0; j < 4; ++j) {
    if (warp == 1) {
      if (lane == 0) {
        const uint64_t bd0 = ptx::umma_desc_k128(ptx::smem_u32(smem + KATOMS * A_ATOM));
        for (int ka = 0; ka < KATOMS; ++ka)
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tmem_base + 256, tmem_base + A_COL + ka * 32 + k * 8,
                        bd0 + uint64_t(ka * (B_ATOM >> 4) + ((j * 64 * 128) >> 4) + k * 2), ptx::umma_idesc_f16(BM, 64), (ka | k) != 0);
        ptx::umma_commit(&done);
      }
      __syncwarp();
    }
    ptx::mbar_wait(&done, j & 1);
    ptx::tc_fence_after();
    if (warp >= 2 && warp < 6) {
      const uint32_t q = warp & 3;
      const int row = q * 32 + lane;
      for (int c = 0; c < 64; c += 32) {
        float ref[32], got[32];
        ptx::tmem_ld32(tmem_base + ((q * 32) << 16) + j * 64 + c, ref);
        ptx::tmem_ld32(tmem_base + ((q * 32) << 16) + 256 + c, got);
        ptx::tmem_ld_wait(ref);
        ptx::tmem_ld_wait(got);
        for (int i = 0; i < 32; ++i) {
          float want = 0.f;
          const int n = j * 64 + c + i;
          for (int k = 0; k < 128; ++k) want += a_val(row, k) * b_val(n, k);     // exact in fp32: small dyadic values
          if (__float_as_uint(ref[i]) != __float_as_uint(got[i]) || ref[i] != want) ++bad;
          if (row == 5 && n == 77) { sample[0] = ref[i]; sample[1] = got[i]; sample[2] = want; }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  if (bad) atomicAdd(mismatches, bad);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ring of S stages of N columns; `ts` != 0: A from TMEM, else from shared memory.  tiles = number of 256-row gallery tiles.
__global__ void __launch_bounds__(64 + 8 * 32, 1)
ring_kernel(int S, int N, int ts, int tiles, int nepi, long long* cycles, float* sink, int variant) {
  // variant bits: 1 = the MMA warp skips tcgen05.fence::after_thread_sync after waiting for a drained stage,
  //               2 = the epilogue warps skip tcgen05.fence::before_thread_sync before releasing a stage,
  //               4 = only lane 0 of the MMA warp waits on the barriers (the other lanes idle at the __syncwarp)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[34];
  __shared__ uint32_t tmem_ptr;
  uint64_t* tfull = bars;        // [S <= 16]
  uint64_t* tempty = bars + 16;  // [S]
  uint64_t* done = bars + 32;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], nepi * 32); }
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  fill_operands(smem, tmem_base, warp, lane);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const int subs = BN / N;                       // sub-tiles per 256-row gallery tile
  const long long t0 = clock64();
  if (warp == 1) {
    const uint32_t idesc = ptx::umma_idesc_f16(BM, N);
    const uint64_t ad0 = ptx::umma_desc_k128(ptx::smem_u32(smem));
    const uint64_t bd0 = ptx::umma_desc_k128(ptx::smem_u32(smem + KATOMS * A_ATOM));
    // (the issue loop must be cheap: every tcgen05.mma below has compile-time descriptor offsets -- a first version with
    // runtime offsets and a branch per MMA was ISSUE-bound at ~175 cycles per MMA, whatever the ring looked like)
    const uint64_t bstep = uint64_t((N * 128) >> 4);
    int it = 0, st = 0, ph = 1;
    for (int t = 0; t < tiles; ++t) {
      uint64_t bd = bd0;
      for (int j = 0; j < subs; ++j, ++it, bd += bstep) {
        if (!(variant & 4) || lane == 0) ptx::mbar_wait(&tempty[st], ph);
        if (!(variant & 1)) ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t d = tmem_base + st * N;
          if (ts) {
            const uint32_t a = tmem_base + A_COL;
#pragma unroll
            for (int ka = 0; ka < KATOMS; ++ka)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(d, a + ka * 32 + k * 8, bd + uint64_t(ka * (B_ATOM >> 4) + k * 2), idesc, (ka | k) != 0);
          } else {
#pragma unroll
            for (int ka = 0; ka < KATOMS; ++ka)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16(d, ad0 + uint64_t(ka * (A_ATOM >> 4) + k * 2), bd + uint64_t(ka * (B_ATOM >> 4) + k * 2), idesc, (ka | k) != 0);
          }
          ptx::umma_commit(&tfull[st]);
        }
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1; }
      }
    }
    if (lane == 0) ptx::umma_commit(done);
    __syncwarp();
    ptx::mbar_wait(done, 0);
  } else if (warp >= 2 && int(warp) < 2 + nepi) {
    const uint32_t q = warp & 3, h = (warp - 2) >> 2;
    const int nh = nepi / 4;
    const int chunks = N / 32;                   // 32-column chunks per stage
    float acc = 0.f;
    int st = -1, ph = 1;
    for (int t = 0; t < tiles; ++t)
      for (int j = 0; j < subs; ++j) {
        if (++st == S) st = 0;
        if (st == 0) ph ^= 1;
        ptx::mbar_wait(&tfull[st], ph);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((q * 32) << 16) + st * N;
        float v[2][32];
        if (int(h) >= chunks) { ptx::tc_fence_before(); ptx::mbar_arrive(&tempty[st]); }
        for (int c = h; c < chunks; c += 2 * nh) {
          const bool two = c + nh < chunks, last = c + 2 * nh >= chunks;
          ptx::tmem_ld32(taddr + c * 32, v[0]);
          if (two) ptx::tmem_ld32(taddr + (c + nh) * 32, v[1]);
          ptx::tmem_ld_wait(v[0]);
          if (two) ptx::tmem_ld_wait(v[1]);
          if (last) { if (!(variant & 2)) ptx::tc_fence_before(); ptx::mbar_arrive(&tempty[st]); }
          auto scan = [&](const float (&x)[32]) {          // (static register indexing: no local memory)
            float m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
#pragma unroll
            for (int i = 0; i < 32; i += 2) m[(i / 2) & 3] = fminf(fminf(m[(i / 2) & 3], x[i]), x[i + 1]);
            acc += fminf(fminf(m[0], m[1]), fminf(m[2], m[3]));
          };
          scan(v[0]);
          if (two) scan(v[1]);
        }
      }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 32) cycles[blockIdx.x] = t1 - t0;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

int main() {
  const int smem = KATOMS * A_ATOM + KATOMS * B_ATOM;
  int* bad;
  float* sample;
  cudaMalloc(&bad, 4);
  cudaMalloc(&sample, 16);
  cudaMemset(bad, 0, 4);
  cudaFuncSetAttribute(numeric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  numeric_kernel<<<1, 64 + 8 * 32, smem>>>(bad, sample);
  cudaError_t err = cudaDeviceSynchronize();
  int hbad = -1;
  float hs[3] = {0, 0, 0};
  cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs, sample, 12, cudaMemcpyDeviceToHost);
  printf("numeric: %s, mismatching accumulator elements (A from TMEM, N = 64 sub-tiles vs A from shared memory, N = 256, vs fp32 "
         "reference) = %d of 32768; D[5][77] = %g / %g / %g\n", cudaGetErrorString(err), hbad, hs[0], hs[1], hs[2]);
  if (err != cudaSuccess || hbad != 0) return 1;

  const int grid = 148, tiles = 4000;
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, grid * sizeof(long long));
  cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int S, N, ts, nepi, variant; };
  const Cfg cfgs[] = {{2, 256, 0, 8, 0}, {2, 256, 0, 8, 1}, {2, 256, 0, 8, 3}, {2, 256, 0, 8, 4}, {6, 64, 1, 8, 0}, {6, 64, 1, 8, 1}, {6, 64, 1, 8, 2},
                      {6, 64, 1, 8, 3}, {6, 64, 1, 8, 4}, {6, 64, 1, 8, 7}, {3, 128, 1, 8, 7}, {3, 128, 0, 8, 7}};
  for (const Cfg& c : cfgs) {
    ring_kernel<<<grid, 64 + 8 * 32, smem>>>(c.S, c.N, c.ts, tiles / 4, c.nepi, cyc, sink, c.variant);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ring_kernel<<<grid, 64 + 8 * 32, smem>>>(c.S, c.N, c.ts, tiles, c.nepi, cyc, sink, c.variant);
    cudaEventRecord(e1);
    err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("ring S=%d N=%d ts=%d: %s\n", c.S, c.N, c.ts, cudaGetErrorString(err)); return 1; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < grid; ++i) s += h[i];
    s /= grid;
    printf("ring: %d stages x %3d columns, A from %s, %d epilogue warps, variant %d: %7.1f cycles per 256 gallery rows, "
           "%.3f ms -> %.0f TFLOP/s\n", c.S, c.N, c.ts ? "TMEM  " : "shared", c.nepi, c.variant, s / tiles, ms,
           2.0 * BM * BN * 128 * double(tiles) * grid / ms / 1e9);
  }
  return 0;
}
