// Microbenchmarks behind DESIGN.md's floor for the kNN sweep: what bounds a 128x256 (K=128) tile on one SM --
// tcgen05.mma dispatch rate, tcgen05.ld (TMEM -> registers) throughput, or both together?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I multimodal_similarity_b200/csrc \
//        scripts/ubench/tmem_mma.cu -o scripts/ubench/tmem_mma
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mmsim;

constexpr int BM = 128, BN = 256;
constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;

template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[X]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// wait flavours (round 2): 0 = mbarrier.try_wait with the library's 200 us suspend hint, 1 = pure test_wait spin,
// 2 = try_wait without a hint (system default suspend time)
__device__ int g_wait_mma = 0, g_wait_epi = 0;
__device__ __forceinline__ void wait_as(int flavour, uint64_t* bar, uint32_t parity) {
  if (flavour == 1) {
    while (!ptx::mbar_test_wait(bar, parity)) {}
  } else if (flavour == 2) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
  } else {
    ptx::mbar_wait(bar, parity);
  }
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode bits: 16 = the MMA warp commits per tile but never waits for a drained accumulator (cost of the commits alone; use
// without bit 1), 32 = the epilogue warps take part in the hand-off but load nothing, 64 = ... load a quarter of the columns
// mode bits: 1 = epilogue warps drain TMEM, 2 = MMA warp issues tiles, 4 = drain waits for the tile's MMA (pipelined
// like the real kernel: 2 accumulator stages, tfull/tempty), 8 = min-tree over the loaded values (ALU work)
template <int X, int OUT = 2, int NT = 64 + 16 * 32>
__global__ void __launch_bounds__(NT, 1)
bench_kernel(int mode, int nepi, int tiles, int katoms, int outstanding, long long* cycles, float* sink, int dcols, int ns,
             uint32_t idesc_in) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // dcols: 32-bit TMEM columns per accumulator (256: fp32 D; 128: fp16 D, two per column); ns: accumulator stages (<= 4)
  __shared__ uint64_t bars[12];
  __shared__ uint32_t tmem_ptr;
  uint64_t* tfull = bars;       // [ns]
  uint64_t* tempty = bars + 4;  // [ns]
  uint64_t* done = bars + 8;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // finite fp16 pattern in the operand tiles
  for (int i = threadIdx.x; i < (4 * A_BYTES + 4 * B_BYTES) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003800u ^ ((i * 2654435761u) & 0x03ff03ffu);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], nepi * 32); }
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 1) ptx::tmem_alloc(&tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  const int wm = g_wait_mma, we = g_wait_epi;
  const long long t0 = clock64();
  if (warp == 1) {
    if (mode & 2) {
      const uint32_t idesc = idesc_in;
      const uint64_t ad0 = ptx::umma_desc_k128(ptx::smem_u32(smem));
      const uint64_t bd0 = ptx::umma_desc_k128(ptx::smem_u32(smem + 4 * A_BYTES));
      for (int t = 0; t < tiles; ++t) {
        const uint32_t as = t % ns;
        if ((mode & 4) && !(mode & 16)) { wait_as(wm, &tempty[as], ((t / ns) & 1) ^ 1); ptx::tc_fence_after(); }
        if (lane == 0) {
          for (int ka = 0; ka < katoms; ++ka)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(tmem_base + as * dcols, ad0 + uint64_t(ka * (A_BYTES >> 4) + k * 2),
                            bd0 + uint64_t(((t + ka) & 3) * (B_BYTES >> 4) + k * 2), idesc, (ka | k) != 0);
          if (mode & 4) ptx::umma_commit(&tfull[as]);
        }
        __syncwarp();
      }
      if (lane == 0) ptx::umma_commit(done);
      __syncwarp();
      ptx::mbar_wait(done, 0);
    }
  } else if (warp >= 2 && int(warp) < 2 + nepi && (mode & 1)) {
    const uint32_t q = warp & 3, h = (warp - 2) >> 2;
    const int nh = nepi / 4;
    const int cpw = (mode & 32) ? 0 : ((mode & 64) ? (dcols / 4 / X) / nh : (dcols / X) / nh);   // chunks per warp per tile (32: none, 64: a quarter)
    float acc = 0.f;
    uint32_t v[OUT][X];
    for (int t = 0; t < tiles; ++t) {
      const uint32_t as = t % ns;
      if (mode & 4) { wait_as(we, &tfull[as], (t / ns) & 1); ptx::tc_fence_after(); }
      const uint32_t taddr = tmem_base + ((q * 32) << 16) + as * dcols;
      const int step = (outstanding >= OUT && cpw >= OUT) ? OUT : 1;
      for (int c = 0; c < cpw; c += step) {
        if (step == OUT) {
#pragma unroll
          for (int o = 0; o < OUT; ++o) tmem_ld<X>(taddr + (h + (c + o) * nh) * X, v[o]);
        } else {
          tmem_ld<X>(taddr + (h + c * nh) * X, v[0]);
        }
        ld_wait();
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
          if (o >= step) break;
          if (mode & 8) {
            float m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
#pragma unroll
            for (int i = 0; i < X; i += 2) m[(i / 2) & 3] = fminf(fminf(m[(i / 2) & 3], __uint_as_float(v[o][i])), __uint_as_float(v[o][i + 1]));
            acc += fminf(fminf(m[0], m[1]), fminf(m[2], m[3]));
          } else {
            uint32_t x = 0;
#pragma unroll
            for (int i = 0; i < X; i += 2) x ^= v[o][i] ^ v[o][i + 1];
            acc += __uint_as_float(x);
          }
        }
      }
      if (mode & 4) { ptx::tc_fence_before(); ptx::mbar_arrive(&tempty[as]); }
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 64) cycles[blockIdx.x] = t1 - t0;       // first epilogue warp
  if (threadIdx.x == 32) cycles[gridDim.x + blockIdx.x] = t1 - t0;  // MMA warp
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

template <int X, int OUT = 2, int NT = 64 + 16 * 32>
void run(const char* name, int mode, int nepi, int tiles, int katoms, int outstanding, int dcols = 256, int ns = 2) {
  const uint32_t idesc = dcols == 256 ? ptx::umma_idesc_f16(BM, BN) : (ptx::umma_idesc_f16(BM, BN) & ~(1u << 4));   // D = fp32 / fp16
  const int grid = 148;
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 2 * grid * sizeof(long long));
  cudaMalloc(&sink, 4096);
  const int smem = 4 * A_BYTES + 4 * B_BYTES;
  cudaFuncSetAttribute(bench_kernel<X, OUT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench_kernel<X, OUT, NT><<<grid, NT, smem>>>(mode, nepi, tiles / 4, katoms, outstanding, cyc, sink, dcols, ns, idesc);
  cudaEventRecord(e0);
  bench_kernel<X, OUT, NT><<<grid, NT, smem>>>(mode, nepi, tiles, katoms, outstanding, cyc, sink, dcols, ns, idesc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); exit(1); }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[2 * 148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double ce = 0, cm = 0;
  for (int i = 0; i < grid; ++i) { ce += h[i]; cm += h[grid + i]; }
  ce /= grid; cm /= grid;
  const double c = (mode & 1) ? ce : cm;
  printf("%-44s x%-2d nepi=%2d out=%d katoms=%d D=%s stages=%d: %8.1f cyc/tile  (%6.1f B/cyc TMEM->RF, %6.0f MAC/cyc)  %.3f ms -> %.0f MHz, %.0f TFLOP/s\n",
         name, X, nepi, outstanding, katoms, dcols == 256 ? "f32" : "f16", ns, c / tiles, (mode & 1) ? 512.0 * dcols * tiles / c : 0.0,
         (mode & 2) ? double(BM) * BN * 64 * katoms * tiles / c : 0.0, ms, c / ms / 1e3,
         (mode & 2) ? 2.0 * BM * BN * 64 * katoms * tiles * grid / ms / 1e9 : 0.0);
  cudaFree(cyc); cudaFree(sink);
}

int main(int argc, char** argv) {
  const int T = 4000;
  if (argc > 1 && argv[1][0] == 'h') {     // round 2: where do the ~150 cycles per tile between "mma only" and "pipelined" go?
    run<32>("mma only", 2, 16, T, 2, 2);
    run<32>("mma + commit per tile (no waits)", 2 | 4 | 16, 16, T, 2, 2);
    for (int wm : {0, 1, 2})
      for (int we : {0, 1, 2}) {
        cudaMemcpyToSymbol(g_wait_mma, &wm, 4);
        cudaMemcpyToSymbol(g_wait_epi, &we, 4);
        printf("-- wait flavour: MMA warp %d, epilogue warps %d (0 try_wait + 200 us hint, 1 test_wait spin, 2 try_wait, no hint)\n", wm, we);
        run<32>("pipelined, hand-off only (no loads)", 1 | 2 | 4 | 32, 16, T, 2, 2);
        run<32>("pipelined, all columns loaded + min-tree", 1 | 2 | 4 | 8, 8, T, 2, 2);
        run<32>("pipelined, all columns loaded + min-tree", 1 | 2 | 4 | 8, 16, T, 2, 2);
      }
    return 0;
  }
  for (int nepi : {4, 8, 16}) {
    run<32>("drain only", 1, nepi, T, 2, 1);
    run<32>("drain only", 1, nepi, T, 2, 2);
    run<16>("drain only", 1, nepi, T, 2, 2);
  }
  run<32>("drain + min-tree", 1 | 8, 16, T, 2, 2);
  run<32>("drain + min-tree", 1 | 8, 8, T, 2, 2);
  for (int ka : {1, 2, 4}) run<32>("mma only", 2, 16, T, ka, 2);
  for (int nepi : {8, 16}) {
    run<32>("mma + independent drain", 1 | 2, nepi, T, 2, 2);
    run<32>("mma -> drain pipelined", 1 | 2 | 4, nepi, T, 2, 2);
    run<32>("mma -> drain pipelined + min-tree", 1 | 2 | 4 | 8, nepi, T, 2, 2);
    run<32>("mma -> drain pipelined + min-tree", 1 | 2 | 4 | 8, nepi, T, 4, 2);
  }
  return 0;
}
