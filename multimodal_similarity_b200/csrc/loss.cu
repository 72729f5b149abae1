// K2/K3: batch-hard and lifted-structured loss, forward + backward in ONE cooperative launch.
//
// Reference semantics:
//   dists = cdist_tf(all_diffs_tf(E, E))                       src/utils.py:302-311,343-360
//   batch_hard(dists, pids, margin, weighted)                  src/networks.py:797-833
//   lifted_loss(dists, pids, margin, weighted)                 src/networks.py:835-870
// and the gradient TF autodiff produces for them.  The reference materialises [N,N,D]; here the N x N
// matrix lives only in registers:
//   phase 1  every CTA owns a TI x TJ tile of pairs: squared distances in difference form (the reference's
//            arithmetic, D_ii == 0 exactly), label masks, and the tile's contribution to each row reduction
//            (hardest positive / hardest negative with index and tie count, or online-logsumexp partials)
//   barrier  one grid-wide barrier (cooperative launch => all CTAs co-resident)
//   phase 2  row owners (one warp per row) combine the partials in a fixed order, write the reference's
//            per-row outputs and the batch-hard sparse gradient; tile CTAs of the lifted loss reuse the
//            distances still sitting in their registers for the dense gradient (G + G^T)(e_i - e_j)
//   last CTA sums the per-row loss terms in a fixed order (deterministic loss value).
// dE is accumulated with fp32 atomics (order-dependent in the last bits only).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "loss.h"

namespace mmsim {
namespace loss {

constexpr int THREADS = 128;

struct Params {
  const float* E;
  const float* pids;
  int N, D, kind, soft, weighted;
  float margin;
  float *loss, *num_active, *diff, *w, *fp, *cn;
  int *pos_idx, *neg_idx;
  float* dE;
  // workspace
  float *p_a, *p_b, *p_c, *p_d;   // partial tables [NBJ][Npad]
  int *p_ia, *p_ib, *p_ca, *p_cb; // partial index / tie-count tables
  int* same_cnt;                  // [N] number of same-label columns (incl. self), integer atomics
  float* row_loss;                // [N] l_i * w_i
  float* row_active;              // [N]
  unsigned int* sync;             // [0] barrier counter, [1] done counter
  unsigned long long* trace;      // [8] globaltimer stamps of CTA 0 at the phase boundaries (diagnostics, ~free)
  int NBI, NBJ, Npad;
};

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// hardest-positive style (max, smallest index among ties, tie count)
struct Arg {
  float v;
  int i, c;
};
__device__ __forceinline__ Arg arg_max(Arg a, Arg b) {
  if (a.v > b.v) return a;
  if (b.v > a.v) return b;
  return Arg{a.v, min(a.i, b.i) < 0 ? max(a.i, b.i) : min(a.i, b.i), a.c + b.c};
}
__device__ __forceinline__ Arg arg_min(Arg a, Arg b) {
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  return Arg{a.v, min(a.i, b.i) < 0 ? max(a.i, b.i) : min(a.i, b.i), a.c + b.c};
}
struct Lse {
  float m, s;
};
__device__ __forceinline__ Lse lse_add(Lse a, float x) {
  if (x <= a.m) {
    a.s += expf(x - a.m);
  } else {
    a.s = a.s * expf(a.m - x) + 1.f;  // a.m == -inf -> exp(-inf) = 0
    a.m = x;
  }
  return a;
}
__device__ __forceinline__ Lse lse_merge(Lse a, Lse b) {
  const float m = fmaxf(a.m, b.m);
  if (m == -kInf) return Lse{-kInf, 0.f};
  return Lse{m, a.s * expf(a.m - m) + b.s * expf(b.m - m)};
}
__device__ __forceinline__ float lse_value(Lse a) { return a.s > 0.f ? a.m + logf(a.s) : -kInf; }

__device__ __noinline__ void barrier_timeout() {
  printf("mmsim: loss grid barrier timed out\n");
  __trap();
}

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= expected) break;
      if (clock64() - t0 > 4000000000LL) barrier_timeout();
    }
  }
  __syncthreads();
}

// distance of row i to row j in the tile kernel's arithmetic (sequential fmaf over d) -- used by the tie path so
// that equality tests against the mined extreme reproduce bit for bit
__device__ __noinline__ float seq_sqdist(const float* __restrict__ a, const float* __restrict__ b, int D) {
  float acc = 0.f;
#pragma unroll 1
  for (int d = 0; d < D; ++d) {
    const float x = a[d] - b[d];
    acc = fmaf(x, x, acc);
  }
  return acc;
}

// Combine one row's NBJ online-logsumexp partials in the fixed order b = 0 .. NBJ-1.  Loads are issued eight partials
// at a time, ahead of the (sequential) merges, so a row costs NBJ/8 memory round trips instead of NBJ.
__device__ __forceinline__ void combine_lse_row(const Params& p, int i, Lse& lp, Lse& ln) {
  lp = Lse{-kInf, 0.f};
  ln = Lse{-kInf, 0.f};
  constexpr int B = 8;
  for (int b0 = 0; b0 < p.NBJ; b0 += B) {
    float va[B], vb[B], vc[B], vd[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const bool in = b0 + u < p.NBJ;
      const int slot = (in ? b0 + u : 0) * p.Npad + i;
      va[u] = in ? __ldcg(&p.p_a[slot]) : -kInf;
      vb[u] = in ? __ldcg(&p.p_b[slot]) : 0.f;
      vc[u] = in ? __ldcg(&p.p_c[slot]) : -kInf;
      vd[u] = in ? __ldcg(&p.p_d[slot]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      lp = lse_merge(lp, Lse{va[u], vb[u]});
      ln = lse_merge(ln, Lse{vc[u], vd[u]});
    }
  }
}

// Rare path of the batch-hard gradient, out of line: exact distance ties.  TF splits the gradient of reduce_max /
// reduce_min evenly over the tied entries (math_grad._MinOrMaxGrad).  One warp; distances are recomputed in the tile
// kernel's arithmetic so the equality tests against the mined extremes reproduce bit for bit.
__device__ __noinline__ void bh_tie_backward(const Params& p, int i, int lane, float c, float fp, float cn, bool use_pos,
                                             int n_pos_ties, int n_neg_ties) {
  const int N = p.N, D = p.D;
  const float* ei = p.E + size_t(i) * D;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    if (j == i) continue;
    const bool same_id = __ldcg(&p.pids[j]) == __ldcg(&p.pids[i]);
    float dist = 0.f;
    if (lane == 0) dist = seq_sqdist(ei, p.E + size_t(j) * D, D);
    dist = __shfl_sync(0xffffffffu, dist, 0);
    float g = 0.f;
    if (same_id && use_pos && dist == fp) g = c / float(n_pos_ties);
    if (!same_id && dist == cn) g = -c / float(n_neg_ties);
    if (g != 0.f) {
      const float* ej = p.E + size_t(j) * D;
#pragma unroll 1
      for (int d = lane; d < D; d += 32) {
        const float dv = 2.f * g * (ei[d] - ej[d]);
        atomicAdd(&p.dE[size_t(i) * D + d], dv);
        atomicAdd(&p.dE[size_t(j) * D + d], -dv);
      }
    }
  }
}

__device__ __forceinline__ void stamp(const Params& p, int slot) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

// KIND (0 batch-hard, 1 lifted) is a template parameter: the kernel runs once per SM with cold instruction caches, so
// dead code of the other loss is pure instruction-fetch latency.
template <int MI, int MJ, int KIND>
__global__ void __launch_bounds__(THREADS) loss_kernel(const Params p) {
  constexpr int TI = 8 * MI, TJ = 16 * MJ;
  extern __shared__ __align__(16) float sm[];
  const int N = p.N, D = p.D;
  const int D4 = (D + 3) & ~3;
  const int DP = D4 + 4;
  float* Ei = sm;                 // [TI][DP]
  float* Ej = Ei + TI * DP;       // [TJ][DP]
  float* pid_i = Ej + TJ * DP;    // [TI]
  float* pid_j = pid_i + TI;      // [TJ]
  float* st = pid_j + TJ;         // lifted: per-row stats for TI + TJ rows: fp, cn, coef -> 3 * (TI + TJ)
  float* Ss = st + 3 * (TI + TJ); // lifted: S tile, transposed [TJ][TI + 4] (16-byte aligned rows)
  __shared__ int s_red[THREADS / 32];
  __shared__ int s_W;
  __shared__ int s_last;

  const int t = threadIdx.x;
  const int bi = blockIdx.x / p.NBJ, bj = blockIdx.x % p.NBJ;
  const int i0 = bi * TI, j0 = bj * TJ;
  const int ti = t >> 4, tj = t & 15;
  const unsigned int G = gridDim.x;

  stamp(p, 0);
  // ---- zero this CTA's slice of dE (ordered before phase 2 by the grid barrier)
  if (p.dE) {
    const int64_t total = int64_t(N) * D;
    for (int64_t x = int64_t(blockIdx.x) * THREADS + t; x < total; x += int64_t(G) * THREADS) p.dE[x] = 0.f;
  }

  // ---- stage the two row panels (coalesced, zero padded); 128-bit loads when the rows are 16-byte aligned
  {
    const int Q4 = D4 >> 2;                       // float4 slots per row
    const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0);
    // all loads of a batch are issued before the first store so that they overlap (one memory round trip per batch)
    const int total4 = (TI + TJ) * Q4;
    constexpr int UNR = 4;
    for (int base = t; base < total4; base += THREADS * UNR) {
      float4 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int x = base + u * THREADS;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < total4) {
          const int r = x / Q4, c4 = x - r * Q4;
          const int gr = r < TI ? i0 + r : j0 + (r - TI);
          if (gr < N) {
            const float* src = p.E + size_t(gr) * D + c4 * 4;
            if (vec) {
              v[u] = *reinterpret_cast<const float4*>(src);
            } else {
              const int c = c4 * 4;
              v[u].x = c + 0 < D ? src[0] : 0.f; v[u].y = c + 1 < D ? src[1] : 0.f;
              v[u].z = c + 2 < D ? src[2] : 0.f; v[u].w = c + 3 < D ? src[3] : 0.f;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int x = base + u * THREADS;
        if (x < total4) {
          const int r = x / Q4, c4 = x - r * Q4;
          float* dst = (r < TI ? Ei + r * DP : Ej + (r - TI) * DP) + c4 * 4;
          *reinterpret_cast<float4*>(dst) = v[u];
        }
      }
    }
  }
  if (t < TI) pid_i[t] = i0 + t < N ? p.pids[i0 + t] : 0.f;
  if (t < TJ) pid_j[t] = j0 + t < N ? p.pids[j0 + t] : 0.f;
  __syncthreads();

  stamp(p, 1);
  // ---- phase 1a: MI x MJ squared distances per thread, difference form
  float acc[MI][MJ];
#pragma unroll
  for (int a = 0; a < MI; ++a)
#pragma unroll
    for (int b = 0; b < MJ; ++b) acc[a][b] = 0.f;
  for (int d = 0; d < D4; d += 4) {
    float4 av[MI], bv[MJ];
#pragma unroll
    for (int a = 0; a < MI; ++a) av[a] = *reinterpret_cast<const float4*>(Ei + (ti + 8 * a) * DP + d);
#pragma unroll
    for (int b = 0; b < MJ; ++b) bv[b] = *reinterpret_cast<const float4*>(Ej + (tj + 16 * b) * DP + d);
#pragma unroll
    for (int a = 0; a < MI; ++a)
#pragma unroll
      for (int b = 0; b < MJ; ++b) {
        float x;
        x = av[a].x - bv[b].x; acc[a][b] = fmaf(x, x, acc[a][b]);
        x = av[a].y - bv[b].y; acc[a][b] = fmaf(x, x, acc[a][b]);
        x = av[a].z - bv[b].z; acc[a][b] = fmaf(x, x, acc[a][b]);
        x = av[a].w - bv[b].w; acc[a][b] = fmaf(x, x, acc[a][b]);
      }
  }

  stamp(p, 2);
  // ---- phase 1b: masked row reductions over this tile's columns
#pragma unroll
  for (int a = 0; a < MI; ++a) {
    const int li = ti + 8 * a, i = i0 + li;
    const float pi = pid_i[li];
    int same = 0;
    Arg hp{-1.f, -1, 0}, hn{kInf, -1, 0};
    Lse lp{-kInf, 0.f}, ln{-kInf, 0.f};
#pragma unroll
    for (int b = 0; b < MJ; ++b) {
      const int lj = tj + 16 * b, j = j0 + lj;
      if (j >= N || i >= N) continue;
      const bool same_id = pid_j[lj] == pi;
      const bool pos = same_id && (i != j);
      const float dist = acc[a][b];
      same += same_id ? 1 : 0;
      if (KIND == 0) {
        if (pos) hp = arg_max(hp, Arg{dist, j, 1});
        if (!same_id) hn = arg_min(hn, Arg{dist, j, 1});
      } else {
        lp = lse_add(lp, pos ? dist : 0.f);              // over ALL columns (networks.py:846)
        if (!same_id) ln = lse_add(ln, p.margin - dist);  // negatives only (:847-848)
      }
    }
    // reduce over the 16 threads that share this row (xor offsets < 16 stay inside the half-warp)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      same += __shfl_xor_sync(0xffffffffu, same, o);
      if (KIND == 0) {
        Arg o1{__shfl_xor_sync(0xffffffffu, hp.v, o), __shfl_xor_sync(0xffffffffu, hp.i, o), __shfl_xor_sync(0xffffffffu, hp.c, o)};
        Arg o2{__shfl_xor_sync(0xffffffffu, hn.v, o), __shfl_xor_sync(0xffffffffu, hn.i, o), __shfl_xor_sync(0xffffffffu, hn.c, o)};
        hp = arg_max(hp, o1);
        hn = arg_min(hn, o2);
      } else {
        Lse o1{__shfl_xor_sync(0xffffffffu, lp.m, o), __shfl_xor_sync(0xffffffffu, lp.s, o)};
        Lse o2{__shfl_xor_sync(0xffffffffu, ln.m, o), __shfl_xor_sync(0xffffffffu, ln.s, o)};
        lp = lse_merge(lp, o1);
        ln = lse_merge(ln, o2);
      }
    }
    if (tj == 0 && i < N) {
      const int slot = bj * p.Npad + i;
      if (KIND == 0) {
        p.p_a[slot] = hp.v; p.p_ia[slot] = hp.i; p.p_ca[slot] = hp.c;
        p.p_b[slot] = hn.v; p.p_ib[slot] = hn.i; p.p_cb[slot] = hn.c;
      } else {
        p.p_a[slot] = lp.m; p.p_b[slot] = lp.s;
        p.p_c[slot] = ln.m; p.p_d[slot] = ln.s;
      }
      if (same) atomicAdd(&p.same_cnt[i], same);
    }
  }

  stamp(p, 3);
  grid_barrier(&p.sync[0], G);
  stamp(p, 4);

  // ---- phase 2a: W = sum_i (#negatives_i) * [pid_i != 0], exact in integers (every CTA, fixed order)
  {
    int part = 0;
    for (int i = t; i < N; i += THREADS) part += (__ldcg(&p.pids[i]) != 0.f) ? (N - __ldcg(&p.same_cnt[i])) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((t & 31) == 0) s_red[t >> 5] = part;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int x = 0; x < THREADS / 32; ++x) tot += s_red[x];
      s_W = tot;
    }
    __syncthreads();
  }
  const float Wf = float(s_W);
  const float invN = 1.0f / float(N);
  auto weight_of = [&](int i) -> float {
    if (!p.weighted) return invN;
    const float wn = (__ldcg(&p.pids[i]) != 0.f) ? float(N - __ldcg(&p.same_cnt[i])) : 0.f;
    return wn / Wf;
  };

  stamp(p, 5);
  // ---- phase 2b: row owners -- one warp per row, fixed combine order over the NBJ partials
  {
    const int warp = t >> 5, lane = t & 31;
    for (int i = blockIdx.x * (THREADS / 32) + warp; i < N; i += G * (THREADS / 32)) {
      const float wi = weight_of(i);
      const float fg = __ldcg(&p.pids[i]) != 0.f ? 1.f : 0.f;
      if (KIND == 0) {
        Arg hp{-1.f, -1, 0}, hn{kInf, -1, 0};
        if (lane < p.NBJ) {
          const int slot = lane * p.Npad + i;
          hp = Arg{__ldcg(&p.p_a[slot]), __ldcg(&p.p_ia[slot]), __ldcg(&p.p_ca[slot])};
          hn = Arg{__ldcg(&p.p_b[slot]), __ldcg(&p.p_ib[slot]), __ldcg(&p.p_cb[slot])};
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Arg o1{__shfl_xor_sync(0xffffffffu, hp.v, o), __shfl_xor_sync(0xffffffffu, hp.i, o), __shfl_xor_sync(0xffffffffu, hp.c, o)};
          Arg o2{__shfl_xor_sync(0xffffffffu, hn.v, o), __shfl_xor_sync(0xffffffffu, hn.i, o), __shfl_xor_sync(0xffffffffu, hn.c, o)};
          hp = arg_max(hp, o1);
          hn = arg_min(hn, o2);
        }
        const float fp = fmaxf(hp.v, 0.f);  // max_j(D_ij * pos_ij) >= 0: masked zeros take part (networks.py:808)
        const float cn = hn.v;              // +inf when the row has no negatives
        const float x = fp - cn;
        float l, dl;
        if (p.soft) {
          l = softplus_f(x);
          dl = 1.f / (1.f + expf(-x));
        } else {
          l = fmaxf(x + p.margin, 0.f);
          dl = (x + p.margin >= 0.f) ? 1.f : 0.f;
        }
        if (hn.i < 0) { l = 0.f; dl = 0.f; }  // empty negative set: the term is exactly 0
        if (lane == 0) {
          p.diff[i] = l; p.w[i] = wi; p.fp[i] = fp; p.cn[i] = cn;
          p.pos_idx[i] = hp.i; p.neg_idx[i] = hn.i;
          p.row_loss[i] = l * wi;
          p.row_active[i] = (l * fg > 1e-5f) ? 1.f : 0.f;
        }
        // sparse gradient: dL/dD_ip = +c/|P|, dL/dD_in = -c/|N| ; dD_ij/de_i = 2(e_i - e_j)
        const float c = wi * dl;
        if (p.dE && c != 0.f) {
          const float* ei = p.E + size_t(i) * D;
          const bool use_pos = hp.i >= 0 && fp > 0.f;  // fp == 0: every tied entry has zero derivative
          if (hp.c <= 1 && hn.c <= 1) {
            const float* ep = p.E + size_t(use_pos ? hp.i : i) * D;
            const float* en = p.E + size_t(hn.i) * D;
            for (int d = lane; d < D; d += 32) {
              const float vi = ei[d], vp = ep[d], vn = en[d];
              float gi = -2.f * c * (vi - vn);
              if (use_pos) {
                gi += 2.f * c * (vi - vp);
                atomicAdd(&p.dE[size_t(hp.i) * D + d], 2.f * c * (vp - vi));
              }
              atomicAdd(&p.dE[size_t(i) * D + d], gi);
              atomicAdd(&p.dE[size_t(hn.i) * D + d], 2.f * c * (vi - vn));
            }
          } else {
            bh_tie_backward(p, i, lane, c, fp, cn, use_pos, hp.c, hn.c);
          }
        }
      } else {
        Lse lp, ln;
        if (lane == 0) {
          combine_lse_row(p, i, lp, ln);     // same fixed order as the tile CTAs below: identical bits
          const float fp = lse_value(lp), cn = lse_value(ln);
          const float l = (cn > -kInf) ? fmaxf(fp + cn, 0.f) : 0.f;
          p.diff[i] = l; p.w[i] = wi; p.fp[i] = fp; p.cn[i] = cn;
          p.pos_idx[i] = -1; p.neg_idx[i] = -1;
          p.row_loss[i] = l * wi;
          p.row_active[i] = 1.f;
        }
      }
    }
  }

  // ---- phase 2c (lifted): dense gradient from the distances still held in registers
  if (KIND == 1 && p.dE) {
    // stats of the TI rows (as anchors) and the TJ rows (as anchors of the transposed entries)
    for (int r = t; r < TI + TJ; r += THREADS) {
      const int i = r < TI ? i0 + r : j0 + (r - TI);
      float fp = 0.f, cn = -kInf, coef = 0.f;
      if (i < N) {
        Lse lp, ln;
        combine_lse_row(p, i, lp, ln);
        fp = lse_value(lp);
        cn = lse_value(ln);
        coef = (cn > -kInf && fp + cn >= 0.f) ? weight_of(i) : 0.f;  // w_i * [l_i active]
      }
      st[r] = fp;
      st[(TI + TJ) + r] = cn;
      st[2 * (TI + TJ) + r] = coef;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < MI; ++a) {
      const int li = ti + 8 * a, i = i0 + li;
#pragma unroll
      for (int b = 0; b < MJ; ++b) {
        const int lj = tj + 16 * b, j = j0 + lj;
        float s = 0.f;
        if (i < N && j < N && i != j) {
          const bool same_id = pid_j[lj] == pid_i[li];
          const float dist = acc[a][b];
          const float ci = st[2 * (TI + TJ) + li], cj = st[2 * (TI + TJ) + TI + lj];
          if (same_id) {
            // G_ij = c_i exp(D_ij - fp_i), G_ji = c_j exp(D_ij - fp_j)
            if (ci != 0.f) s += ci * expf(dist - st[li]);
            if (cj != 0.f) s += cj * expf(dist - st[TI + lj]);
          } else {
            if (ci != 0.f) s -= ci * expf(p.margin - dist - st[(TI + TJ) + li]);
            if (cj != 0.f) s -= cj * expf(p.margin - dist - st[(TI + TJ) + TI + lj]);
          }
        }
        Ss[lj * (TI + 4) + li] = s;                 // transposed: the gradient loop reads four rows of a column at once
      }
    }
    __syncthreads();
    // dE[i][d] += 2 * sum_j S_ij (e_i[d] - e_j[d]).  A thread owns one feature d and FOUR rows at a time, so every e_j[d]
    // it loads from shared memory is used four times (the one-row form did two shared loads per FMA and the loop was
    // shared-memory-issue bound).
    static_assert(TI % 4 == 0, "rows are processed four at a time");
    for (int x = t; x < (TI / 4) * D; x += THREADS) {
      const int lq = x / D, d = x - lq * D, li0 = lq * 4;
      float ei[4], o[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { ei[a] = Ei[(li0 + a) * DP + d]; o[a] = 0.f; }
#pragma unroll 4
      for (int lj = 0; lj < TJ; ++lj) {
        const float ej = Ej[lj * DP + d];
        const float4 s4 = *reinterpret_cast<const float4*>(Ss + lj * (TI + 4) + li0);   // one broadcast load
        o[0] = fmaf(s4.x, ei[0] - ej, o[0]);
        o[1] = fmaf(s4.y, ei[1] - ej, o[1]);
        o[2] = fmaf(s4.z, ei[2] - ej, o[2]);
        o[3] = fmaf(s4.w, ei[3] - ej, o[3]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (i0 + li0 + a < N && o[a] != 0.f) atomicAdd(&p.dE[size_t(i0 + li0 + a) * D + d], 2.f * o[a]);
    }
  }

  stamp(p, 6);
  // ---- last CTA: deterministic sum of the per-row terms
  __threadfence();
  __syncthreads();
  if (t == 0) {
    __threadfence();
    s_last = (atomicAdd(&p.sync[1], 1u) == G - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    __shared__ float s_l[THREADS], s_a[THREADS], s_f[THREADS];
    float l = 0.f, a = 0.f, f = 0.f;
    for (int i = t; i < N; i += THREADS) {
      l += __ldcg(&p.row_loss[i]);
      a += __ldcg(&p.row_active[i]);
      f += __ldcg(&p.pids[i]) != 0.f ? 1.f : 0.f;
    }
    s_l[t] = l; s_a[t] = a; s_f[t] = f;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
      if (t < o) { s_l[t] += s_l[t + o]; s_a[t] += s_a[t + o]; s_f[t] += s_f[t + o]; }
      __syncthreads();
    }
    if (t == 0) {
      *p.loss = s_l[0];
      *p.num_active = KIND == 0 ? s_a[0] / s_f[0] : 1.0f;
      p.sync[0] = 0;   // every CTA has passed the barrier and finished phase 2: leave the workspace ready for the
      p.sync[1] = 0;   // next launch (contract: zero-filled before the first call, left zero-filled by every call)
    }
    for (int i = t; i < N; i += THREADS) p.same_cnt[i] = 0;
    if (t == 0) {
      unsigned long long tt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      p.trace[7] = tt;
    }
  }
}

// ------------------------------------------------------------------------------------------------ batch-hard, row-owned
// K2': batch-hard forward + backward WITHOUT a grid barrier.  The tile kernel above spends most of its 19 us on L2 round
// trips, not arithmetic (gpurun_out/loss_trace.log: staging 3.6, distances 1.2, partials 1.4, grid barrier 2.0, W 1.3,
// row combine + gradient 6.2, last CTA 3.7).  Batch-hard needs none of the exchanges: if a CTA owns COMPLETE rows -- 4
// anchors against all N columns, streamed through shared memory in 128-row tiles with cp.async double buffering (N/4
// CTAs of 128 threads; fewer rows per CTA and the L2 -> SM traffic of every CTA streaming all of E dominates) -- the
// hardest positive / negative of its rows are final inside the CTA, the weight normaliser W depends on the labels only
// (every CTA counts it itself while the first tile is in flight), and the sparse gradient of an anchor touches rows
// i, p, n with a coefficient that depends on row i alone.  What is left of the global synchronisation is the
// deterministic sum of the per-row terms by the last CTA.  dE is zeroed by a memset node ahead of the launch (the
// contributions arrive as atomics from arbitrary CTAs).  Distances use the tile kernel's arithmetic (sequential fmaf
// over d), so values, mined indices and the tie path are bit-identical to it.
constexpr int BR_THREADS = 128, BR_R = 4, BR_TJ = 128;
constexpr int BR_HASH = 2048;              // label hash table slots (power of two, >= 2 N)

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int n = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__global__ void __launch_bounds__(BR_THREADS) bh_rows_kernel(const Params p) {
  extern __shared__ __align__(16) float sm[];
  const int N = p.N, D = p.D;
  const int D4 = (D + 3) & ~3, DP = D4 + 4, Q4 = D4 >> 2;
  float* Ei = sm;                              // [BR_R][DP]   this CTA's anchors
  float* Ej = Ei + BR_R * DP;                  // [2][BR_TJ][DP] streamed column tiles
  float* pid_all = Ej + 2 * BR_TJ * DP;        // [Npad]
  int* same_all = reinterpret_cast<int*>(pid_all + p.Npad);   // [Npad] same-label columns of every row (incl. self)
  float* part = reinterpret_cast<float*>(same_all + p.Npad);  // [BR_R][warps][6] partial (hp, hn) of a row
  int* hkey = reinterpret_cast<int*>(part + BR_R * (BR_THREADS / 32) * 6);   // [BR_HASH] label hash table
  int* hcnt = hkey + BR_HASH;
  __shared__ int s_red[BR_THREADS / 32];
  __shared__ int s_W, s_last;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int i0 = blockIdx.x * BR_R;
  const unsigned int G = gridDim.x;
  const int n_tiles = (N + BR_TJ - 1) / BR_TJ;
  const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0);

  stamp(p, 0);
  auto load_tile = [&](int tile, int buf) {
    float* dst = Ej + buf * BR_TJ * DP;
    const int j0 = tile * BR_TJ;
    if (vec) {
      for (int x = t; x < BR_TJ * Q4; x += BR_THREADS) {
        const int r = x / Q4, c4 = x - r * Q4;
        const bool in = j0 + r < N;
        cp_async16(dst + r * DP + c4 * 4, p.E + size_t(in ? j0 + r : 0) * D + c4 * 4, in);
      }
    } else {
      for (int x = t; x < BR_TJ * D4; x += BR_THREADS) {
        const int r = x / D4, c = x - r * D4;
        dst[r * DP + c] = (j0 + r < N && c < D) ? p.E[size_t(j0 + r) * D + c] : 0.f;
      }
    }
    cp_async_commit();
  };
  load_tile(0, 0);
  // anchors + labels (plain loads, overlapped with the first tile)
  for (int x = t; x < BR_R * D4; x += BR_THREADS) {
    const int r = x / D4, c = x - r * D4;
    Ei[r * DP + c] = (i0 + r < N && c < D) ? p.E[size_t(i0 + r) * D + c] : 0.f;
  }
  for (int i = t; i < p.Npad; i += BR_THREADS) pid_all[i] = i < N ? p.pids[i] : 0.f;
  __syncthreads();
  // same-label counts of every row (a shared-memory hash table keyed by the label's bits: O(N) instead of N^2 compares,
  // which alone took 6 us) and W = sum_i (#negatives_i) * [pid_i != 0], exact in integers
  {
    for (int x = t; x < BR_HASH; x += BR_THREADS) { hkey[x] = 0x7fc00123; hcnt[x] = 0; }   // empty = an unused NaN pattern
    __syncthreads();
    for (int i = t; i < N; i += BR_THREADS) {
      const int key = __float_as_int(pid_all[i] + 0.f);          // -0 and +0 are the same label
      int slot = (unsigned(key) * 2654435761u >> 16) & (BR_HASH - 1);
      for (;;) {
        const int prev = atomicCAS(&hkey[slot], 0x7fc00123, key);
        if (prev == 0x7fc00123 || prev == key) break;
        slot = (slot + 1) & (BR_HASH - 1);
      }
      atomicAdd(&hcnt[slot], 1);
      same_all[i] = slot;
    }
    __syncthreads();
    int part_w = 0;
    for (int i = t; i < N; i += BR_THREADS) {
      const int c = hcnt[same_all[i]];
      part_w += pid_all[i] != 0.f ? N - c : 0;
      same_all[i] = c;      // only this thread reads and writes entry i
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part_w += __shfl_xor_sync(0xffffffffu, part_w, o);
    if (lane == 0) s_red[warp] = part_w;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int x = 0; x < BR_THREADS / 32; ++x) tot += s_red[x];
      s_W = tot;
    }
  }
  stamp(p, 1);

  // ---- distances + masked row extremes: a thread owns one column of every tile for the CTA's four anchors (the column
  // values are read once for all of them; four independent fmaf chains)
  float pr[BR_R];
  Arg hp4[BR_R], hn4[BR_R];
#pragma unroll
  for (int r = 0; r < BR_R; ++r) {
    pr[r] = i0 + r < N ? pid_all[i0 + r] : 0.f;
    hp4[r] = Arg{-1.f, -1, 0};
    hn4[r] = Arg{kInf, -1, 0};
  }
  for (int tile = 0; tile < n_tiles; ++tile) {
    if (tile + 1 < n_tiles) {
      load_tile(tile + 1, (tile + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* ej = Ej + (tile & 1) * BR_TJ * DP + t * DP;
    float dd[BR_R];
#pragma unroll
    for (int r = 0; r < BR_R; ++r) dd[r] = 0.f;
#pragma unroll 2
    for (int d = 0; d < D4; d += 4) {
      const float4 v = *reinterpret_cast<const float4*>(ej + d);
#pragma unroll
      for (int r = 0; r < BR_R; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(Ei + r * DP + d);
        float x;
        x = a.x - v.x; dd[r] = fmaf(x, x, dd[r]);
        x = a.y - v.y; dd[r] = fmaf(x, x, dd[r]);
        x = a.z - v.z; dd[r] = fmaf(x, x, dd[r]);
        x = a.w - v.w; dd[r] = fmaf(x, x, dd[r]);
      }
    }
    const int j = tile * BR_TJ + t;
    if (j < N) {
      const float pj = pid_all[j];
#pragma unroll
      for (int r = 0; r < BR_R; ++r) {
        if (i0 + r >= N) continue;
        if (pj == pr[r]) { if (j != i0 + r) hp4[r] = arg_max(hp4[r], Arg{dd[r], j, 1}); } else hn4[r] = arg_min(hn4[r], Arg{dd[r], j, 1});
      }
    }
    __syncthreads();   // the tile buffer is reloaded two iterations later
  }
  stamp(p, 2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    auto sh = [&](Arg x) { return Arg{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o), __shfl_xor_sync(0xffffffffu, x.c, o)}; };
#pragma unroll
    for (int r = 0; r < BR_R; ++r) {
      hp4[r] = arg_max(hp4[r], sh(hp4[r]));
      hn4[r] = arg_min(hn4[r], sh(hn4[r]));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < BR_R; ++r) {
      float* q = part + (r * (BR_THREADS / 32) + warp) * 6;
      q[0] = hp4[r].v; q[1] = __int_as_float(hp4[r].i); q[2] = __int_as_float(hp4[r].c);
      q[3] = hn4[r].v; q[4] = __int_as_float(hn4[r].i); q[5] = __int_as_float(hn4[r].c);
    }
  }
  __syncthreads();
  stamp(p, 3);
  stamp(p, 4);

  // ---- per-row results and the sparse gradient: warp w owns anchor i0 + w
  const float Wf = float(s_W), invN = 1.0f / float(N);
  {
    const int i = i0 + warp;
    if (warp < BR_R && i < N) {
      Arg hp{-1.f, -1, 0}, hn{kInf, -1, 0};
#pragma unroll
      for (int w2 = 0; w2 < BR_THREADS / 32; ++w2) {
        const float* q = part + (warp * (BR_THREADS / 32) + w2) * 6;
        hp = arg_max(hp, Arg{q[0], __float_as_int(q[1]), __float_as_int(q[2])});
        hn = arg_min(hn, Arg{q[3], __float_as_int(q[4]), __float_as_int(q[5])});
      }
      const float pi = pid_all[i];
      const float fg = pi != 0.f ? 1.f : 0.f;
      const float wi = p.weighted ? (pi != 0.f ? float(N - same_all[i]) : 0.f) / Wf : invN;
      const float fp = fmaxf(hp.v, 0.f);  // max_j(D_ij * pos_ij) >= 0: masked zeros take part (networks.py:808)
      const float cn = hn.v;              // +inf when the row has no negatives
      const float x = fp - cn;
      float l, dl;
      if (p.soft) {
        l = softplus_f(x);
        dl = 1.f / (1.f + expf(-x));
      } else {
        l = fmaxf(x + p.margin, 0.f);
        dl = (x + p.margin >= 0.f) ? 1.f : 0.f;
      }
      if (hn.i < 0) { l = 0.f; dl = 0.f; }  // empty negative set: the term is exactly 0
      if (lane == 0) {
        p.diff[i] = l; p.w[i] = wi; p.fp[i] = fp; p.cn[i] = cn;
        p.pos_idx[i] = hp.i; p.neg_idx[i] = hn.i;
        p.row_loss[i] = l * wi;
        p.row_active[i] = (l * fg > 1e-5f) ? 1.f : 0.f;
      }
      // sparse gradient: dL/dD_ip = +c/|P|, dL/dD_in = -c/|N| ; dD_ij/de_i = 2(e_i - e_j)
      const float c = wi * dl;
      if (p.dE && c != 0.f) {
        const bool use_pos = hp.i >= 0 && fp > 0.f;  // fp == 0: every tied entry has zero derivative
        if (hp.c <= 1 && hn.c <= 1) {
          const float* ei = Ei + warp * DP;
          const float* ep = p.E + size_t(use_pos ? hp.i : i) * D;
          const float* en = p.E + size_t(hn.i) * D;
          for (int d = lane; d < D; d += 32) {
            const float vi = ei[d], vp = ep[d], vn = en[d];
            float gi = -2.f * c * (vi - vn);
            if (use_pos) {
              gi += 2.f * c * (vi - vp);
              atomicAdd(&p.dE[size_t(hp.i) * D + d], 2.f * c * (vp - vi));
            }
            atomicAdd(&p.dE[size_t(i) * D + d], gi);
            atomicAdd(&p.dE[size_t(hn.i) * D + d], 2.f * c * (vi - vn));
          }
        } else {
          bh_tie_backward(p, i, lane, c, fp, cn, use_pos, hp.c, hn.c);
        }
      }
    }
  }
  stamp(p, 5);
  stamp(p, 6);
  // ---- last CTA: deterministic sum of the per-row terms
  __threadfence();
  __syncthreads();
  if (t == 0) {
    __threadfence();
    s_last = (atomicAdd(&p.sync[1], 1u) == G - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    __shared__ float s_l[BR_THREADS], s_a[BR_THREADS], s_f[BR_THREADS];
    float l = 0.f, a = 0.f, f = 0.f;
    for (int i = t; i < N; i += BR_THREADS) {
      l += __ldcg(&p.row_loss[i]);
      a += __ldcg(&p.row_active[i]);
      f += pid_all[i] != 0.f ? 1.f : 0.f;
    }
    s_l[t] = l; s_a[t] = a; s_f[t] = f;
    __syncthreads();
    for (int o = BR_THREADS / 2; o > 0; o >>= 1) {
      if (t < o) { s_l[t] += s_l[t + o]; s_a[t] += s_a[t + o]; s_f[t] += s_f[t + o]; }
      __syncthreads();
    }
    if (t == 0) {
      *p.loss = s_l[0];
      *p.num_active = s_a[0] / s_f[0];
      p.sync[1] = 0;   // leave the workspace zero-filled for the next launch
      unsigned long long tt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      p.trace[7] = tt;
    }
  }
}

static size_t bh_rows_smem(int64_t N, int64_t D) {
  const size_t D4 = size_t((D + 3) & ~int64_t(3)), DP = D4 + 4, Npad = align_up(size_t(N), 32);
  return (BR_R * DP + 2 * BR_TJ * DP + 2 * Npad + BR_R * (BR_THREADS / 32) * 6 + 2 * BR_HASH) * 4;
}
// MMSIM_BH_ROWS=1 selects the row-owned kernel for batch-hard (A/B runs).  It is NOT the default: measured on B200 at
// 256 x 128 it is no faster than the tile kernel (18.0 vs 19.3 us in-kernel, 20.7 vs 21.8 us per CUDA-graph replay, and
// slower per eager call because of the extra memset launch; gpurun_out/loss_trace4.log) -- every CTA streaming all of E
// through L2 costs what the grid barrier and the partial tables cost the tile kernel.
static bool bh_rows_ok(int64_t N, int64_t D) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MMSIM_BH_ROWS");
    enabled = (e && atoi(e) == 1) ? 1 : 0;
  }
  return enabled && bh_rows_smem(N, D) <= 200 * 1024;
}

// ------------------------------------------------------------------------------------------------ host side
// Tile shapes (MI, MJ) -> TI = 8 MI rows x TJ = 16 MJ columns per CTA, smallest first.
static const int kShapes[3][2] = {{2, 2}, {4, 4}, {8, 8}};

static size_t smem_for(int TI, int TJ, int64_t D) {
  const int D4 = int((D + 3) & ~int64_t(3));
  return size_t(TI + TJ) * (D4 + 4) * 4 + size_t(TI + TJ) * 4 + size_t(3) * (TI + TJ) * 4 + size_t(TJ) * (TI + 4) * 4 + 16;
}

static const void* kernel_for(int shape, int kind) {
  switch (shape * 2 + kind) {
    case 0: return reinterpret_cast<const void*>(loss_kernel<2, 2, 0>);
    case 1: return reinterpret_cast<const void*>(loss_kernel<2, 2, 1>);
    case 2: return reinterpret_cast<const void*>(loss_kernel<4, 4, 0>);
    case 3: return reinterpret_cast<const void*>(loss_kernel<4, 4, 1>);
    case 4: return reinterpret_cast<const void*>(loss_kernel<8, 8, 0>);
    default: return reinterpret_cast<const void*>(loss_kernel<8, 8, 1>);
  }
}

// The dynamic shared-memory limit is an attribute of the KERNEL, while the amount a launch needs depends on (N, D): the
// attribute is only ever raised (a later, smaller problem must not lower it under a layout that is already cached --
// that made a cached 256 x 128 launch fail with "invalid argument" after a 64 x 16 one in the same process).
// It is also an attribute PER DEVICE (the current one when it is set), and a process may drive several GPUs: everything
// cached here -- the raised limits, the occupancy-derived tile shape, the SM count -- is keyed by the device ordinal
// (round-1 advisor finding: the second GPU of a process never had its limit raised).  Slot kMaxDev = "no device"
// (workspace-size queries on a CPU box).
constexpr int kMaxDev = 64;
static int current_dev_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return kMaxDev;
  }
  return dev >= 0 && dev < kMaxDev ? dev : kMaxDev;
}

static void raise_smem_attr(int shape, int kind, size_t bytes) {
  static std::mutex mu;
  static size_t have[kMaxDev + 1][3][2] = {};
  const int slot = current_dev_slot();
  std::lock_guard<std::mutex> lock(mu);
  if (slot == kMaxDev) return;
  if (bytes > have[slot][shape][kind] &&
      cudaFuncSetAttribute(kernel_for(shape, kind), cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) == cudaSuccess)
    have[slot][shape][kind] = bytes;
}

// Pick the smallest tile whose grid fits co-resident on the device (a cooperative launch requires it).
static int pick_shape(int64_t N, int64_t D) {
  if (const char* e = getenv("MMSIM_LOSS_SHAPE")) {   // experiment switch: force tile shape 0 / 1 / 2
    const int s = atoi(e);
    if (s >= 0 && s <= 2) return s;
  }
  int num_sms = 0;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      num_sms = 0;
    }
  }
  for (int s = 0; s < 3; ++s) {
    const int TI = 8 * kShapes[s][0], TJ = 16 * kShapes[s][1];
    const size_t smem = smem_for(TI, TJ, D);
    if (smem > 200 * 1024) continue;
    const int64_t grid = ((N + TI - 1) / TI) * ((N + TJ - 1) / TJ);
    int per_sm = 0;
    if (num_sms > 0) {
      int per_kind[2] = {0, 0};
      for (int kd = 0; kd < 2; ++kd) {
        raise_smem_attr(s, kd, smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_kind[kd], kernel_for(s, kd), THREADS, smem) != cudaSuccess) per_kind[kd] = 0;
      }
      per_sm = per_kind[0] < per_kind[1] ? per_kind[0] : per_kind[1];
    } else {
      per_sm = int((227 * 1024) / (smem + 1024));  // no device (layout queries on a CPU box): assume 148 SMs
    }
    const int64_t cap = int64_t(per_sm) * (num_sms > 0 ? num_sms : 148);
    if (grid <= cap) return s;
  }
  return -1;
}

static Layout compute_layout(int64_t N, int64_t D) {
  Layout L{};
  L.shape = pick_shape(N, D);
  const int s = L.shape < 0 ? 2 : L.shape;
  L.TI = 8 * kShapes[s][0];
  L.TJ = 16 * kShapes[s][1];
  L.NBI = int((N + L.TI - 1) / L.TI);
  L.NBJ = int((N + L.TJ - 1) / L.TJ);
  L.Npad = int(align_up(size_t(N), 32));
  const size_t tab = size_t((N + 31) / 32) * L.Npad * 4;  // sized for the smallest TJ so the size is device independent
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  L.off_sync = take(256);
  L.off_same = take(size_t(N) * 4);
  L.zero_bytes = off;  // [sync | same_cnt] must be zeroed before every launch
  for (int x = 0; x < 8; ++x) L.off_tab[x] = take(tab);
  L.off_row_loss = take(size_t(N) * 4);
  L.off_row_active = take(size_t(N) * 4);
  L.off_trace = take(64);
  L.total_bytes = off;
  L.smem_bytes = smem_for(L.TI, L.TJ, D);
  return L;
}

// The layout (tile shape by occupancy query, shared-memory attribute) is computed once per (N, D): the losses are
// microsecond-scale, so the per-call host path must stay a memset + one launch.
Layout make_layout(int64_t N, int64_t D) {
  static std::mutex mu;
  static std::unordered_map<int64_t, Layout> cache;
  const int64_t key = (int64_t(current_dev_slot()) << 40) | (N << 20) | D;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  Layout L = compute_layout(N, D);
  if (L.shape >= 0)
    for (int kd = 0; kd < 2; ++kd)
      raise_smem_attr(L.shape, kd, L.smem_bytes);
  cache.emplace(key, L);
  return L;
}

int run(int kind, const float* E, const float* pids, int64_t N, int64_t D, int soft, float margin, int weighted,
        float* loss, float* num_active, float* diff, float* w, float* fp, float* cn, int* pos_idx, int* neg_idx,
        float* dE, void* ws, size_t ws_bytes, cudaStream_t stream) {
  MMSIM_REQUIRE(E && pids && loss && num_active && diff && w && fp && cn && pos_idx && neg_idx && ws, MMSIM_ERR_ARG,
                "loss: null pointer argument");
  MMSIM_REQUIRE(kind == 0 || kind == 1, MMSIM_ERR_ARG, "loss: kind must be 0 (batch_hard) or 1 (lifted)");
  MMSIM_REQUIRE(N >= 1 && N <= 1024, MMSIM_ERR_UNSUPPORTED, "loss: batch size N=%lld unsupported (1..1024)", (long long)N);
  MMSIM_REQUIRE(D >= 1 && D <= 512, MMSIM_ERR_UNSUPPORTED, "loss: embedding width D=%lld unsupported (1..512)", (long long)D);
  MMSIM_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MMSIM_ERR_WORKSPACE, "loss: workspace must be 256-byte aligned");
  const Layout L = make_layout(N, D);
  MMSIM_REQUIRE(L.shape >= 0, MMSIM_ERR_UNSUPPORTED, "loss: N=%lld, D=%lld does not fit a co-resident grid on this device",
                (long long)N, (long long)D);
  MMSIM_REQUIRE(ws_bytes >= L.total_bytes, MMSIM_ERR_WORKSPACE, "loss: workspace too small (%zu < %zu)", ws_bytes, L.total_bytes);

  uint8_t* b = static_cast<uint8_t*>(ws);
  Params p{};
  p.E = E; p.pids = pids; p.N = int(N); p.D = int(D); p.kind = kind; p.soft = soft; p.weighted = weighted; p.margin = margin;
  p.loss = loss; p.num_active = num_active; p.diff = diff; p.w = w; p.fp = fp; p.cn = cn;
  p.pos_idx = pos_idx; p.neg_idx = neg_idx; p.dE = dE;
  p.sync = reinterpret_cast<unsigned int*>(b + L.off_sync);
  p.same_cnt = reinterpret_cast<int*>(b + L.off_same);
  p.p_a = reinterpret_cast<float*>(b + L.off_tab[0]); p.p_b = reinterpret_cast<float*>(b + L.off_tab[1]);
  p.p_c = reinterpret_cast<float*>(b + L.off_tab[2]); p.p_d = reinterpret_cast<float*>(b + L.off_tab[3]);
  p.p_ia = reinterpret_cast<int*>(b + L.off_tab[4]); p.p_ib = reinterpret_cast<int*>(b + L.off_tab[5]);
  p.p_ca = reinterpret_cast<int*>(b + L.off_tab[6]); p.p_cb = reinterpret_cast<int*>(b + L.off_tab[7]);
  p.row_loss = reinterpret_cast<float*>(b + L.off_row_loss);
  p.row_active = reinterpret_cast<float*>(b + L.off_row_active);
  p.trace = reinterpret_cast<unsigned long long*>(b + L.off_trace);
  p.NBI = L.NBI; p.NBJ = L.NBJ; p.Npad = L.Npad;

  if (kind == 0 && bh_rows_ok(N, D)) {
    static bool attr_set[kMaxDev + 1] = {};
    const int slot = current_dev_slot();
    if (!attr_set[slot]) {
      MMSIM_CUDA_CHECK(cudaFuncSetAttribute(bh_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[slot] = true;
    }
    if (dE) MMSIM_CUDA_CHECK(cudaMemsetAsync(dE, 0, size_t(N) * D * sizeof(float), stream));
    bh_rows_kernel<<<unsigned((N + BR_R - 1) / BR_R), BR_THREADS, bh_rows_smem(N, D), stream>>>(p);
    MMSIM_CUDA_CHECK(::mmsim::launched());
    return MMSIM_OK;
  }

  void* args[] = {const_cast<Params*>(&p)};
  const dim3 grid(unsigned(L.NBI * L.NBJ)), block(THREADS);
  const void* fn = kernel_for(L.shape, kind);
  MMSIM_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, grid, block, args, L.smem_bytes, stream));
  MMSIM_CUDA_CHECK(::mmsim::launched());
  return MMSIM_OK;
}

}  // namespace loss
}  // namespace mmsim
