// Internal declarations shared by the kNN translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace mmsim {
namespace knn {

constexpr int KP = 128;  // approximate candidates re-ranked exactly per query; k <= KP - 16

// Launch geometry + workspace layout of one mmsim_knn_f32 call (pure function of the problem size).
struct Plan {
  int Dp, katoms;             // padded feature width (multiple of 64) and number of 64-wide K atoms
  int n_qblocks, n_tiles;     // 128-query blocks, 256-row gallery tiles
  int n_splits, tiles_per_split, grid;
  int unc_cap;                // max uncertified queries handled by the exact fallback
  int logcap, use_pivots, n_sample_tiles, sample_cols, pivot_grid;   // candidate log / pivot pre-pass geometry
  int n_anchor, group_blocks, group_default; // query grouping: anchors (0: never), blocks of the counting sort, on outside shard mode
  size_t off_ah, off_apack, off_aidx, off_assign, off_perm, off_ghist;
  size_t off_qh, off_gh, off_gpack, off_qnorm, off_qerr, off_stats, off_piv16, off_ladder, off_log, off_log_cnt, off_log_tau, off_split_done;
  size_t off_unc_query, off_unc_bound, off_fb_count, off_fb_dist, off_fb_idx;
  size_t off_s32, off_sh, off_spack;   // host-buffer mode: compact copy of the sampled gallery tiles (fp32, fp16, norm pack)
  size_t total_bytes;
};

// host_mode: the plan of mmsim_knn_host_f32 (more gallery splits: the unit of its copy / sweep pipeline)
Plan make_plan(int64_t nq, int64_t ng, int64_t D, int k, int num_sms, bool host_mode = false);

enum : int { kPhasePrep = 1, kPhaseTensor = 2, kPhaseRerank = 4, kPhaseFallback = 8, kPhasePivot = 16, kPhaseLadder = 32, kPhaseAll = 63 };

// Host-buffer mode (mmsim_knn_host_f32): Q and G given to run() are device STAGING buffers; run() fills them from these
// host arrays on a copy stream, one gallery split at a time, while the sweep of the previous split is running.
struct HostPipe {
  const float* q_host;
  const float* g_host;
};

int run(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self, int64_t self_offset,
        float* out_dist, int* out_idx, int* status, void* ws, size_t ws_bytes, cudaStream_t stream, int phases = kPhaseAll,
        int shard_kp = 0, float* out_lb = nullptr, const HostPipe* host = nullptr);

constexpr int kPivotsPerRow = 16;  // floats per query row in the pivot region of the workspace
int merge_pivots(const float* parts, int nparts, int64_t part_stride, int64_t rows, float* out, cudaStream_t stream);

}  // namespace knn
}  // namespace mmsim
