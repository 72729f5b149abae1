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
  // query-streaming sweep (knn_sweepq.cuh; the default): one candidate log per query (n_splits == 1), work items are
  // (gallery tile, query chunk); host-buffer mode copies / sweeps the gallery in host_splits tile ranges
  int sweepq, host_splits;
  int dual;                   // sweep with two query blocks resident per CTA (K <= 128): work items are PAIRS of query blocks
  size_t off_tau, off_state;
  int unc_cap;                // capacity of the uncertified-query lists (= nq: every query may take the exact fallback)
  int logcap, use_pivots;     // candidate log capacity per (query, split); pivot pre-pass used (gallery larger than a log)
  // pivot pre-pass: a systematic sample of the gallery rows (every sample_div-th row, the offset inside the stride changes
  // from segment to segment), gathered into a compact block of n_sample_tiles tiles -- independent of the row ORDER
  int n_sample, n_sample_tiles, sample_div, sample_seg, pivot_grid;
  int n_anchor, group_blocks, group_default; // query grouping: anchors (0: never), blocks of the counting sort, on outside shard mode
  size_t off_ah, off_apack, off_aidx, off_assign, off_perm, off_ghist;
  size_t off_qh, off_gh, off_gpack, off_qnorm, off_qerr, off_stats, off_piv16, off_ladder, off_log, off_log_cnt, off_log_tau, off_split_done;
  size_t off_unc_query, off_unc_bound, off_fb2_list, off_fb_qh, off_fb_ladder, off_fb2_dist, off_fb2_idx;   // exact fallback
  size_t off_sidx, off_s32, off_sh, off_spack;   // the gallery sample: row indices, fp32 rows (host-buffer mode), fp16, norm pack
  size_t total_bytes;
};

// host_mode: the plan of mmsim_knn_host_f32 (more gallery splits: the unit of its copy / sweep pipeline)
Plan make_plan(int64_t nq, int64_t ng, int64_t D, int k, int num_sms, bool host_mode = false);

enum : int { kPhasePrep = 1, kPhaseTensor = 2, kPhaseRerank = 4, kPhaseFallback = 8, kPhasePivot = 16, kPhaseLadder = 32, kPhaseAll = 63,
             kPhaseFinish = 64 /* one more tier-2 wave of the exact fallback (mmsim_knn_finish_f32); not part of kPhaseAll */,
             kPhasePrepQ = 128, kPhasePrepG = 256 /* the two halves of kPhasePrep on their own */ };

// gallery row of sample j (host and device agree: the host-buffer call copies these rows with strided 2-D copies)
__host__ __device__ inline int64_t sample_row(int64_t j, int div, int seg) {
  const uint32_t h = uint32_t(j / seg) * 2654435761u;
  return j * div + int64_t((h >> 7) % uint32_t(div));
}

// Host-buffer mode (mmsim_knn_host_f32): Q and G given to run() are device STAGING buffers; run() fills them from these
// host arrays on a copy stream, one gallery split at a time, while the sweep of the previous split is running.
struct HostPipe {
  const float* q_host;
  const float* g_host;
};

int run(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self, int64_t self_offset,
        float* out_dist, int* out_idx, int* status, void* ws, size_t ws_bytes, cudaStream_t stream, int phases = kPhaseAll,
        int shard_kp = 0, float* out_lb = nullptr, const HostPipe* host = nullptr, int64_t slice_rows = 0,
        int64_t slice_stride = 0);

constexpr int kPivotsPerRow = 16;  // floats per query row in the pivot region of the workspace
int merge_pivots(const float* parts, int nparts, int64_t part_stride, int64_t rows, float* out, cudaStream_t stream);

// Gallery-shard mode: exact top-k INSIDE this shard for the queries flagged uncertified by the global certificate
// (flag[q] >= 0: the merged k-th distance, an upper bound of the true one; < 0: certified).  Uses the gallery operand
// copies a previous run() left in the workspace.  out_dist / out_idx: compact [cap, k] blocks, slot order = ascending
// query index (identical on every rank); out_query[cap] the query of each slot; status[0] = flagged queries (may exceed
// cap: nothing beyond cap is computed), status[1] - status[2] = tier-2 queries still pending (see knn_fallback.cuh).
int shard_fallback(const float* Q, int64_t nq, const float* G, int64_t ng, int64_t D, int k, int exclude_self, int64_t self_offset,
                   const float* flag, int cap, float* out_dist, int* out_idx, int* out_query, int* status, void* ws,
                   size_t ws_bytes, cudaStream_t stream, bool host_layout = false /* the workspace was filled by host-buffer calls */);

}  // namespace knn
}  // namespace mmsim
