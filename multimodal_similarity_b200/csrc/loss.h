// Internal declarations for the fused loss kernels (loss.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace mmsim {
namespace loss {

struct Layout {
  int shape, TI, TJ, NBI, NBJ, Npad;
  size_t off_sync, off_same, zero_bytes, off_tab[8], off_row_loss, off_row_active, off_trace, total_bytes, smem_bytes;
};
// Workspace layout for a batch of N rows of width D (grid shape depends on the device: pass its co-resident CTA capacity).
Layout make_layout(int64_t N, int64_t D);

int run(int kind, const float* E, const float* pids, int64_t N, int64_t D, int soft, float margin, int weighted,
        float* loss, float* num_active, float* diff, float* w, float* fp, float* cn, int* pos_idx, int* neg_idx,
        float* dE, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace loss
}  // namespace mmsim
