// kernels_breakthrough.cu -- the Breakthrough instantiations of rollout_kernel (one translation unit per game: parallel build).
#include "kernel_table.cuh"

namespace mr {

kernel_fn pick_breakthrough(int policy, int v) { return pick_all_policies<Breakthrough>(policy, v); }
cudaError_t upload_tables_breakthrough(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp) {
    return upload_tables_here(stride, inv, rcp);
}

} // namespace mr
