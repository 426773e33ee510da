// kernels_knightthrough.cu -- the Knightthrough instantiations of rollout_kernel (one translation unit per game: parallel build).
#include "kernel_table.cuh"

namespace mr {

kernel_fn pick_knightthrough(int policy, int v) { return pick_all_policies<Knightthrough>(policy, v); }
cudaError_t upload_tables_knightthrough(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp) {
    return upload_tables_here(stride, inv, rcp);
}

} // namespace mr
