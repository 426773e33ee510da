// kernels_hex11.cu -- the Hex 11x11 kernels: fill-then-search (uniform) and the per-ply state machine (the rest).
#include "hex_kernel.cuh"
#include "hex_step.cuh"
#include "kernel_table.cuh"

namespace mr {

static kernel_fn pick_fill_then_search(int v) { // AMAF from board masks; no MAST delta
    switch (v) {
    case 0: return hex_rollout_kernel<0>;
    case 1: return hex_rollout_kernel<MR_WANT_AMAF>;
    default: return hex_rollout_kernel<MR_WANT_AMAF | MR_WANT_TRACE | MR_WANT_FINAL_STATE>;
    }
}

kernel_fn pick_hex11(int policy, int v, uint32_t flags) {
    // uniform playouts never look at the position: fill-then-search (hex_kernel.cuh), unless the caller wants the
    // MAST delta, which only the per-ply kernel (hex_step.cuh) produces; every other policy needs the position
    if (policy == MR_UNIFORM && !(flags & MR_WANT_MAST_DELTA)) return pick_fill_then_search(v);
    return pick_all_policies<HexStep>(policy, v);
}
cudaError_t upload_tables_hex11(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp) {
    return upload_tables_here(stride, inv, rcp);
}

} // namespace mr
