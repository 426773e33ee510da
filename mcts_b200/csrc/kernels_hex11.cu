// kernels_hex11.cu -- the Hex 11x11 kernels: fill-then-search (uniform) and the per-ply state machine (the rest).
#include "hex_kernel.cuh"
#include "kernel_table.cuh"

namespace mr {

template <int POLICY>
static kernel_fn pick_fill_then_search(int v) { // AMAF and MAST delta from board masks
    switch (v) {
    case 0: return hex_rollout_kernel<POLICY, kFlagVariant[0]>;
    case 1: return hex_rollout_kernel<POLICY, kFlagVariant[1]>;
    case 2: return hex_rollout_kernel<POLICY, kFlagVariant[2]>;
    case 3: return hex_rollout_kernel<POLICY, kFlagVariant[3]>;
    default: return hex_rollout_kernel<POLICY, kFlagVariant[4]>;
    }
}

kernel_fn pick_hex11(int policy, int v, uint32_t flags) {
    // the uniform and epsilon-greedy MAST policies never look at who is connected: fill-then-search (hex_kernel.cuh);
    // every other policy needs the position's connectivity or a trace: the per-ply kernel (hex_step.cuh)
    (void)flags;
    if (policy == MR_UNIFORM) return pick_fill_then_search<MR_UNIFORM>(v);
    if (policy == MR_EPS_MAST) return pick_fill_then_search<MR_EPS_MAST>(v);
    return pick_all_policies<HexStep>(policy, v);
}
cudaError_t upload_tables_hex11(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp) {
    return upload_tables_here(stride, inv, rcp);
}

} // namespace mr
