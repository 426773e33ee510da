// kernels_othello.cu -- the Othello instantiations of rollout_kernel (one translation unit per game: parallel build).
#include "kernel_table.cuh"

namespace mr {

kernel_fn pick_othello(int policy, int v) {
    // The decisive scan provably never changes an Othello pick (DESIGN.md "Decisive move"), so DecisiveMove<Win> shares
    // the uniform kernel and DecisiveMove<EpsilonGreedy<Mast>> the MAST kernel.
    if (policy == MR_UNIFORM || policy == MR_DECISIVE_WIN) return pick_variant<Othello, MR_UNIFORM>(v);
    if (policy == MR_EPS_MAST || policy == MR_DECISIVE_MAST_WIN) return pick_variant<Othello, MR_EPS_MAST>(v);
    if (policy == MR_DECISIVE_ANTI) return pick_variant<Othello, MR_DECISIVE_ANTI>(v);
    if (policy == MR_EPS_LGR) return pick_variant<Othello, MR_EPS_LGR>(v);
    if (policy == MR_EPS_NST) return pick_variant<Othello, MR_EPS_NST>(v);
    if (policy == MR_EPS_LGR2) return pick_variant<Othello, MR_EPS_LGR2>(v);
    if (policy == MR_EPS_LGR2_MAST) return pick_variant<Othello, MR_EPS_LGR2_MAST>(v);
    return nullptr;
}
cudaError_t upload_tables_othello(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp) {
    return upload_tables_here(stride, inv, rcp);
}

} // namespace mr
