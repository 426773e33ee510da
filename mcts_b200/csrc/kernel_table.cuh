// kernel_table.cuh -- shared by the kernels_<game>.cu units: variant dispatch and the table upload of this unit.
#pragma once

#include "kernels.h"
#include "rollout_kernel.cuh"

namespace mr {

template <class G, int POLICY>
kernel_fn pick_variant(int v) {
    switch (v) {
    case 0: return rollout_kernel<G, POLICY, kFlagVariant[0]>;
    case 1: return rollout_kernel<G, POLICY, kFlagVariant[1]>;
    case 2: return rollout_kernel<G, POLICY, kFlagVariant[2]>;
    case 3: return rollout_kernel<G, POLICY, kFlagVariant[3]>;
    default: return rollout_kernel<G, POLICY, kFlagVariant[4]>;
    }
}

// copies the host-computed random_best tables into THIS translation unit's device globals (current device)
inline cudaError_t upload_tables_here(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp) {
    cudaError_t e = cudaMemcpyToSymbol(MR_STRIDE, stride, sizeof(uint8_t) * 129 * 16);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(MR_STRIDE_INV, inv, sizeof(uint8_t) * 129 * 16);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(MR_RCP32, rcp, sizeof(uint32_t) * 129);
}

// the policies every per-ply game is built with
template <class G>
kernel_fn pick_all_policies(int policy, int v) {
    if (policy == MR_UNIFORM) return pick_variant<G, MR_UNIFORM>(v);
    if (policy == MR_EPS_MAST) return pick_variant<G, MR_EPS_MAST>(v);
    if (policy == MR_DECISIVE_WIN) return pick_variant<G, MR_DECISIVE_WIN>(v);
    if (policy == MR_DECISIVE_ANTI) return pick_variant<G, MR_DECISIVE_ANTI>(v);
    if (policy == MR_EPS_LGR) return pick_variant<G, MR_EPS_LGR>(v);
    if (policy == MR_DECISIVE_MAST_WIN) return pick_variant<G, MR_DECISIVE_MAST>(v);
    if (policy == MR_EPS_NST) return pick_variant<G, MR_EPS_NST>(v);
    if (policy == MR_EPS_LGR2) return pick_variant<G, MR_EPS_LGR2>(v);
    if (policy == MR_EPS_LGR2_MAST) return pick_variant<G, MR_EPS_LGR2_MAST>(v);
    return nullptr;
}

} // namespace mr
