// kernels.h -- what capi.cu sees of the per-game kernel translation units (kernels_<game>.cu).
//
// Every game's kernel instantiations live in their own .cu so that the library builds in parallel.  Device
// globals (the random_best tables of common.cuh) are per translation unit without relocatable device code, so
// every unit also exports an upload function for its own copy.
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace mr {

typedef void (*kernel_fn)(const KernelParams);

// flag variants: 0 = none, 1 = AMAF, 2 = MAST delta, 3 = AMAF|MAST, 4 = everything (debug / parity outputs)
constexpr int kFlagVariant[5] = {0, MR_WANT_AMAF, MR_WANT_MAST_DELTA, MR_WANT_AMAF | MR_WANT_MAST_DELTA,
                                 MR_WANT_AMAF | MR_WANT_MAST_DELTA | MR_WANT_TRACE | MR_WANT_FINAL_STATE};

// (policy, flag variant) -> kernel, nullptr when the policy is not built; `policy` is already canonical
// (WinLoss / WinLossDraw -> Win, DecisiveMast WinLoss -> Win)
kernel_fn pick_connect4(int policy, int v);
kernel_fn pick_breakthrough(int policy, int v);
kernel_fn pick_knightthrough(int policy, int v);
kernel_fn pick_othello(int policy, int v);
kernel_fn pick_hex11(int policy, int v, uint32_t flags);

typedef cudaError_t (*table_upload_fn)(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp);
cudaError_t upload_tables_connect4(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp);
cudaError_t upload_tables_breakthrough(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp);
cudaError_t upload_tables_knightthrough(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp);
cudaError_t upload_tables_othello(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp);
cudaError_t upload_tables_hex11(const uint8_t (*stride)[16], const uint8_t (*inv)[16], const uint32_t *rcp);

} // namespace mr
