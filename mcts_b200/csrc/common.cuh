// common.cuh -- shared device-side definitions of the rollout engine (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mcts_rollout.h"

namespace mr {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int WARPS_PER_BLOCK = 4;
constexpr int THREADS_PER_BLOCK = WARPS_PER_BLOCK * 32;
// kernel-side name of DecisiveMove<EpsilonGreedy<Mast>>: the ABI's Win and WinLoss modes share one instantiation
constexpr int MR_DECISIVE_MAST = MR_DECISIVE_MAST_WIN;
__host__ __device__ constexpr bool policy_uses_mast(int policy) {
    return policy == MR_EPS_MAST || policy == MR_DECISIVE_MAST_WIN || policy == MR_DECISIVE_MAST_WINLOSS || policy == MR_EPS_NST ||
           policy == MR_EPS_LGR2_MAST;
}

__host__ __device__ constexpr bool policy_is_lgr2(int policy) { return policy == MR_EPS_LGR2 || policy == MR_EPS_LGR2_MAST; }
__host__ __device__ constexpr bool policy_is_lgr(int policy) { return policy == MR_EPS_LGR || policy_is_lgr2(policy); }
// the policies whose last resort is the (epsilon-)greedy MAST pick
__host__ __device__ constexpr bool policy_mast_inner(int policy) {
    return policy == MR_EPS_MAST || policy == MR_DECISIVE_MAST_WIN || policy == MR_EPS_NST || policy == MR_EPS_LGR2_MAST;
}

constexpr int MAST_KINDS = 8;  // move kinds per source cell: 3 (Breakthrough) or 8 (Knightthrough)
constexpr int MAST_BITS = 10;  // <= 513 distinct score values over the <= 512 real actions of a player

// step() result codes.  Bit 3 says a ply was made; the low bits say how the playout ended.
enum : int {
    ST_CONTINUE = 0,
    ST_WIN_P0 = 1,
    ST_WIN_P1 = 2,
    ST_DRAW = 3,       // terminal draw
    ST_NO_ACTIONS = 4, // no actions but not terminal: NaturalEnd/NotTerminal, scored as a draw (simulate.rs:112-116)
    ST_END_MASK = 7,
    ST_MOVED = 8
};

// Everything one launch needs; passed by value (lives in the constant bank, so warp-uniform reads of
// it are free operands of ALU instructions).
struct KernelParams {
    const mr_leaf *leaves;
    mr_leaf_result *results;
    unsigned long long *work_counter; // persistent-warp work queue head: next global playout index (relative to g_begin)
    uint32_t n_leaves;
    uint32_t k;               // playouts per leaf
    // Work = global playout indices g in [g_begin, g_end), g = leaf * k + playout.  Warps grab contiguous
    // ranges with GUIDED self-scheduling (grab size ~ remaining / (2 * warps), between grab_min and grab_max):
    // large grabs early keep the per-grab tail short, small grabs at the end balance the warps.
    unsigned long long g_begin;
    unsigned long long g_end;
    uint32_t total_warps;
    uint32_t grab_min, grab_max; // multiples of 32
    uint32_t grab_div;           // grab = remaining / (grab_div * warps)
    uint32_t max_depth; // 0xffffffff = unlimited
    uint32_t seed_lo, seed_hi;
    // the ten Philox round keys (seed_lo + r * W0, seed_hi + r * W1), computed on the host: the rounds read them as
    // constant-bank operands of their LOP3s instead of bumping the key with an add per round
    uint32_t philox_keys[10][2];
    uint32_t leaf_id_base;
    uint32_t eps_thr_m1; // explore iff eps_enable && x1 <= eps_thr_m1
    uint32_t eps_enable;
    uint32_t max_trace;
    // optional outputs
    mr_action_stat *amaf;       // [n_leaves][2][NA]
    mr_action_stat *mast_delta; // [2][NA]
    uint16_t *trace;            // [n_leaves*k][max_trace]
    mr_playout_record *rec;     // [n_leaves*k]
    mr_leaf *final_state;       // [n_leaves*k]
    // MAST snapshot as BIT-PLANES of dense order-preserving ranks of score/visits (unvisited == 1.0):
    // bit s of mast_planes[player][kind][b] = bit b of rank(player, action (src = s, kind)).  The greedy
    // pick is then a bit-sliced arg-max over the legal-move masks, with no table lookups at all.
    uint64_t mast_planes[2][MAST_KINDS][MAST_BITS];
    uint32_t mast_nbits[2];
    uint16_t mast_rank_small[2][8]; // Connect4: the same dense ranks, one per column, read directly
    uint16_t *trace_scratch;   // [resident warps][trace_rows][32] u16, MAST-delta / AMAF staging
    uint32_t trace_cap;
    uint32_t trace_rows;       // rows of a warp's scratch: trace_cap (LGR / NST: one row per ply), else the per-lane log's ring size (a power of two)
    // Last Good Reply (MR_EPS_LGR): frozen reply table in, last-winning-occurrence stamps out
    const uint16_t *lgr_in;          // [2][NA]: 1 + compact code of the reply to the preceding action id, 0 = none
    unsigned long long *lgr_stamps;  // [2][NA]: ((g + 1) << 26) | ((4096 + ply) << 13) | reply action id, max-merged
    uint32_t *lgr_last;              // [n_leaves][2]: 1 + last playout index the player did not lose, max-merged
    // LGRF-2 (MR_EPS_LGR2), on top of the LGR-1 fields: the frozen 2-ply table in; pass 1 writes the last insert per
    // context and each leaf's last lost playout; pass 2 (a replay with lgr2_target set) marks the contexts whose
    // candidate entry a later losing playout removes
    const uint16_t *lgr2_in;                // [2][S][S]: 1 + compact code of the reply to (own previous slot, opponent's slot)
    unsigned long long *lgr2_stamps;        // [2][S][S]: ((g + 1) << 26) | ((4096 + ply) << 13) | reply slot, max-merged
    uint32_t *lgr2_last_lose;               // [n_leaves][2]: 1 + last playout index the player lost, max-merged
    const unsigned long long *lgr2_target;  // pass 2, [2][S][S]: ((g* + 1) << 16) | (1 + candidate slot), 0 = no candidate
    uint8_t *lgr2_removed;                  // pass 2, [2][S][S]
    // N-gram Selection Technique (MR_EPS_NST): rank planes per CONTEXT (mover, 1 + slot of the preceding action; 0 = no
    // preceding action = the unigram ranks), resolved on the host from the bigram / unigram tables and the backoff
    const uint64_t *nst_planes;      // [2][NSLOTS + 1][KINDS][MAST_BITS]
    const uint32_t *nst_nbits;       // [2][NSLOTS + 1]
    const uint16_t *nst_rank_small;  // Connect4: [2][NSLOTS + 1][8] ranks per column
    unsigned long long *nst_delta;   // [2][NSLOTS][NSLOTS]: visits in the low word, score (two's complement) in the high word
    // MR_TABLES_IN_SLOTS: amaf / mast_delta are indexed [2][NSLOTS] by compact per-player action slot instead of
    // [2][NA] by dense action id (192 / 512 instead of 4096 entries per player for Breakthrough / Knightthrough)
    uint32_t tables_in_slots;
    uint32_t amaf_stride;      // entries per leaf in `amaf`: 2 * NA, or 2 * NSLOTS with tables_in_slots
    // Fused epilogue of the low-latency submit path (batch_epilogue below), active when done_counter != nullptr: the
    // last block to finish copies the per-leaf results and the MAST delta into HOST-mapped memory, re-zeroes the device
    // copies and the two counters for the next batch, and publishes done_seq in *host_flag -- one kernel launch per
    // batch, no memset / copy / synchronize calls around it; the host spins on the flag.
    unsigned int *done_counter;
    mr_leaf_result *host_results;     // [n_leaves], host-mapped
    mr_action_stat *host_mast_delta;  // [mast_delta_len], host-mapped (nullptr without a MAST delta)
    uint32_t mast_delta_len;
    uint32_t done_seq;
    volatile uint32_t *host_flag;
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint64_t mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint64_t bit64(uint32_t pos) { return 1ull << pos; }

// ---- Philox4x32-10 (Salmon et al. SC'11), the draw discipline's generator ----
// ctr = (leaf_id, playout_id, ply >> 2, stream), key = (seed_lo, seed_hi); word [ply & 3] serves ply.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
        const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}
// the same block with the round keys precomputed (KernelParams::philox_keys)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&keys)[10][2],
                                              uint32_t out[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
        const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
        c0 = h1 ^ c1 ^ keys[r][0];
        c1 = l1;
        c2 = h0 ^ c3 ^ keys[r][1];
        c3 = l0;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// ---- fused batch epilogue (low-latency submit path) ----
// Every block calls this after its warps have run out of work.  The block that takes the last ticket sees every other
// block's result atomics (each block fences before taking its ticket), copies them out and leaves the device buffers
// zeroed for the slot's next batch.
__device__ __forceinline__ void batch_epilogue(const KernelParams &p) {
    if (!p.done_counter) return;
    __shared__ uint32_t s_is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_is_last = atomicAdd(p.done_counter, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    {
        uint2 *src = reinterpret_cast<uint2 *>(p.results);
        uint2 *dst = reinterpret_cast<uint2 *>(p.host_results);
        const uint32_t cnt = p.n_leaves * (uint32_t)(sizeof(mr_leaf_result) / sizeof(uint2));
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            dst[i] = __ldcg(src + i);
            src[i] = make_uint2(0u, 0u);
        }
    }
    if (p.host_mast_delta && p.mast_delta) {
        uint2 *src = reinterpret_cast<uint2 *>(p.mast_delta);
        uint2 *dst = reinterpret_cast<uint2 *>(p.host_mast_delta);
        for (uint32_t i = threadIdx.x; i < p.mast_delta_len; i += blockDim.x) {
            dst[i] = __ldcg(src + i);
            src[i] = make_uint2(0u, 0u);
        }
    }
    if (threadIdx.x == 0) {
        *p.work_counter = 0ull;
        *p.done_counter = 0u;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *p.host_flag = p.done_seq;
        __threadfence_system();
    }
}

// random_best helpers, filled by mr_init (capi.cu): for n legal actions and prime index k,
//   MR_STRIDE[n][k]     = PRIMES[k] mod n                       (the visiting stride, util.rs:147-161)
//   MR_STRIDE_INV[n][k] = its inverse mod n (PRIMES[k] is a prime > n, hence coprime with n)
//   MR_RCP32[n]         = floor(2^32 / n) + 1, so that floor(t / n) == umulhi(t, MR_RCP32[n]) for t * n < 2^32
__device__ uint8_t MR_STRIDE[129][16];
__device__ uint8_t MR_STRIDE_INV[129][16];
__device__ uint32_t MR_RCP32[129];

// util.rs:129-132 PRIMES, packed for random_best's stride
__device__ __constant__ uint32_t MR_PRIMES[16] = {14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991,
                                                  53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693};

// ---- ALU -> FMA pipe balancing ----
// These kernels are bound by the ALU pipe (LOP3 / SHF / SEL / ISETP at 77-92 % busy, ncu) while the FMA pipe idles at
// ~15 %.  A conditional move or a conditional add written as a PREDICATED IMAD runs on the FMA pipe instead of taking an
// ALU slot as a SEL.  The multiplier 1 / -1 comes from constant memory, so the compiler cannot fold the IMAD back into a
// MOV / SEL / IADD3; a constant-bank operand costs no register and no extra instruction.
#ifndef MR_FMA_BALANCE
#define MR_FMA_BALANCE 1
#endif
__device__ __constant__ uint32_t MR_ONE = 1u;
__device__ __constant__ uint32_t MR_NEG_ONE = 0xffffffffu;

// dst = src when `nonzero` != 0
__device__ __forceinline__ void cmov_nz(uint32_t &dst, uint32_t src, uint32_t nonzero) {
#if MR_FMA_BALANCE
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p mad.lo.u32 %0, %1, %3, 0;\n\t}" : "+r"(dst) : "r"(src), "r"(nonzero), "r"(MR_ONE));
#else
    dst = nonzero ? src : dst;
#endif
}
// if (idx >= cnt) { idx -= cnt; pos += step; }   -- one step of a rank/select binary search
__device__ __forceinline__ void select_step(uint32_t &idx, uint32_t &pos, uint32_t cnt, uint32_t step) {
#if MR_FMA_BALANCE
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %0, %2;\n\t@p mad.lo.u32 %0, %2, %5, %0;\n\t@p mad.lo.u32 %1, %3, %4, %1;\n\t}"
        : "+r"(idx), "+r"(pos)
        : "r"(cnt), "r"(step), "r"(MR_ONE), "r"(MR_NEG_ONE));
#else
    if (idx >= cnt) {
        idx -= cnt;
        pos += step;
    }
#endif
}
// v << s with PTX semantics: a shift amount >= 32 (including "negative" amounts as unsigned) gives 0, so word q of a
// multi-word board receives bit `cell` as shl_clamp(1, cell - 32 q) with no compare / select
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

} // namespace mr
