// hex_kernel.cuh -- Hex 11x11 playouts as FILL-THEN-SEARCH (sm_100a).
//
// The reference checks the last mover's connection with a full flood6 fix-point after every ply
// (games/hex-gen/src/lib.rs:130-144,208-215 through the default terminal_status, mcts/src/game.rs:202-211).
// That answer is MONOTONE along a playout -- a connected chain stays connected, and in Hex at most one
// player can ever connect -- and the moves themselves never depend on it (the idx-th empty cell, ascending,
// lib.rs:115-120).  So a playout is computed as:
//   1. fill  : play the uniform policy to the ply limit D = min(#empty cells, max_playout_depth) with NO
//              connectivity work, recording the cell of every ply.  All playouts of a leaf have the same
//              length, so the 32 lanes of a warp run in perfect lock-step (no refill, no dead lanes).
//   2. who   : one flood per player on the position at ply D decides the winner (none => depth cutoff).
//   3. when  : binary search over the winner's moves for the FIRST one that connects its two edges; each
//              probe grows the stone mask and the edge-connected set incrementally from the last
//              "not yet connected" probe.
// The stop ply, winner, trace, final state and AMAF occurrences are exactly those of the per-ply loop.
// Moves generated past the stop ply are discarded; they consume Philox words the reference would never
// have drawn, which is unobservable because draws are keyed by (leaf, playout, ply).
//
// The same holds for EpsilonGreedy<Mast> (simulate.rs:538-570, 824-848): its pick looks at the position only through
// the set of EMPTY cells (the legal actions) and at the frozen MAST table -- never at who is connected -- so the ply
// sequence of an epsilon-MAST playout is independent of where the game ends, and fill-then-search applies unchanged
// with the bit-sliced arg-max + random_best tie-break as the fill step (POLICY == MR_EPS_MAST; two Philox words per
// ply, as in the per-ply kernels).  Every Hex cell is played at most once per playout and by a known player, so the
// MAST delta (like the AMAF table) is "stones at the stop ply minus leaf stones", added without a trace.
// The decisive / reply / n-gram policies do read connectivity or need traces: they run in hex_step.cuh.
#pragma once

#include "hex_step.cuh"
#include "rollout_kernel.cuh"

namespace mr {

struct HexLeafShared {
    Hex11::B128 stones[2];
    Hex11::B128 conn[2]; // flood6(first edge) of each player at the leaf
    uint32_t turn;
    uint32_t n_empty;
};

#ifndef MR_HEX_MAST_BLOCKS
#define MR_HEX_MAST_BLOCKS 6
#endif
template <int POLICY, int FLAGS>
__global__ void __launch_bounds__(THREADS_PER_BLOCK, POLICY == MR_EPS_MAST ? MR_HEX_MAST_BLOCKS : 8 /* 64 registers: +3.7% measured */)
    hex_rollout_kernel(const KernelParams p) {
    using H = Hex11;
    using B128 = H::B128;
    static_assert(POLICY == MR_UNIFORM || POLICY == MR_EPS_MAST, "fill-then-search covers the policies that never read connectivity");
    constexpr bool IS_MAST = POLICY == MR_EPS_MAST;
    constexpr uint32_t WINDOW = IS_MAST ? 2u : 4u; // plies per Philox block
    constexpr bool WANT_AMAF = (FLAGS & MR_WANT_AMAF) != 0;
    constexpr bool WANT_MAST = (FLAGS & MR_WANT_MAST_DELTA) != 0;
    constexpr bool WANT_TRACE = (FLAGS & MR_WANT_TRACE) != 0;
    constexpr bool WANT_FINAL = (FLAGS & MR_WANT_FINAL_STATE) != 0;
    constexpr bool NEED_BOARDS = WANT_AMAF || WANT_MAST || WANT_FINAL; // both players' stones at the stop ply
    constexpr B128 NC0 = H::inv(H::col_mask(0)), NC10 = H::inv(H::col_mask(10)), BM = H::board_mask();

    __shared__ __align__(16) mr_leaf s_leaf[WARPS_PER_BLOCK];
    __shared__ HexLeafShared s_hex[WARPS_PER_BLOCK];
    __shared__ uint8_t s_order[WARPS_PER_BLOCK][121][32]; // cell played at [ply][lane]
    __shared__ uint32_t s_amaf[WARPS_PER_BLOCK][WANT_AMAF ? 2 * H::NSLOTS : 1];
    __shared__ uint32_t s_mast[WARPS_PER_BLOCK][WANT_MAST ? 2 * H::NSLOTS : 1];
    // s_nth[v][r] = position of the r-th set bit of the byte v: the last step of the fill phase's rank -> cell select
    __shared__ uint8_t s_nth[256][8];
    for (uint32_t v = threadIdx.x; v < 256u; v += THREADS_PER_BLOCK) {
        uint32_t r = 0;
#pragma unroll
        for (uint32_t b = 0; b < 8u; ++b) {
            s_nth[v][b] = 0;
            if ((v >> b) & 1u) s_nth[v][r++] = (uint8_t)b;
        }
    }
    __syncthreads();

    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    HexLeafShared &sh = s_hex[warp];
    uint8_t(*order)[32] = s_order[warp];
    uint32_t *amaf_tab = s_amaf[warp];
    uint32_t *mast_tab = s_mast[warp];
    uint32_t amaf_plies = 0, mast_plies = 0;
    if (WANT_AMAF)
        for (uint32_t e = lane; e < 2u * H::NSLOTS; e += 32) amaf_tab[e] = 0;
    if (WANT_MAST)
        for (uint32_t e = lane; e < 2u * H::NSLOTS; e += 32) mast_tab[e] = 0;
    __syncwarp();

    unsigned long long grab_lo = 0, grab_hi = 0;
    const unsigned long long work_total = p.g_end - p.g_begin;
    for (;;) {
        if (grab_lo >= grab_hi) {
            unsigned long long lo = 0;
            uint32_t want = 0;
            if (lane == 0) { // guided self-scheduling, as in rollout_kernel
                const unsigned long long seen = *reinterpret_cast<volatile unsigned long long *>(p.work_counter);
                const unsigned long long remaining = seen < work_total ? work_total - seen : 0ull;
                unsigned long long w = remaining / ((unsigned long long)p.grab_div * p.total_warps);
                // half the grab sizes of the per-ply kernels: a Hex playout is ~90 plies of fill plus the search, so a grab
                // is long and the end-of-kernel tail matters more than the per-task set-up (13.88 -> 13.65 ms measured)
                const uint32_t gmax = p.grab_max > 64u ? p.grab_max >> 1 : p.grab_max, gmin = p.grab_min > 64u ? p.grab_min >> 1 : p.grab_min;
                w = w > gmax ? gmax : w;
                w = w < gmin ? gmin : w;
                want = (uint32_t)w & ~31u;
                lo = atomicAdd(p.work_counter, (unsigned long long)want);
            }
            lo = __shfl_sync(FULL_MASK, lo, 0);
            want = __shfl_sync(FULL_MASK, want, 0);
            if (lo >= work_total) break;
            grab_lo = lo;
            grab_hi = lo + want < work_total ? lo + want : work_total;
        }
        const unsigned long long g0 = p.g_begin + grab_lo;
        const uint32_t li = (uint32_t)(g0 / p.k);
        const uint32_t pid_begin = (uint32_t)(g0 - (unsigned long long)li * p.k);
        const unsigned long long room = (unsigned long long)p.k - pid_begin;
        const unsigned long long take = grab_hi - grab_lo < room ? grab_hi - grab_lo : room;
        const uint32_t pid_end = pid_begin + (uint32_t)take;
        grab_lo += take;
        const uint32_t leaf_id = p.leaf_id_base + li;

        // ---- stage the leaf, flood both players' first edges once per task ----
        __syncwarp();
        if (lane < sizeof(mr_leaf) / 4)
            reinterpret_cast<uint32_t *>(&s_leaf[warp])[lane] = reinterpret_cast<const uint32_t *>(p.leaves + li)[lane];
        __syncwarp();
        int leaf_status;
        {
            const mr_leaf &leaf = s_leaf[warp];
            B128 st[2], cn[2];
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
                st[pl].w[0] = (uint32_t)leaf.board[2 * pl];
                st[pl].w[1] = (uint32_t)(leaf.board[2 * pl] >> 32);
                st[pl].w[2] = (uint32_t)leaf.board[2 * pl + 1];
                st[pl].w[3] = (uint32_t)(leaf.board[2 * pl + 1] >> 32) & 0x01ffffffu;
                cn[pl] = H::flood(H::first_edge(pl), st[pl], NC0, NC10, BM);
            }
            const B128 occ = H::bor(st[0], st[1]);
            const uint32_t n_occ = __popc(occ.w[0]) + __popc(occ.w[1]) + __popc(occ.w[2]) + __popc(occ.w[3]);
            if (lane == 0) {
                sh.stones[0] = st[0];
                sh.stones[1] = st[1];
                sh.conn[0] = cn[0];
                sh.conn[1] = cn[1];
                sh.turn = leaf.turn;
                sh.n_empty = 121 - n_occ;
            }
            const int last = (leaf.turn + 1) & 1; // only the player who moved last is checked (lib.rs:130-144)
            if (H::any(H::band(cn[last], H::second_edge(last)))) leaf_status = last ? MR_WINNER_P1 : MR_WINNER_P0;
            else if (n_occ == 121) leaf_status = MR_DRAW;
            else leaf_status = MR_NOT_TERMINAL;
        }
        __syncwarp();
        const uint32_t leaf_turn = sh.turn;
        const uint32_t E = sh.n_empty;
        const uint32_t D = min(E, p.max_depth); // plies every playout of this leaf is filled to

        uint32_t acc_w0 = 0, acc_w1 = 0, acc_draw = 0, acc_cut = 0, acc_depth = 0;

        if (leaf_status != MR_NOT_TERMINAL) {
            const uint32_t cnt = pid_end - pid_begin;
            if (lane == 0) {
                if (leaf_status == MR_WINNER_P0) acc_w0 = cnt;
                else if (leaf_status == MR_WINNER_P1) acc_w1 = cnt;
                else acc_draw = cnt;
            }
            if (WANT_TRACE || WANT_FINAL) {
                for (uint32_t pid = pid_begin + lane; pid < pid_end; pid += 32) {
                    const size_t g = (size_t)li * p.k + pid;
                    if (WANT_TRACE && p.rec) {
                        mr_playout_record r;
                        r.depth = 0;
                        r.outcome = (uint8_t)leaf_status;
                        r.end_type = MR_END_TERMINAL;
                        r.pad[0] = r.pad[1] = 0;
                        p.rec[g] = r;
                    }
                    if (WANT_FINAL && p.final_state) {
                        mr_leaf fs = s_leaf[warp];
                        fs.prev_action_plus1 = 0; // call context, not game state
                        fs.prev2_action_plus1 = 0;
                        p.final_state[g] = fs;
                    }
                }
            }
        } else {
            for (uint32_t round = pid_begin; round < pid_end; round += 32) {
                const uint32_t pid = round + lane;
                const bool active = pid < pid_end;

                // ---- 1. fill: D plies, lock-step, no connectivity work ----
                B128 full[2] = {sh.stones[0], sh.stones[1]};
                uint32_t n_empty = E; // warp-uniform
                for (uint32_t base = 0; base < D; base += WINDOW) {
                    uint32_t r[4];
                    philox4x32_10(leaf_id, pid, IS_MAST ? base >> 1 : base >> 2, IS_MAST ? 1u : 0u, p.philox_keys, r);
#pragma unroll
                    for (uint32_t j = 0; j < WINDOW; ++j) {
                        const uint32_t ply = base + j;
                        if (ply < D) { // warp-uniform
                            const uint32_t x0 = IS_MAST ? r[2u * j] : r[j];
                            const uint32_t mover = (leaf_turn + ply) & 1u; // warp-uniform
                            uint32_t idx = __umulhi(x0, n_empty);
                            const uint32_t e0 = ~(full[0].w[0] | full[1].w[0]);
                            const uint32_t e1 = ~(full[0].w[1] | full[1].w[1]);
                            const uint32_t e2 = ~(full[0].w[2] | full[1].w[2]);
                            const uint32_t e3 = ~(full[0].w[3] | full[1].w[3]) & 0x01ffffffu;
                            uint32_t cell = 0;
                            bool have = false;
                            if constexpr (IS_MAST) {
                                // EpsilonGreedy (simulate.rs:538-570): explore iff x1 < floor(eps * 2^32), else Mast::select_move
                                // (:824-848): bit-sliced arg-max over the empty cells (the two 64-bit halves of the board ride the
                                // "kind" dimension of the rank planes), then random_best's visiting order (util.rs:135-163)
                                const uint32_t x1 = r[2u * j + 1u];
                                if (!(p.eps_enable && x1 <= p.eps_thr_m1)) {
                                    uint64_t cand[2] = {mk64(e0, e1), mk64(e2, e3)};
                                    mast_argmax<2>(cand, p.mast_planes[mover], p.mast_nbits[mover]);
                                    const B128 c = B128{{lo32(cand[0]), hi32(cand[0]), lo32(cand[1]), hi32(cand[1])}};
                                    const B128 empties = B128{{e0, e1, e2, e3}};
                                    const uint32_t m = HexStep::count(c);
                                    const uint32_t rr = __umulhi(x0, 16u * n_empty);
                                    const uint32_t i0 = rr >> 4;
                                    if (m == 1) {
                                        cell = HexStep::first_cell(c);
                                        have = true;
                                    } else if (m == n_empty) {
                                        idx = i0; // every empty cell is maximal: the first visited list position
                                    } else if (m <= 8) {
                                        const uint32_t invs = MR_STRIDE_INV[n_empty][rr & 15u], rcp = MR_RCP32[n_empty];
                                        uint32_t best_j = 0xffffffffu;
                                        B128 rest = c;
                                        for (uint32_t q = 0; q < m; ++q) {
                                            const uint32_t c0 = HexStep::first_cell(rest);
                                            rest = H::bandn(rest, HexStep::cell_bit(c0));
                                            const uint32_t jr = visiting_rank(HexStep::rank128(empties, c0), i0, n_empty, invs, rcp);
                                            if (jr < best_j) {
                                                best_j = jr;
                                                cell = c0;
                                            }
                                        }
                                        have = true;
                                    } else {
                                        const uint32_t stride = MR_STRIDE[n_empty][rr & 15u];
                                        uint32_t i = i0;
                                        for (uint32_t q = 0; q < n_empty; ++q) {
                                            const uint32_t c0 = HexStep::select128(empties, i);
                                            if (HexStep::has_cell(c, c0)) {
                                                cell = c0;
                                                break;
                                            }
                                            i += stride;
                                            if (i >= n_empty) i -= n_empty;
                                        }
                                        have = true;
                                    }
                                }
                            }
                            if (!have) {
                                const uint32_t c0 = __popc(e0), c1 = __popc(e1), c2 = __popc(e2);
                                // the idx-th empty cell (ascending), branch-free: three compare-and-select steps pick the
                                // word, SWAR prefix popcounts the byte, a table the bit.  (Measured against the 5-step binary
                                // search it replaces: 8 % fewer instructions, 14.05 -> 13.82 ms.  The same word pick with
                                // predicated IMADs: 14.02 ms; prefix POPCs of the word's low 1, 2, 3 bytes instead of the
                                // SWAR counts: 13.92 ms.)
                                uint32_t wbase = 0, word = e0;
                                {
                                    const uint32_t p1 = c0, p2 = c0 + c1, p3 = p2 + c2;
                                    uint32_t sub = 0;
                                    if (idx >= p1) {
                                        word = e1;
                                        wbase = 32;
                                        sub = p1;
                                    }
                                    if (idx >= p2) {
                                        word = e2;
                                        wbase = 64;
                                        sub = p2;
                                    }
                                    if (idx >= p3) {
                                        word = e3;
                                        wbase = 96;
                                        sub = p3;
                                    }
                                    idx -= sub;
                                }
                                uint32_t bc = word - ((word >> 1) & 0x55555555u);
                                bc = (bc & 0x33333333u) + ((bc >> 2) & 0x33333333u);
                                bc = (bc + (bc >> 4)) & 0x0f0f0f0fu;            // set bits per byte
                                const uint32_t incl = bc * 0x01010101u;          // byte j: set bits of bytes 0..j
                                // bit 7 of byte j is set iff incl_j > idx (incl_j + 127 - idx >= 128; no carry between bytes)
                                const uint32_t hit = (incl + (0x7fu - idx) * 0x01010101u) & 0x80808080u;
                                const uint32_t sh8 = ((uint32_t)__ffs((int)hit) - 1u) & ~7u; // 8 * (index of the byte)
                                const uint32_t below = ((incl << 8) >> sh8) & 0xffu;         // set bits in the bytes before it
                                cell = wbase + sh8 + s_nth[(word >> sh8) & 0xffu][idx - below];
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) { // word q receives the bit iff 0 <= cell - 32 q < 32: a clamped shift
                                const uint32_t add = shl_clamp(1u, cell - 32u * (uint32_t)q);
                                if (mover == 0) full[0].w[q] |= add;
                                else full[1].w[q] |= add;
                            }
                            order[ply][lane] = (uint8_t)cell;
                            --n_empty;
                        }
                    }
                }
                __syncwarp();

                // ---- 2. who: the position at ply D decides the winner ----
                // On a full board exactly one player connects (Hex has no draws), so when P0 does not, P1 does and
                // its flood can be skipped; a depth-cut position needs both tests.
                int winner = -1;
                {
                    const B128 c0 = H::flood(H::bor(sh.conn[0], H::first_edge(0)), full[0], NC0, NC10, BM);
                    if (H::any(H::band(c0, H::second_edge(0)))) {
                        winner = 0;
                    } else if (D == E) {
                        winner = 1;
                    } else {
                        const B128 c1 = H::flood(H::bor(sh.conn[1], H::first_edge(1)), full[1], NC0, NC10, BM);
                        if (H::any(H::band(c1, H::second_edge(1)))) winner = 1;
                    }
                }

                // ---- 3. when: first of the winner's moves that connects ----
                uint32_t depth = D;
                B128 win_mask = H::zero(); // the winner's stones at the stop ply
                if (winner >= 0) {
                    // the winner's moves are plies first, first+2, ...; move i is ply first + 2i
                    const uint32_t first = ((uint32_t)winner ^ leaf_turn) & 1u;
                    const int n_w = (int)((D - first + 1) >> 1);
                    int lo = -1, hi = n_w - 1; // not connected after move lo, connected after move hi
                    B128 mask_lo = sh.stones[winner], conn_lo = sh.conn[winner];
                    const B128 edge1 = H::first_edge(winner), edge2 = H::second_edge(winner);
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        B128 m = mask_lo;
                        for (int i = lo + 1; i <= mid; ++i) {
                            const uint32_t cell = order[first + 2u * (uint32_t)i][lane];
                            #pragma unroll
                            for (int q = 0; q < 4; ++q) m.w[q] |= shl_clamp(1u, cell - 32u * (uint32_t)q);
                        }
                        // everything connected at `lo` stays connected; grow it through the added stones
                        const B128 c = H::flood(H::bor(conn_lo, edge1), m, NC0, NC10, BM);
                        if (H::any(H::band(c, edge2))) {
                            hi = mid;
                        } else {
                            lo = mid;
                            mask_lo = m;
                            conn_lo = c;
                        }
                    }
                    depth = first + 2u * (uint32_t)hi + 1u;
                    if (NEED_BOARDS) {
                        win_mask = mask_lo;
                        for (int i = lo + 1; i <= hi; ++i) {
                            const uint32_t cell = order[first + 2u * (uint32_t)i][lane];
#pragma unroll
                            for (int q = 0; q < 4; ++q) win_mask.w[q] |= shl_clamp(1u, cell - 32u * (uint32_t)q);
                        }
                    }
                }

                // ---- 4. results ----
                const bool cut = winner < 0 && D < E; // no winner within max_playout_depth plies
                if (active) {
                    acc_w0 += (winner == 0);
                    acc_w1 += (winner == 1);
                    acc_draw += (winner < 0);
                    acc_cut += cut;
                    acc_depth += depth;
                }
                B128 fin[2];
                if (NEED_BOARDS) {
                    // both players' stones at the stop ply: the winner's from the search, the other's by replay
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        if (pl == winner) {
                            fin[pl] = win_mask;
                        } else if (winner < 0) {
                            fin[pl] = full[pl];
                        } else {
                            B128 m = sh.stones[pl];
                            const uint32_t first = ((uint32_t)pl ^ leaf_turn) & 1u;
                            for (uint32_t ply = first; ply < depth; ply += 2) {
                                const uint32_t cell = order[ply][lane];
                                #pragma unroll
                                for (int q = 0; q < 4; ++q) m.w[q] |= shl_clamp(1u, cell - 32u * (uint32_t)q);
                            }
                            fin[pl] = m;
                        }
                    }
                }
                if (active) {
                    const size_t g = (size_t)li * p.k + pid;
                    if (WANT_TRACE && p.trace) {
                        const uint32_t nt = min(depth, p.max_trace);
                        for (uint32_t t = 0; t < nt; ++t) p.trace[g * p.max_trace + t] = order[t][lane];
                    }
                    if (WANT_TRACE && p.rec) {
                        mr_playout_record rr;
                        rr.depth = depth;
                        rr.outcome = (uint8_t)(winner == 0 ? MR_WINNER_P0 : (winner == 1 ? MR_WINNER_P1 : MR_DRAW));
                        rr.end_type = (uint8_t)(cut ? MR_END_TURN_LIMIT : MR_END_TERMINAL);
                        rr.pad[0] = rr.pad[1] = 0;
                        p.rec[g] = rr;
                    }
                    if (WANT_FINAL && p.final_state) {
                        mr_leaf *out = &p.final_state[g];
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) {
                            out->board[2 * pl] = mk64(fin[pl].w[0], fin[pl].w[1]);
                            out->board[2 * pl + 1] = mk64(fin[pl].w[2], fin[pl].w[3]);
                        }
                        out->turn = (uint8_t)((leaf_turn + depth) & 1u);
                        out->flag = 0;
                        out->prev_action_plus1 = 0;
                        out->prev2_action_plus1 = 0;
                        out->pad[0] = out->pad[1] = 0;
                    }
                }
                if ((WANT_AMAF && p.amaf) || (WANT_MAST && p.mast_delta)) {
                    // occurrences = stones at the stop ply minus leaf stones (every cell is played at most
                    // once); lane L owns cells L, L+32, L+64, L+96 of both players: no atomics.  The per-leaf AMAF table
                    // (backprop.rs:619-640) and the batch-wide MAST delta (:857-870) count the same occurrences.
                    __syncwarp();
                    const uint32_t n_act = min(32u, pid_end - round);
                    for (uint32_t f = 0; f < n_act; ++f) {
                        const uint32_t d_f = __shfl_sync(FULL_MASK, depth, f);
                        const int w_f = __shfl_sync(FULL_MASK, winner, f);
                        if (WANT_AMAF && p.amaf) {
                            if (amaf_plies + d_f > STAT_FLUSH_PLIES) {
                                flush_stat_table<H>(amaf_tab, p.amaf + (size_t)li * p.amaf_stride, lane, p.tables_in_slots);
                                amaf_plies = 0;
                            }
                            amaf_plies += d_f;
                        }
                        if (WANT_MAST && p.mast_delta) {
                            if (mast_plies + d_f > STAT_FLUSH_PLIES) {
                                flush_stat_table<H>(mast_tab, p.mast_delta, lane, p.tables_in_slots);
                                mast_plies = 0;
                            }
                            mast_plies += d_f;
                        }
#pragma unroll
                        for (int pl = 0; pl < 2; ++pl) {
                            const int u = w_f < 0 ? 0 : (w_f == pl ? 1 : -1);
                            const uint32_t inc = 1u + ((uint32_t)u << 16);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t played = __shfl_sync(FULL_MASK, fin[pl].w[q], f) & ~sh.stones[pl].w[q];
                                if ((played >> lane) & 1u) {
                                    if (WANT_AMAF && p.amaf) amaf_tab[pl * H::NSLOTS + q * 32 + lane] += inc;
                                    if (WANT_MAST && p.mast_delta) mast_tab[pl * H::NSLOTS + q * 32 + lane] += inc;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                __syncwarp();
            }
            if (WANT_AMAF && p.amaf) {
                flush_stat_table<H>(amaf_tab, p.amaf + (size_t)li * p.amaf_stride, lane, p.tables_in_slots);
                amaf_plies = 0;
            }
        }

        acc_w0 = __reduce_add_sync(FULL_MASK, acc_w0);
        acc_w1 = __reduce_add_sync(FULL_MASK, acc_w1);
        acc_draw = __reduce_add_sync(FULL_MASK, acc_draw);
        acc_cut = __reduce_add_sync(FULL_MASK, acc_cut);
        acc_depth = __reduce_add_sync(FULL_MASK, acc_depth);
        if (lane == 0) {
            mr_leaf_result *res = p.results + li;
            if (acc_w0) atomicAdd(&res->wins[0], acc_w0);
            if (acc_w1) atomicAdd(&res->wins[1], acc_w1);
            if (acc_draw) atomicAdd(&res->draws, acc_draw);
            if (acc_cut) atomicAdd(&res->cutoffs, acc_cut);
            if (acc_depth) atomicAdd(reinterpret_cast<unsigned long long *>(&res->sum_depth), (unsigned long long)acc_depth);
        }
    }
    // the MAST delta is batch-wide: it leaves shared memory when the warp runs out of tasks
    if (WANT_MAST && p.mast_delta) flush_stat_table<H>(mast_tab, p.mast_delta, lane, p.tables_in_slots);
    batch_epilogue(p);
}

} // namespace mr
