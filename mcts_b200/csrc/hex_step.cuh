// hex_step.cuh -- Hex 11x11 as a per-ply register state machine for rollout_kernel (sm_100a).
//
// hex_kernel.cuh's fill-then-search is the fast path for the UNIFORM policy, whose moves never look at the
// position.  Policies that do (MAST, decisive moves, Last Good Reply) need the position after every ply; this
// game struct gives them one, with the same rules (games/hex-gen/src/lib.rs:115-144, 208-215) and the same
// action order (empty cells ascending).
//
// The reference answers "did the mover just connect its edges?" with a flood6 fix-point over the mover's stones
// after every ply (lib.rs:130-144 via game.rs:202-211).  Here each lane keeps, per player, the two sets
//     e1 = stones connected to the player's FIRST edge,   e2 = stones connected to its SECOND edge
// (P0: north row 10 / south row 0; P1: west col 0 / east col 10) and updates them incrementally: a new stone x
// joins e1 iff it lies on the edge or touches a stone of e1, and then drags in the not-yet-connected own stones
// reachable from it (a flood that only runs when something changes).  The mover has connected iff x is now in
// both sets -- the same predicate as the reference's flood from one edge reaching the other.
// The sets also give the decisive policies their answer without trying moves: a cell wins for a player iff it is
// empty and touches (or lies on) both of that player's edge-connected sets.
#pragma once

#include "games_device.cuh"

namespace mr {

struct HexStep : Hex11 {
    static constexpr bool MASK_AMAF = false; // statistics come from the slot trace, like the other per-ply games
    template <class S>
    __device__ static void init_shared(S &, uint32_t) {} // no per-warp tables
    static constexpr bool ROLLED = true;
    __host__ __device__ static constexpr bool rolled(int) { return ROLLED; }
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int min_blocks(int policy) { return MIN_BLOCKS; }
    __device__ static uint32_t code_to_slot(uint32_t pl, uint32_t code) { return code < 121u ? code : 0u; }

    struct Shared {
        B128 stones[2];
        B128 e1[2], e2[2];
        uint32_t turn, n_empty;
    };

    B128 s[2], e1[2], e2[2]; // indexed by compile-time player only (registers)
    uint32_t n_empty;
    uint32_t rt_parity;
    uint32_t turn_u; // the leaf's mover as a warp-uniform value (see Connect4::turn_u)
    uint32_t lgr_reply;
    uint32_t lgr_reply2; // MR_EPS_LGR2: the 2-ply table's reply, tried before lgr_reply
    const uint64_t (*nst_planes)[MAST_BITS]; // MR_EPS_NST: this ply's context rank planes, set by the kernel
    uint32_t nst_nbits;
    const uint16_t *nst_ranks;

    __device__ static constexpr B128 nc0() { return inv(col_mask(0)); }
    __device__ static constexpr B128 nc10() { return inv(col_mask(10)); }
    __device__ static B128 nb(B128 f) {
        constexpr B128 NC0 = inv(col_mask(0)), NC10 = inv(col_mask(10)), BM = board_mask();
        return neighbours(f, NC0, NC10, BM);
    }
    __device__ static B128 flood_from(B128 seed, B128 own) {
        constexpr B128 NC0 = inv(col_mask(0)), NC10 = inv(col_mask(10)), BM = board_mask();
        return flood(seed, own, NC0, NC10, BM);
    }
    __device__ static uint32_t count(B128 a) { return __popc(a.w[0]) + __popc(a.w[1]) + __popc(a.w[2]) + __popc(a.w[3]); }
    __device__ static B128 cell_bit(uint32_t cell) {
        return B128{{shl_clamp(1u, cell), shl_clamp(1u, cell - 32u), shl_clamp(1u, cell - 64u), shl_clamp(1u, cell - 96u)}};
    }
    __device__ static bool has_cell(B128 a, uint32_t cell) {
        const uint32_t ws = cell >> 5;
        const uint32_t w = ws == 0 ? a.w[0] : (ws == 1 ? a.w[1] : (ws == 2 ? a.w[2] : a.w[3]));
        return (w >> (cell & 31u)) & 1u;
    }
    __device__ static uint32_t first_cell(B128 a) {
        if (a.w[0]) return (uint32_t)__ffs((int)a.w[0]) - 1u;
        if (a.w[1]) return 32u + (uint32_t)__ffs((int)a.w[1]) - 1u;
        if (a.w[2]) return 64u + (uint32_t)__ffs((int)a.w[2]) - 1u;
        return 96u + (uint32_t)__ffs((int)a.w[3]) - 1u;
    }
    // the idx-th set cell, ascending
    __device__ static uint32_t select128(B128 e, uint32_t idx) {
        const uint32_t c0 = __popc(e.w[0]), c1 = __popc(e.w[1]), c2 = __popc(e.w[2]);
        uint32_t base = 0, word = e.w[0];
        if (idx >= c0) {
            idx -= c0;
            base = 32;
            word = e.w[1];
            if (idx >= c1) {
                idx -= c1;
                base = 64;
                word = e.w[2];
                if (idx >= c2) {
                    idx -= c2;
                    base = 96;
                    word = e.w[3];
                }
            }
        }
        return base + select32(word, idx);
    }
    // number of set cells below `cell`
    __device__ static uint32_t rank128(B128 e, uint32_t cell) {
        const uint32_t ws = cell >> 5, below = (1u << (cell & 31u)) - 1u;
        uint32_t r = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) r += (uint32_t)q < ws ? __popc(e.w[q]) : ((uint32_t)q == ws ? __popc(e.w[q] & below) : 0u);
        return r;
    }

    __device__ static int prepare(Shared &sh, const mr_leaf &leaf, uint32_t lane) {
        B128 st[2], a1[2], a2[2];
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            st[pl].w[0] = (uint32_t)leaf.board[2 * pl];
            st[pl].w[1] = (uint32_t)(leaf.board[2 * pl] >> 32);
            st[pl].w[2] = (uint32_t)leaf.board[2 * pl + 1];
            st[pl].w[3] = (uint32_t)(leaf.board[2 * pl + 1] >> 32) & 0x01ffffffu;
            a1[pl] = flood_from(first_edge(pl), st[pl]);
            a2[pl] = flood_from(second_edge(pl), st[pl]);
        }
        const uint32_t n_occ = count(bor(st[0], st[1]));
        if (lane == 0) {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
                sh.stones[pl] = st[pl];
                sh.e1[pl] = a1[pl];
                sh.e2[pl] = a2[pl];
            }
            sh.turn = leaf.turn;
            sh.n_empty = 121 - n_occ;
        }
        const int last = (leaf.turn + 1) & 1; // only the player who moved last is checked (lib.rs:130-144)
        if (any(band(a1[last], second_edge(last)))) return last ? MR_WINNER_P1 : MR_WINNER_P0;
        if (n_occ == 121) return MR_DRAW;
        return MR_NOT_TERMINAL;
    }
    __device__ void reset(const Shared &sh) {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            s[pl] = sh.stones[pl];
            e1[pl] = sh.e1[pl];
            e2[pl] = sh.e2[pl];
        }
        n_empty = sh.n_empty;
    }

    // cells that would connect player PL's edges at once: empty, on-or-touching e1 and on-or-touching e2
    template <int PL>
    __device__ B128 win_cells(B128 empties) const {
        const B128 t1 = bor(nb(e1[PL]), first_edge(PL)), t2 = bor(nb(e2[PL]), second_edge(PL));
        return band(empties, band(t1, t2));
    }

    template <int POLICY, int PL>
    __device__ int step_pl(uint32_t x0, uint32_t x1, const KernelParams &p, uint32_t &action, uint32_t &slot) {
        constexpr int OP = 1 - PL;
        constexpr bool DECISIVE = POLICY == MR_DECISIVE_WIN || POLICY == MR_DECISIVE_ANTI || POLICY == MR_DECISIVE_MAST;
        constexpr bool MAST_INNER = policy_mast_inner(POLICY);
        constexpr B128 BM = board_mask();
        const B128 empties = bandn(BM, bor(s[0], s[1]));
        const uint32_t n = n_empty;
        if (n == 0) return ST_NO_ACTIONS; // (a full board is terminal and never reaches here)
        uint32_t cell = 0;
        bool have = false;
        if (DECISIVE) {
            // DecisiveMove::choose (simulate.rs:636-735).  A Hex move can end the game only by connecting the mover
            // (a full board without a connection of the last mover does not occur in play), so Win, WinLoss and
            // WinLossDraw all take the first winning cell.
            const B128 mine = win_cells<PL>(empties);
            if (any(mine)) {
                cell = first_cell(mine);
                have = true;
            } else if (POLICY == MR_DECISIVE_ANTI && n > 1) {
                // pass 2: the first cell after which the opponent has no winning reply.  Our stone only takes one
                // cell away from the opponent: no threat -> the first empty cell; one -> that cell; more -> none.
                const B128 theirs = win_cells<OP>(empties);
                const uint32_t nt = count(theirs);
                if (nt <= 1) {
                    cell = first_cell(nt == 0 ? empties : theirs);
                    have = true;
                }
            }
        }
        if (policy_is_lgr(POLICY)) { // the recorded reply, still empty (simulate.rs:1022-1026); LGRF-2's 2-ply reply first
            if (policy_is_lgr2(POLICY) && lgr_reply2 && lgr_reply2 <= 121u && has_cell(empties, lgr_reply2 - 1u)) {
                cell = lgr_reply2 - 1u;
                have = true;
            } else if (lgr_reply && lgr_reply <= 121u && has_cell(empties, lgr_reply - 1u)) {
                cell = lgr_reply - 1u;
                have = true;
            }
        }
        if (MAST_INNER && !have && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
            // Mast::select_move (simulate.rs:824-848): bit-sliced arg-max over the empty cells (the two 64-bit halves
            // of the board ride the "kind" dimension of the rank planes), then random_best's visiting order
            uint64_t cand[2] = {mk64(empties.w[0], empties.w[1]), mk64(empties.w[2], empties.w[3])};
            if (POLICY == MR_EPS_NST) mast_argmax<2>(cand, nst_planes, nst_nbits);
            else mast_argmax<2>(cand, p.mast_planes[PL], p.mast_nbits[PL]);
            const B128 c = B128{{lo32(cand[0]), hi32(cand[0]), lo32(cand[1]), hi32(cand[1])}};
            const uint32_t m = count(c);
            const uint32_t r = mulhi(x0, 16u * n);
            const uint32_t i0 = r >> 4;
            if (m == 1) {
                cell = first_cell(c);
            } else if (m == n) {
                cell = select128(empties, i0);
            } else if (m <= 8) {
                const uint32_t invs = MR_STRIDE_INV[n][r & 15u], rcp = MR_RCP32[n];
                uint32_t best_j = 0xffffffffu;
                B128 rest = c;
                for (uint32_t q = 0; q < m; ++q) {
                    const uint32_t c0 = first_cell(rest);
                    rest = bandn(rest, cell_bit(c0));
                    const uint32_t j = visiting_rank(rank128(empties, c0), i0, n, invs, rcp);
                    if (j < best_j) {
                        best_j = j;
                        cell = c0;
                    }
                }
            } else {
                const uint32_t stride = MR_STRIDE[n][r & 15u];
                uint32_t i = i0;
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t c0 = select128(empties, i);
                    if (has_cell(c, c0)) {
                        cell = c0;
                        break;
                    }
                    i += stride;
                    if (i >= n) i -= n;
                }
            }
            have = true;
        }
        if (!have) cell = select128(empties, mulhi(x0, n));

        // ---- apply (lib.rs:122-125) and the incremental connection test ----
        const B128 x = cell_bit(cell);
        s[PL] = bor(s[PL], x);
        --n_empty;
        action = cell;
        slot = cell;
        const B128 nx = nb(x);
        const bool t1 = any(band(x, first_edge(PL))) || any(band(nx, e1[PL]));
        const bool t2 = any(band(x, second_edge(PL))) || any(band(nx, e2[PL]));
        // x joins a set it touches; own stones next to x that were not in the set yet come along, with whatever they
        // reach -- usually there are none, and the flood (divergent, long) is skipped
        if (t1) {
            const B128 adj = bandn(band(nx, s[PL]), e1[PL]);
            e1[PL] = bor(e1[PL], x);
            if (any(adj)) e1[PL] = bor(e1[PL], flood_from(adj, bandn(s[PL], e1[PL])));
        }
        if (t2) {
            const B128 adj = bandn(band(nx, s[PL]), e2[PL]);
            e2[PL] = bor(e2[PL], x);
            if (any(adj)) e2[PL] = bor(e2[PL], flood_from(adj, bandn(s[PL], e2[PL])));
        }
        if (t1 && t2) return ST_MOVED | (PL ? ST_WIN_P1 : ST_WIN_P0);
        if (n_empty == 0) return ST_MOVED | ST_DRAW; // is_terminal on a full board, no connection of the mover (:208-215)
        return ST_MOVED;
    }

    template <int POLICY, int PARITY>
    __device__ int step(const Shared &sh, uint32_t x0, uint32_t x1, const KernelParams &p, uint32_t &action, uint32_t &slot) {
        // the mover is warp-uniform (one leaf per task, lanes in ply lock-step): the branch never diverges, and each
        // side's code indexes the per-player register sets with a compile-time constant
        const uint32_t pl = (turn_u ^ (PARITY == 2 ? rt_parity : (uint32_t)PARITY)) & 1u;
        if (pl) return step_pl<POLICY, 1>(x0, x1, p, action, slot);
        return step_pl<POLICY, 0>(x0, x1, p, action, slot);
    }

    __device__ void store_final(const Shared &sh, mr_leaf *out, uint32_t plies, int end_code) const {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            out->board[2 * pl] = mk64(s[pl].w[0], s[pl].w[1]);
            out->board[2 * pl + 1] = mk64(s[pl].w[2], s[pl].w[3]);
        }
        out->turn = (uint8_t)((sh.turn + plies) & 1u);
        out->flag = 0;
        out->prev_action_plus1 = 0;
        out->prev2_action_plus1 = 0;
        out->pad[0] = out->pad[1] = 0;
    }
};

} // namespace mr
