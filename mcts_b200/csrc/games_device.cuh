// games_device.cuh -- the five bitboard games as per-lane register state machines (sm_100a).
//
// Each game keeps the reference's RULES and ACTION ORDER (so that "the idx-th legal action" means the
// same action as in the reference's generate_actions list) but not its data layout or algorithm:
// moves are never materialised; legal-move sets live in bitboards, the idx-th action is found with
// popcount rank/select, and terminal tests are incremental.  Reference lines each rule follows are
// cited at the rule.
//
// Common contract (used by rollout_kernel.cuh):
//   struct Shared            per-warp leaf-derived data in shared memory
//   prepare(sh, leaf, lane)  all 32 lanes; fills sh; returns the leaf's own TerminalStatus (mr_outcome)
//   reset(sh)                lane state := leaf
//   step<POLICY, PARITY>(..) one ply for the lane; mover = sh.turn ^ PARITY; returns ST_* code
//   store_final(sh, out, parity_after)  debug: lane state -> reference-layout mr_leaf
#pragma once

#include <type_traits>

#include "common.cuh"

namespace mr {

__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

// a | b for disjoint a, b and a & ~b for b a subset of a, as carry-free 32-bit adds / subs (both integer pipes)
__device__ __forceinline__ uint64_t add_disjoint(uint64_t a, uint64_t b) { return mk64(lo32(a) + lo32(b), hi32(a) + hi32(b)); }
__device__ __forceinline__ uint64_t sub_subset(uint64_t a, uint64_t b) { return mk64(lo32(a) - lo32(b), hi32(a) - hi32(b)); }

// visiting rank j of list position i in random_best's order i = (i0 + j*stride) mod n (util.rs:147-161):
// j = ((i - i0) mod n) * stride^-1 mod n
__device__ __forceinline__ uint32_t visiting_rank(uint32_t i, uint32_t i0, uint32_t n, uint32_t inv, uint32_t rcp) {
    const uint32_t d = i >= i0 ? i - i0 : i + n - i0;
    const uint32_t t = d * inv;
    return t - __umulhi(t, rcp) * n;
}

// Position of the idx-th (0-based) set bit of a 32-bit word, by binary search on popcounts.
__device__ __forceinline__ uint32_t select32(uint32_t w, uint32_t idx) {
    uint32_t pos = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t m = ((1u << step) - 1u) << pos;
        select_step(idx, pos, __popc(w & m), (uint32_t)step);
    }
    return pos;
}

// ==========================================================================================
// Connect4 6x7 -- games/connect4/src/lib.rs
//   rules: legal = non-full columns ascending (:303-312); drop at height (:314-339); win = four in a
//   row N/E/NE/NW (:195-207); terminal = winner or 42 stones (:372-374); winner keeps the turn.
//   device layout: one BYTE per column (bit = col*8 + row), rows 6-7 of every byte stay zero, so
//   `all + 0x01..01` yields the drop cell of every column at once and shifted-AND line tests need no
//   wall masks.
// ==========================================================================================
struct Connect4 {
    static constexpr int NA = 7;
    // windows unrolled: measured 6-7% faster than a rolled window for this small step (profiles/roll_compare_r01.txt)
    static constexpr bool ROLLED = false;
    __host__ __device__ static constexpr bool rolled(int) { return ROLLED; }
    static constexpr int MIN_BLOCKS = 1; // a 64-register cap (8 blocks/SM) measured 4% slower
    static constexpr int min_blocks(int policy) { return MIN_BLOCKS; }
    static constexpr int NSLOTS = 7; // compact per-player statistic slots (== action ids here)
    static constexpr int KINDS = 1;
    static constexpr bool MASK_AMAF = false;
    template <class S>
    __device__ static void init_shared(S &, uint32_t) {} // no per-warp tables
    __device__ static uint32_t slot_action_id(uint32_t pl, uint32_t slot) { return slot; }
    __device__ static uint32_t code_to_slot(uint32_t pl, uint32_t code) { return code < 7u ? code : 0u; }
    static constexpr uint64_t BOTTOM = 0x0001010101010101ull;
    static constexpr uint64_t BOARD = 0x003F3F3F3F3F3F3Full;
    static constexpr uint64_t TOPROW = 0x0020202020202020ull;

    struct Shared {
        uint64_t cur; // stones of the player to move, column-byte layout
        uint64_t all;
        uint32_t turn;
        uint32_t n_open; // columns that are not full
    };

    uint64_t cur, all;
    uint32_t n_open;
    uint32_t rt_parity;
    uint32_t turn_u; // the leaf's mover, as a WARP-UNIFORM value (kernel: a REDUX result), so that everything derived from
                     // the mover's colour -- table indices, loop bounds, constant-bank addresses -- runs on the uniform datapath
    uint32_t lgr_reply; // MR_EPS_LGR: 1 + compact code of the recorded reply for this ply, 0 = none / exploring
    uint32_t lgr_reply2; // MR_EPS_LGR2: the 2-ply table's reply, tried before lgr_reply
    // MR_EPS_NST: this ply's context (mover, preceding action) rank planes / per-column ranks, set by the kernel
    const uint64_t (*nst_planes)[MAST_BITS];
    uint32_t nst_nbits;
    const uint16_t *nst_ranks;

    __device__ static uint64_t to_colbyte(uint64_t rm) { // reference row-major (idx = row*7 + col) -> col*8 + row
        uint64_t out = 0;
        for (int row = 0; row < 6; ++row)
            for (int col = 0; col < 7; ++col)
                if ((rm >> (row * 7 + col)) & 1) out |= 1ull << (col * 8 + row);
        return out;
    }
    __device__ static uint64_t to_rowmajor(uint64_t cb) {
        uint64_t out = 0;
        for (int row = 0; row < 6; ++row)
            for (int col = 0; col < 7; ++col)
                if ((cb >> (col * 8 + row)) & 1) out |= 1ull << (row * 7 + col);
        return out;
    }

    // four in a row through LEFT shifts: the low word of a 64-bit left shift is a multiply (IMAD.SHL, FMA pipe),
    // only the high word needs a funnel shift (ALU pipe) -- half the ALU work of right shifts, and the ALU pipe
    // is what bounds this kernel.  Spare rows 6-7 of every column byte (and byte 7) are zero, so nothing wraps.
    __device__ static bool four(uint64_t me) {
        uint64_t m, acc;
        m = me & (me << 1);
        acc = m & (m << 2);
        m = me & (me << 8);
        acc |= m & (m << 16);
        m = me & (me << 7);
        acc |= m & (m << 14);
        m = me & (me << 9);
        acc |= m & (m << 18);
        return acc != 0;
    }

    // cells where one more stone of `p` completes four in a row (occupied cells included; callers mask)
    __device__ static uint64_t win_cells(uint64_t p) {
        uint64_t r = (p << 1) & (p << 2) & (p << 3), pp;
#define MR_C4_LINE(S)                        \
    pp = (p << (S)) & (p << (2 * (S)));      \
    r |= pp & ((p << (3 * (S))) | (p >> (S))); \
    pp = (p >> (S)) & (p >> (2 * (S)));      \
    r |= pp & ((p << (S)) | (p >> (3 * (S))));
        MR_C4_LINE(8)
        MR_C4_LINE(7)
        MR_C4_LINE(9)
#undef MR_C4_LINE
        return r & BOARD;
    }

    __device__ static int prepare(Shared &sh, const mr_leaf &leaf, uint32_t lane) {
        // row-major -> column-byte, cooperatively: lane L converts cells L and L + 32, then a warp OR
        uint32_t blo = 0, bhi = 0, wlo = 0, whi = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t cell = lane + 32u * h; // reference index row*7 + col
            if (cell < 42) {
                const uint32_t row = cell / 7u, col = cell - row * 7u, pos = col * 8u + row;
                const uint32_t b = (uint32_t)(leaf.board[0] >> cell) & 1u, w = (uint32_t)(leaf.board[1] >> cell) & 1u;
                if (pos < 32) {
                    blo |= b << pos;
                    wlo |= w << pos;
                } else {
                    bhi |= b << (pos - 32);
                    whi |= w << (pos - 32);
                }
            }
        }
        const uint64_t black = mk64(__reduce_or_sync(FULL_MASK, blo), __reduce_or_sync(FULL_MASK, bhi));
        const uint64_t white = mk64(__reduce_or_sync(FULL_MASK, wlo), __reduce_or_sync(FULL_MASK, whi));
        const uint64_t all = black | white;
        if (lane == 0) {
            sh.cur = leaf.turn ? white : black;
            sh.all = all;
            sh.turn = leaf.turn;
            sh.n_open = __popcll((all + BOTTOM) & BOARD);
        }
        if (leaf.flag) return leaf.turn ? MR_WINNER_P1 : MR_WINNER_P0; // :380-386 winner == turn
        if (__popcll(all) == 42) return MR_DRAW;
        return MR_NOT_TERMINAL;
    }

    __device__ void reset(const Shared &sh) {
        cur = sh.cur;
        all = sh.all;
        n_open = sh.n_open;
    }

    template <int POLICY, int PARITY>
    __device__ int step(const Shared &sh, uint32_t x0, uint32_t x1, const KernelParams &p, uint32_t &action, uint32_t &slot) {
        const uint64_t moves = (all + BOTTOM) & BOARD; // one bit per open column, ascending column order
        uint64_t bit;
        constexpr bool MAST_INNER = policy_mast_inner(POLICY);
        if (POLICY == MR_DECISIVE_WIN || POLICY == MR_DECISIVE_MAST) {
            // first legal column (ascending) whose drop wins for the mover (simulate.rs:662-670); a Connect4
            // move cannot lose immediately, a drawing move exists only as the 42nd stone (the single action)
            uint64_t winning = 0;
            for (uint64_t m = moves; m; m &= m - 1) {
                const uint64_t b = m & (0 - m);
                if (four(cur | b)) {
                    winning = b;
                    break;
                }
            }
            bit = winning;
        } else if (POLICY == MR_DECISIVE_ANTI) {
            // DecisiveMove<AntiDecisive> (simulate.rs:689-735).  Pass 1: the first column whose drop wins.
            // Pass 2: the first column after which the opponent has no winning drop.  The opponent's winning
            // cells W do not depend on our stone except that it fills one cell and opens the one above:
            // with T = W & drop cells, no threat -> any drop x whose upper neighbour is not in W; one threat ->
            // only x = T (same condition); two or more -> none.  (A board-filling drop is the single action.)
            const uint64_t mine = win_cells(cur) & moves;
            if (mine) {
                bit = mine & (0 - mine);
            } else {
                const uint64_t W = win_cells(all ^ cur) & ~all;
                const uint64_t T = W & moves;
                const uint64_t cand = (T == 0 ? moves : ((T & (T - 1)) == 0 ? T : 0ull)) & ~(W >> 1);
                bit = cand & (0 - cand);
            }
        } else if (policy_is_lgr(POLICY)) {
            // the recorded reply column, if it is still open (simulate.rs:1022-1026); LGRF-2 tries its 2-ply reply first
            bit = (lgr_reply && lgr_reply <= 7u) ? (moves & (0xFFull << (8u * (lgr_reply - 1u)))) : 0ull;
            if (policy_is_lgr2(POLICY)) {
                const uint64_t b2 = (lgr_reply2 && lgr_reply2 <= 7u) ? (moves & (0xFFull << (8u * (lgr_reply2 - 1u)))) : 0ull;
                bit = b2 ? b2 : bit;
            }
        } else {
            bit = 0;
        }
        if (MAST_INNER && bit == 0 && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
            // Mast::select_move (simulate.rs:824-848) = random_best over the open columns: the first strict maximum in
            // visiting order, i.e. the largest (rank, -visiting rank) key.  At most 7 actions: every column is scored.
            const uint32_t pl = (turn_u ^ (PARITY == 2 ? rt_parity : (uint32_t)PARITY)) & 1u;
            const uint32_t r = mulhi(x0, 16u * n_open);
            const uint32_t i0 = r >> 4, inv = MR_STRIDE_INV[n_open][r & 15u], rcp = MR_RCP32[n_open];
            uint32_t best_key = 0, i = 0;
            bit = 0;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const uint64_t b = moves & (0xFFull << (8 * c));
                if (b) {
                    const uint32_t rk = POLICY == MR_EPS_NST ? nst_ranks[c] : p.mast_rank_small[pl][c];
                    const uint32_t key = (rk << 8) | (255u - visiting_rank(i, i0, n_open, inv, rcp));
                    if (key >= best_key) { // keys of open columns are distinct; >= lets a zero key win the first time
                        best_key = key;
                        bit = b;
                    }
                    ++i;
                }
            }
        }
        if (bit == 0) {
            uint32_t idx = mulhi(x0, n_open);
            const uint32_t lo = lo32(moves), hi = hi32(moves);
            const uint32_t nlo = __popc(lo);
            const bool in_hi = idx >= nlo;
            uint32_t w = in_hi ? hi : lo;
            idx = in_hi ? idx - nlo : idx;
            if (idx > 0) w &= w - 1;
            if (idx > 1) w &= w - 1;
            if (idx > 2) w &= w - 1;
            const uint32_t b = w & (0u - w);
            bit = in_hi ? mk64(0, b) : mk64(b, 0);
        }
        action = (uint32_t)(__ffsll((long long)bit) - 1) >> 3; // column
        slot = action;
        // `bit` is an empty cell: OR == ADD and, below, XOR with a subset == SUB, word by word without carries.
        // Adds issue on both integer pipes, LOP3 only on the (saturated) ALU pipe.
        all = add_disjoint(all, bit);
        const uint64_t me = add_disjoint(cur, bit);
        n_open -= (bit & TOPROW) ? 1u : 0u;
        cur = sub_subset(all, me); // the opponent moves next
        const uint32_t mover = (turn_u ^ (PARITY == 2 ? rt_parity : (uint32_t)PARITY)) & 1u;
        if (four(me)) return ST_MOVED | (mover ? ST_WIN_P1 : ST_WIN_P0);
        if (n_open == 0) return ST_MOVED | ST_DRAW;
        return ST_MOVED;
    }

    // parity_next: (plies played) & 1
    __device__ void store_final(const Shared &sh, mr_leaf *out, uint32_t plies, int end_code) const {
        // `cur` is the stones of the player who would move next (or, after a win, of the loser)
        const uint32_t next = (sh.turn + plies) & 1;
        const uint64_t nxt = to_rowmajor(cur), oth = to_rowmajor(all ^ cur);
        const bool won = end_code == ST_WIN_P0 || end_code == ST_WIN_P1;
        out->board[0] = next == 0 ? nxt : oth;
        out->board[1] = next == 0 ? oth : nxt;
        out->board[2] = out->board[3] = 0;
        // the winner keeps the turn (:332-336)
        out->turn = (uint8_t)(won ? next ^ 1 : next);
        out->flag = won ? 1 : 0;
        out->prev_action_plus1 = 0;
        out->prev2_action_plus1 = 0;
        out->pad[0] = out->pad[1] = 0;
    }
};

// ==========================================================================================
// Breakthrough 8x8 -- games/breakthrough/src/lib.rs
//   moves (:131-156): for src ascending, dst ascending = {west diagonal, straight, east diagonal};
//   diagonals may capture (dst not own), straight needs an empty cell.  Black moves south (-8),
//   White north (+8).  apply (:158-188): far wall reached => winner, turn kept.  No capture-all win;
//   a side without moves ends the playout as a draw (simulate.rs:112-116).
//   device: three SOURCE-space masks W,F,E; per-cell move count = W+F+E as a bit-sliced 2-bit number
//   (S = xor, C = majority); the idx-th move in reference order is a popcount rank/select over (S,C).
// ==========================================================================================
struct ThroughBase {
    static constexpr int NA = 4096;
    static constexpr bool MASK_AMAF = false;
    template <class S>
    __device__ static void init_shared(S &, uint32_t) {} // no per-warp tables
    static constexpr uint64_t NOT_COL0 = 0xFEFEFEFEFEFEFEFEull;
    static constexpr uint64_t NOT_COL7 = 0x7F7F7F7F7F7F7F7Full;

    struct Shared {
        uint64_t black, white;
        uint32_t turn;
    };

    uint64_t black, white;
    uint32_t rt_parity;
    uint32_t turn_u; // the leaf's mover as a warp-uniform value (see Connect4::turn_u)
    uint32_t lgr_reply; // MR_EPS_LGR: 1 + compact code (src << 8 | dst) of the recorded reply, 0 = none / exploring
    uint32_t lgr_reply2; // MR_EPS_LGR2: the 2-ply table's reply, tried before lgr_reply
    // MR_EPS_NST: this ply's context (mover, preceding action) rank planes / per-column ranks, set by the kernel
    const uint64_t (*nst_planes)[MAST_BITS];
    uint32_t nst_nbits;
    const uint16_t *nst_ranks;

    __device__ static int prepare(Shared &sh, const mr_leaf &leaf, uint32_t lane) {
        if (lane == 0) {
            sh.black = leaf.board[0];
            sh.white = leaf.board[1];
            sh.turn = leaf.turn;
        }
        if (leaf.flag) return leaf.turn ? MR_WINNER_P1 : MR_WINNER_P0;
        return MR_NOT_TERMINAL;
    }
    __device__ void reset(const Shared &sh) {
        black = sh.black;
        white = sh.white;
    }
    __device__ void store_final(const Shared &sh, mr_leaf *out, uint32_t plies, int end_code) const {
        const uint32_t next = (sh.turn + plies) & 1;
        const bool won = end_code == ST_WIN_P0 || end_code == ST_WIN_P1;
        out->board[0] = black;
        out->board[1] = white;
        out->board[2] = out->board[3] = 0;
        out->turn = (uint8_t)(won ? next ^ 1 : next);
        out->flag = won ? 1 : 0;
        out->prev_action_plus1 = 0;
        out->prev2_action_plus1 = 0;
        out->pad[0] = out->pad[1] = 0;
    }
    // own/opp update for a move src -> dst by `white`; returns true when the far wall is reached
    __device__ bool apply_move(bool is_white, uint32_t src, uint32_t dst) {
        const uint64_t s = bit64(src), d = bit64(dst);
        // src holds an own piece, dst does not: (own | d) & ~s == own + d - s, carry-free per 32-bit word
        if (is_white) {
            white = sub_subset(add_disjoint(white, d), s);
            black &= ~d;
            return dst >= 56;
        } else {
            black = sub_subset(add_disjoint(black, d), s);
            white &= ~d;
            return dst < 8;
        }
    }
};

struct Breakthrough : ThroughBase {
    static constexpr bool ROLLED = false;
    __host__ __device__ static constexpr bool rolled(int) { return ROLLED; }
#ifndef MR_BT_BLOCKS
#define MR_BT_BLOCKS 6
#endif
    // <= 80 registers, 24 warps/SM.  Measured on the MAST kernel after the pipe-balancing / uniform-datapath changes:
    // 4 blocks 4.12 ms, 5: 4.11, 6: 3.73, 7: 3.76, 8: 3.80 (the 7-block cap that was best before them is no longer)
    static constexpr int MIN_BLOCKS = MR_BT_BLOCKS;
    static constexpr int min_blocks(int policy) { return MIN_BLOCKS; }
    static constexpr int NSLOTS = 192;
    static constexpr int KINDS = 3; // src*3 + kind (W,F,E)
    __device__ static uint32_t slot_action_id(uint32_t pl, uint32_t slot) {
        const uint32_t src = slot / 3u, kind = slot - src * 3u;
        return src * 64u + (uint32_t)((int)src + (pl ? 7 : -9) + (int)kind);
    }
    // compact code (src << 8 | dst) of a move by `pl` -> its slot; 0 for a code that is not such a move
    __device__ static uint32_t code_to_slot(uint32_t pl, uint32_t code) {
        const uint32_t src = code >> 8, dst = code & 0xffu;
        const uint32_t kind = dst - src - (pl ? 7u : (uint32_t)-9);
        return (src < 64u && kind < 3u) ? src * 3u + kind : 0u;
    }
    // source-space masks of the three move kinds for the mover
    __device__ static void masks(bool is_white, uint64_t own, uint64_t opp, uint64_t &W, uint64_t &F, uint64_t &E) {
        const uint64_t occ = own | opp;
        if (is_white) { // dst = src+7 (W), src+8 (F), src+9 (E)
            F = own & (~occ >> 8);
            W = own & NOT_COL0 & (~own >> 7);
            E = own & NOT_COL7 & (~own >> 9);
        } else { // dst = src-9 (W), src-8 (F), src-7 (E)
            F = own & (~occ << 8);
            W = own & NOT_COL0 & (~own << 9);
            E = own & NOT_COL7 & (~own << 7);
        }
    }

    // rank/select: the idx-th move in (src ascending; W,F,E) order.  Returns src, sets dir (0=W,1=F,2=E).
    __device__ static uint32_t select_move(uint64_t W, uint64_t F, uint64_t E, uint64_t S, uint64_t C, uint32_t idx,
                                           uint32_t &dir) {
        const uint32_t cl = __popc(lo32(S)) + 2u * __popc(lo32(C));
        const bool in_hi = idx >= cl;
        idx = in_hi ? idx - cl : idx;
        const uint32_t s = in_hi ? hi32(S) : lo32(S), c = in_hi ? hi32(C) : lo32(C);
        uint32_t pos = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const uint32_t m = ((1u << step) - 1u) << pos;
            const uint32_t cnt = __popc(s & m) + 2u * __popc(c & m);
            select_step(idx, pos, cnt, (uint32_t)step);
        }
        const uint32_t w = ((in_hi ? hi32(W) : lo32(W)) >> pos) & 1u;
        const uint32_t f = ((in_hi ? hi32(F) : lo32(F)) >> pos) & 1u;
        // idx is now the residual among this cell's set kinds in W,F,E order
        if (w && idx == 0) {
            dir = 0;
        } else {
            idx -= w;
            dir = (f && idx == 0) ? 1u : 2u;
        }
        return pos + (in_hi ? 32u : 0u);
    }

    // random_best over FEW maximal actions (util.rs:147-161): the pick is the maximal action with the smallest visiting
    // rank.  One 32-bit half of the board at a time (HI = sources 32..63; `before` = the legal actions of the sources in
    // front of this half): the tied actions are as a rule the winning moves of ONE piece, all with score exactly 1
    // (measured with an instrumented oracle: every one of the 12.5 % of plies with a small tie on the headline workload),
    // so one half has the candidates and a 64-bit walk would do twice the work.  Cell by cell: the prefix count of the
    // cell gives the visiting rank of its first kind, every further legal kind is one list position on, i.e. + inv
    // (mod n) in visiting rank.  Returns the minimum of `best` and the packed keys (rank << 8 | source << 2 | kind).
    template <bool HI>
    __device__ static uint32_t half_tie(uint32_t best, const uint64_t cand[3], uint64_t W, uint64_t F, uint64_t S, uint64_t C,
                                        uint32_t before, uint32_t n, uint32_t i0, uint32_t inv, uint32_t rcp) {
        const uint32_t c0 = HI ? hi32(cand[0]) : lo32(cand[0]), c1 = HI ? hi32(cand[1]) : lo32(cand[1]);
        const uint32_t c2 = HI ? hi32(cand[2]) : lo32(cand[2]);
        const uint32_t w = HI ? hi32(W) : lo32(W), f = HI ? hi32(F) : lo32(F);
        const uint32_t sw = HI ? hi32(S) : lo32(S), cw = HI ? hi32(C) : lo32(C);
        for (uint32_t cells = c0 | c1 | c2; cells; cells &= cells - 1u) {
            const uint32_t b = (uint32_t)__ffs((int)cells) - 1u;
            const uint32_t below = (1u << b) - 1u;
            const uint32_t i = before + (uint32_t)(__popc(sw & below) + 2 * __popc(cw & below));
            const uint32_t key = ((HI ? 32u : 0u) + b) << 2;
            uint32_t j = visiting_rank(i, i0, n, inv, rcp);
            if ((c0 >> b) & 1u) best = min(best, (j << 8) | key);
            if ((w >> b) & 1u) {
                j += inv;
                j = min(j, j - n); // j - n wraps to a huge value while j < n
            }
            if ((c1 >> b) & 1u) best = min(best, (j << 8) | key | 1u);
            if ((f >> b) & 1u) {
                j += inv;
                j = min(j, j - n);
            }
            if ((c2 >> b) & 1u) best = min(best, (j << 8) | key | 2u);
        }
        return best;
    }

    template <int POLICY, int PARITY>
    __device__ int step(const Shared &sh, uint32_t x0, uint32_t x1, const KernelParams &p, uint32_t &action, uint32_t &slot);
};

// ==========================================================================================
// Knightthrough 8x8 -- games/knightthrough/src/lib.rs:126-190
//   same state/apply/goal as Breakthrough; moves = all eight knight jumps staying on the board minus own
//   pieces, src ascending then dst ascending, i.e. jump deltas in the order -17,-15,-10,-6,+6,+10,+15,+17.
// ==========================================================================================
struct Knightthrough : ThroughBase {
    static constexpr int MIN_BLOCKS = 5; // <= 96 registers: +13% on the MAST kernel (measured); 4 and 6 are within noise of it
    static constexpr int min_blocks(int policy) { return MIN_BLOCKS; }
    // the step is large (8 masks, adder tree): the MAST kernels run the window as a loop (fits the I-cache, +1.3 %
    // measured); the kernels without the arg-max are 4 % faster unrolled (compile-time ply parity, uniform datapath)
    __host__ __device__ static constexpr bool rolled(int policy) { return policy_mast_inner(policy); }
    static constexpr int NSLOTS = 512; // src*8 + kind (ascending-dst jump order)
    static constexpr int KINDS = 8;
    __device__ static uint32_t slot_action_id(uint32_t pl, uint32_t slot) {
        const uint32_t src = slot >> 3, kind = slot & 7u;
        return src * 64u + (uint32_t)((int)src + delta(kind));
    }
    __device__ static uint32_t code_to_slot(uint32_t pl, uint32_t code) {
        const uint32_t src = code >> 8, dst = code & 0xffu;
        uint32_t kind = 8;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if ((int)dst - (int)src == delta((uint32_t)q)) kind = q;
        return (src < 64u && kind < 8u) ? src * 8u + kind : 0u;
    }
    __device__ static void masks(uint64_t own, uint64_t M[8]) {
        const uint64_t n = ~own;
        constexpr uint64_t GE1 = 0xFEFEFEFEFEFEFEFEull, LE6 = 0x7F7F7F7F7F7F7F7Full;
        constexpr uint64_t GE2 = 0xFCFCFCFCFCFCFCFCull, LE5 = 0x3F3F3F3F3F3F3F3Full;
        M[0] = own & GE1 & (n << 17); // (-2,-1)
        M[1] = own & LE6 & (n << 15); // (-2,+1)
        M[2] = own & GE2 & (n << 10); // (-1,-2)
        M[3] = own & LE5 & (n << 6);  // (-1,+2)
        M[4] = own & GE2 & (n >> 6);  // (+1,-2)
        M[5] = own & LE5 & (n >> 10); // (+1,+2)
        M[6] = own & GE1 & (n >> 15); // (+2,-1)
        M[7] = own & LE6 & (n >> 17); // (+2,+1)
    }
    __device__ static int delta(uint32_t k) {
        // -17,-15,-10,-6,+6,+10,+15,+17 packed as bytes biased by +17
        const uint64_t tbl = 0x22201B170B070200ull;
        return (int)((tbl >> (8 * k)) & 0xff) - 17;
    }

    template <int POLICY, int PARITY>
    __device__ int step(const Shared &sh, uint32_t x0, uint32_t x1, const KernelParams &p, uint32_t &action, uint32_t &slot);
};

// ---- MAST greedy pick (Breakthrough / Knightthrough) ----
// Mast::select_move = random_best over unigram scores (simulate.rs:794-848, util.rs:135-163): visit list
// positions i_j = (i0 + j*stride) mod n and keep the FIRST strict maximum, i.e. the first position in
// visiting order whose score equals the global maximum.
//
// Scores arrive as bit-planes of dense order-preserving ranks (KernelParams::mast_planes), so the set of
// maximal legal actions is found by a bit-sliced arg-max: start from the legal-move masks and, from the
// most significant rank bit down, keep only candidates with that bit set whenever any has it.
// Three outcomes:  exactly one maximal action -> it is the pick, no list index needed;
//                  every legal action maximal (e.g. an empty table) -> the pick is list position i0;
//                  otherwise -> mark the maximal actions' list positions in a bitmask and walk the
//                  visiting order until a marked position is met.
template <int KINDS>
__device__ __forceinline__ void mast_argmax(uint64_t cand[KINDS], const uint64_t (*planes)[MAST_BITS], uint32_t nbits) {
    // (the "keep the candidates that have the bit" update is a conditional move on `any`: predicated IMADs on the FMA
    // pipe instead of SELs on the ALU pipe, +7 % on the Breakthrough MAST kernel, measured)
    uint32_t lo[KINDS], hi[KINDS];
#pragma unroll
    for (int q = 0; q < KINDS; ++q) {
        lo[q] = lo32(cand[q]);
        hi[q] = hi32(cand[q]);
    }
    for (int b = (int)nbits - 1; b >= 0; --b) {
        uint32_t tl[KINDS], th[KINDS];
        uint32_t any = 0;
#pragma unroll
        for (int q = 0; q < KINDS; ++q) {
            const uint64_t pl = planes[q][b];
            tl[q] = lo[q] & lo32(pl);
            th[q] = hi[q] & hi32(pl);
            any |= tl[q] | th[q];
        }
#pragma unroll
        for (int q = 0; q < KINDS; ++q) {
            cmov_nz(lo[q], tl[q], any);
            cmov_nz(hi[q], th[q], any);
        }
    }
#pragma unroll
    for (int q = 0; q < KINDS; ++q) cand[q] = mk64(lo[q], hi[q]);
}

template <int POLICY, int PARITY>
__device__ int Breakthrough::step(const Shared &sh, uint32_t x0, uint32_t x1, const KernelParams &p,
                                  uint32_t &action, uint32_t &slot) {
    const bool is_white = ((turn_u ^ (PARITY == 2 ? rt_parity : (uint32_t)PARITY)) & 1u) != 0;
    const uint64_t own = is_white ? white : black, opp = is_white ? black : white;
    uint64_t W, F, E;
    masks(is_white, own, opp, W, F, E);
    const uint64_t S = W ^ F ^ E, C = (W & F) | (W & E) | (F & E);
    const uint32_t n = __popcll(S) + 2u * __popcll(C);
    if (n == 0) return ST_NO_ACTIONS;
    uint32_t idx = 0, src = 0, dir = 0;
    bool have = false, idx_set = false;
    // DecisiveMove<.., inner> first tries the decisive scan and falls through to its inner strategy (simulate.rs:742-760)
    constexpr bool DECISIVE = POLICY == MR_DECISIVE_WIN || POLICY == MR_DECISIVE_ANTI || POLICY == MR_DECISIVE_MAST;
    constexpr bool MAST_INNER = policy_mast_inner(POLICY);
    if (DECISIVE) {
        // DecisiveMove<Win> (simulate.rs:662-687): the first action in list order whose child is a win for the
        // mover.  A Breakthrough move wins iff it lands on the far wall, i.e. iff its source is on the row next
        // to it; no move can lose or draw on the spot, so the loser/draw fallbacks never fire.
        const uint64_t last_row = is_white ? 0x00FF000000000000ull : 0x000000000000FF00ull;
        const uint64_t any_win = (W | F | E) & last_row;
        if (any_win) {
            src = (uint32_t)__ffsll((long long)any_win) - 1u;
            dir = ((W >> src) & 1ull) ? 0u : (((F >> src) & 1ull) ? 1u : 2u);
            have = true;
        } else if (POLICY == MR_DECISIVE_ANTI) {
            // AntiDecisive pass 2 (simulate.rs:701-733): the first action after which the opponent cannot win at
            // once.  An opposing piece one row from ITS goal always has a legal diagonal onto the goal row (no own
            // piece can stand there), so the opponent can win iff such a piece survives our move: none -> the
            // first legal action; exactly one -> the first action that captures it; more -> the inner pick.
            const uint64_t threat = opp & (is_white ? 0x000000000000FF00ull : 0x00FF000000000000ull);
            idx = mulhi(x0, n);
            if (threat == 0) {
                src = (uint32_t)__ffsll((long long)(W | F | E)) - 1u;
                dir = ((W >> src) & 1ull) ? 0u : (((F >> src) & 1ull) ? 1u : 2u);
                have = true;
            } else if ((threat & (threat - 1)) == 0) {
                // captures of cell t in list order (source ascending): Black E from t+7 then W from t+9,
                // White E from t-9 then W from t-7
                const int t = __ffsll((long long)threat) - 1;
                const int s_e = is_white ? t - 9 : t + 7, s_w = is_white ? t - 7 : t + 9;
                const bool e_ok = s_e >= 0 && s_e < 64 && ((E >> (s_e & 63)) & 1ull);
                const bool w_ok = s_w >= 0 && s_w < 64 && ((W >> (s_w & 63)) & 1ull);
                if (e_ok || w_ok) {
                    src = (uint32_t)(e_ok ? s_e : s_w);
                    dir = e_ok ? 2u : 0u;
                    have = true;
                }
            }
        }
    }
    if (policy_is_lgr(POLICY)) {
        // the recorded reply, if it is legal now (simulate.rs:1022-1026); else the uniform pick.  LGRF-2 tries its 2-ply
        // reply first (simulate.rs:1124-1131)
        idx = mulhi(x0, n);
#pragma unroll
        for (int c = (policy_is_lgr2(POLICY) ? 0 : 1); c < 2; ++c) {
            const uint32_t rep = c == 0 ? lgr_reply2 : lgr_reply;
            if (!have && rep) {
                const uint32_t s0 = (rep - 1u) >> 8, d0 = (rep - 1u) & 0xffu;
                const uint32_t dr = d0 - s0 - (is_white ? 7u : (uint32_t)-9);
                if (s0 < 64u && dr < 3u) {
                    const uint64_t mk = dr == 0 ? W : (dr == 1 ? F : E);
                    if ((mk >> s0) & 1ull) {
                        src = s0;
                        dir = dr;
                        have = true;
                    }
                }
            }
        }
    }
    if (MAST_INNER && !have && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
        const uint32_t pl = is_white ? 1u : 0u;
        uint64_t cand[3] = {W, F, E};
        if (POLICY == MR_EPS_NST) mast_argmax<3>(cand, nst_planes, nst_nbits);
        else mast_argmax<3>(cand, p.mast_planes[pl], p.mast_nbits[pl]);
        const uint64_t cs = cand[0] ^ cand[1] ^ cand[2];
        const uint64_t cc = (cand[0] & cand[1]) | (cand[0] & cand[2]) | (cand[1] & cand[2]);
        const uint32_t m = __popcll(cs) + 2u * __popcll(cc);
        if (m == 1) {
            src = (uint32_t)__ffsll((long long)(cand[0] | cand[1] | cand[2])) - 1u;
            dir = cand[0] ? 0u : (cand[1] ? 1u : 2u);
            have = true;
        } else {
            const uint32_t r = mulhi(x0, 16u * n);
            const uint32_t i0 = r >> 4;
            if (m == n) {
                idx = i0;
                idx_set = true;
            } else if (m <= 6) {
                // few maximal actions: the one with the smallest visiting rank (half_tie)
                const uint32_t pk = r & 15u;
                const uint32_t inv = MR_STRIDE_INV[n][pk], rcp = MR_RCP32[n];
                const uint64_t cells = cand[0] | cand[1] | cand[2];
                uint32_t best = 0xffffffffu;
                if (lo32(cells)) best = half_tie<false>(best, cand, W, F, S, C, 0u, n, i0, inv, rcp);
                if (hi32(cells)) best = half_tie<true>(best, cand, W, F, S, C, __popc(lo32(S)) + 2u * __popc(lo32(C)), n, i0, inv, rcp);
                src = (best >> 2) & 63u;
                dir = best & 3u;
                have = true;
            } else {
                // many maximal actions: walk the visiting order, the first maximal action is close
                const uint32_t stride = MR_STRIDE[n][r & 15u];
                uint32_t i = i0;
                for (uint32_t j = 0; j < n; ++j) {
                    uint32_t d;
                    const uint32_t s = select_move(W, F, E, S, C, i, d);
                    const uint64_t cd = d == 0 ? cand[0] : (d == 1 ? cand[1] : cand[2]);
                    if ((cd >> s) & 1ull) {
                        src = s;
                        dir = d;
                        break;
                    }
                    i += stride;
                    if (i >= n) i -= n;
                }
                have = true;
            }
        }
    }
    if (!have && !idx_set) idx = mulhi(x0, n);
    if (!have) src = select_move(W, F, E, S, C, idx, dir);
    const uint32_t dst = (uint32_t)((int)src + (is_white ? 7 : -9) + (int)dir);
    action = (src << 8) | dst;
    slot = src * 3u + dir;
    const bool won = apply_move(is_white, src, dst);
    if (won) return ST_MOVED | (is_white ? ST_WIN_P1 : ST_WIN_P0);
    return ST_MOVED;
}

template <int POLICY, int PARITY>
__device__ int Knightthrough::step(const Shared &sh, uint32_t x0, uint32_t x1, const KernelParams &p,
                                   uint32_t &action, uint32_t &slot) {
    const bool is_white = ((turn_u ^ (PARITY == 2 ? rt_parity : (uint32_t)PARITY)) & 1u) != 0;
    const uint64_t own = is_white ? white : black;
    uint64_t M[8];
    masks(own, M);
    // per-cell move count as a bit-sliced 4-bit number (carry-save adder tree over the 8 masks)
    const uint64_t s01 = M[0] ^ M[1] ^ M[2], c01 = (M[0] & M[1]) | (M[0] & M[2]) | (M[1] & M[2]);
    const uint64_t s23 = M[3] ^ M[4] ^ M[5], c23 = (M[3] & M[4]) | (M[3] & M[5]) | (M[4] & M[5]);
    const uint64_t s45 = M[6] ^ M[7], c45 = M[6] & M[7];
    const uint64_t B0 = s01 ^ s23 ^ s45;                                               // weight 1
    const uint64_t k0 = (s01 & s23) | (s01 & s45) | (s23 & s45);                       // carry into weight 2
    const uint64_t t1 = c01 ^ c23 ^ c45, u1 = (c01 & c23) | (c01 & c45) | (c23 & c45); // weight 2 sum, weight 4 carry
    const uint64_t B1 = t1 ^ k0, k1 = t1 & k0;                                         // weight 2, carry into 4
    const uint64_t B2 = u1 ^ k1, B3 = u1 & k1;                                         // weight 4, weight 8
    const uint32_t n = __popcll(B0) + 2u * __popcll(B1) + 4u * __popcll(B2) + 8u * __popcll(B3);
    if (n == 0) return ST_NO_ACTIONS;

    auto select_move = [&](uint32_t idx, uint32_t &kind) -> uint32_t {
        const uint32_t cl = __popc(lo32(B0)) + 2u * __popc(lo32(B1)) + 4u * __popc(lo32(B2)) + 8u * __popc(lo32(B3));
        const bool in_hi = idx >= cl;
        idx = in_hi ? idx - cl : idx;
        const uint32_t b0 = in_hi ? hi32(B0) : lo32(B0), b1 = in_hi ? hi32(B1) : lo32(B1);
        const uint32_t b2 = in_hi ? hi32(B2) : lo32(B2), b3 = in_hi ? hi32(B3) : lo32(B3);
        uint32_t pos = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const uint32_t m = ((1u << step) - 1u) << pos;
            const uint32_t cnt = __popc(b0 & m) + 2u * __popc(b1 & m) + 4u * __popc(b2 & m) + 8u * __popc(b3 & m);
            select_step(idx, pos, cnt, (uint32_t)step);
        }
        // residual: the idx-th set kind at this cell, kinds in ascending-dst order -- the cell's kinds as a byte, then a
        // three-step rank/select over it
        uint32_t kb = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) kb |= (((in_hi ? hi32(M[q]) : lo32(M[q])) >> pos) & 1u) << q;
        uint32_t kk = 0;
#pragma unroll
        for (int step = 4; step >= 1; step >>= 1) {
            const uint32_t cnt = __popc(kb & (((1u << step) - 1u) << kk));
            select_step(idx, kk, cnt, (uint32_t)step);
        }
        kind = kk;
        return pos + (in_hi ? 32u : 0u);
    };

    uint32_t idx = 0, src = 0, kind = 0;
    bool have = false, idx_set = false;
    constexpr bool DECISIVE = POLICY == MR_DECISIVE_WIN || POLICY == MR_DECISIVE_ANTI || POLICY == MR_DECISIVE_MAST;
    constexpr bool MAST_INNER = policy_mast_inner(POLICY);
    if (DECISIVE) {
        // first action in list order that lands on the far wall (simulate.rs:662-687); a jump of two rows wins
        // from the two rows next to it, a jump of one row from the row next to it
        const uint64_t r1 = is_white ? 0x00FF000000000000ull : 0x000000000000FF00ull; // one row from the goal
        const uint64_t r2 = is_white ? 0x0000FF0000000000ull : 0x0000000000FF0000ull; // two rows from the goal
        uint64_t win[8];
        uint64_t any_win = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            // kinds 0,1 jump -2 rows, 2,3 jump -1 row, 4,5 jump +1 row, 6,7 jump +2 rows
            const bool toward = is_white ? q >= 4 : q < 4;
            const bool two = q < 2 || q >= 6;
            win[q] = toward ? (M[q] & (two ? (r1 | r2) : r1)) : 0ull;
            any_win |= win[q];
        }
        if (any_win) {
            src = (uint32_t)__ffsll((long long)any_win) - 1u;
#pragma unroll
            for (int q = 7; q >= 0; --q)
                if ((win[q] >> src) & 1ull) kind = q;
            have = true;
        } else if (POLICY == MR_DECISIVE_ANTI) {
            // AntiDecisive pass 2 (simulate.rs:701-733).  An opposing knight within two rows of ITS goal always has
            // a legal jump onto the goal row, so the opponent can win at once iff such a knight survives our move:
            // none -> the first legal action; exactly one -> the first action capturing it; more -> the inner pick.
            const uint64_t opp = is_white ? black : white;
            const uint64_t threat = opp & (is_white ? 0x0000000000FFFF00ull : 0x00FFFF0000000000ull);
            idx = mulhi(x0, n);
            if (threat == 0) {
                uint64_t any = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) any |= M[q];
                src = (uint32_t)__ffsll((long long)any) - 1u;
#pragma unroll
                for (int q = 7; q >= 0; --q)
                    if ((M[q] >> src) & 1ull) kind = q;
                have = true;
            } else if ((threat & (threat - 1)) == 0) {
                const int t = __ffsll((long long)threat) - 1;
                // sources t - delta(kind); the largest delta gives the smallest source, i.e. the first in list order
#pragma unroll
                for (int q = 7; q >= 0; --q) {
                    const int s0 = t - delta((uint32_t)q);
                    if (!have && s0 >= 0 && s0 < 64 && ((M[q] >> (s0 & 63)) & 1ull)) {
                        src = (uint32_t)s0;
                        kind = q;
                        have = true;
                    }
                }
            }
        }
    }
    if (policy_is_lgr(POLICY)) {
        idx = mulhi(x0, n);
#pragma unroll
        for (int c = (policy_is_lgr2(POLICY) ? 0 : 1); c < 2; ++c) {
            const uint32_t rep = c == 0 ? lgr_reply2 : lgr_reply;
            if (!have && rep) {
                const uint32_t s0 = (rep - 1u) >> 8, d0 = (rep - 1u) & 0xffu;
                if (s0 < 64u) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if ((int)d0 - (int)s0 == delta((uint32_t)q) && ((M[q] >> s0) & 1ull)) {
                            src = s0;
                            kind = q;
                            have = true;
                        }
                }
            }
        }
    }
    if (MAST_INNER && !have && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
        const uint32_t pl = is_white ? 1u : 0u;
        uint64_t cand[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) cand[q] = M[q];
        if (POLICY == MR_EPS_NST) mast_argmax<8>(cand, nst_planes, nst_nbits);
        else mast_argmax<8>(cand, p.mast_planes[pl], p.mast_nbits[pl]);
        uint64_t any = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) any |= cand[q];
        if ((any & (any - 1ull)) == 0ull) {
            // Every maximal action starts from ONE cell -- the unique maximum, or (a quarter of the plies) the winning jumps
            // of one knight tied at score 1: the cell's candidate kinds as a byte, no popcounts over the eight boards
            src = (uint32_t)__ffsll((long long)any) - 1u;
            const bool hi = src >= 32u;
            const uint32_t b = src & 31u;
            uint32_t kb = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) kb |= (((hi ? hi32(cand[q]) : lo32(cand[q])) >> b) & 1u) << q;
            if ((kb & (kb - 1u)) == 0u) {
                kind = (uint32_t)__ffs((int)kb) - 1u;
                have = true;
            } else {
                const uint32_t r = mulhi(x0, 16u * n);
                const uint32_t i0 = r >> 4;
                if ((uint32_t)__popc(kb) == n) { // the cell's jumps are all there is
                    idx = i0;
                    idx_set = true;
                } else {
                    // random_best among the cell's tied kinds: visiting rank of the cell's first action, + inv (mod n) per
                    // further legal kind (as in the general loop below)
                    const uint32_t pk = r & 15u;
                    const uint32_t inv = MR_STRIDE_INV[n][pk], rcp = MR_RCP32[n];
                    uint32_t lb = 0;
#pragma unroll
                    for (int q = 0; q < 7; ++q) lb |= (((hi ? hi32(M[q]) : lo32(M[q])) >> b) & 1u) << q;
                    const uint32_t below = (1u << b) - 1u;
                    const uint32_t h0 = hi ? hi32(B0) : lo32(B0), h1 = hi ? hi32(B1) : lo32(B1);
                    const uint32_t h2 = hi ? hi32(B2) : lo32(B2), h3 = hi ? hi32(B3) : lo32(B3);
                    uint32_t i = (uint32_t)(__popc(h0 & below) + 2 * __popc(h1 & below) + 4 * __popc(h2 & below) + 8 * __popc(h3 & below));
                    if (hi) i += __popc(lo32(B0)) + 2u * __popc(lo32(B1)) + 4u * __popc(lo32(B2)) + 8u * __popc(lo32(B3));
                    uint32_t j = visiting_rank(i, i0, n, inv, rcp), best = 0xffffffffu;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if ((kb >> q) & 1u) best = min(best, (j << 3) | (uint32_t)q);
                        if (q < 7 && ((lb >> q) & 1u)) {
                            j += inv;
                            j = min(j, j - n); // j - n wraps to a huge value while j < n
                        }
                    }
                    kind = best & 7u;
                    have = true;
                }
            }
        } else {
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) m += __popcll(cand[q]);
            const uint32_t r = mulhi(x0, 16u * n);
            const uint32_t i0 = r >> 4;
            if (m == n) {
                idx = i0;
                idx_set = true;
            } else if (m <= 6) {
                const uint32_t pk = r & 15u;
                const uint32_t inv = MR_STRIDE_INV[n][pk], rcp = MR_RCP32[n];
                // as for Breakthrough (half_tie): one 32-bit half of the board at a time -- the tied actions are as a
                // rule the winning jumps of one knight -- rank of the cell's first kind, + inv (mod n) per further legal
                // kind, minimum over packed keys (rank << 9 | cell << 3 | kind)
                uint32_t best = 0xffffffffu;
                auto half = [&](auto HI_T, uint32_t before) {
                    constexpr bool HI = decltype(HI_T)::value;
                    const uint32_t b0 = HI ? hi32(B0) : lo32(B0), b1 = HI ? hi32(B1) : lo32(B1);
                    const uint32_t b2 = HI ? hi32(B2) : lo32(B2), b3 = HI ? hi32(B3) : lo32(B3);
                    for (uint32_t cells = HI ? hi32(any) : lo32(any); cells; cells &= cells - 1u) {
                        const uint32_t b = (uint32_t)__ffs((int)cells) - 1u;
                        const uint32_t below = (1u << b) - 1u;
                        const uint32_t i = before + (uint32_t)(__popc(b0 & below) + 2 * __popc(b1 & below) + 4 * __popc(b2 & below) +
                                                               8 * __popc(b3 & below));
                        const uint32_t key = ((HI ? 32u : 0u) + b) << 3;
                        uint32_t j = visiting_rank(i, i0, n, inv, rcp);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (((HI ? hi32(cand[q]) : lo32(cand[q])) >> b) & 1u) best = min(best, (j << 9) | key | (uint32_t)q);
                            if (q < 7 && (((HI ? hi32(M[q]) : lo32(M[q])) >> b) & 1u)) {
                                j += inv;
                                j = min(j, j - n); // j - n wraps to a huge value while j < n
                            }
                        }
                    }
                };
                if (lo32(any)) half(std::false_type{}, 0u);
                if (hi32(any))
                    half(std::true_type{}, __popc(lo32(B0)) + 2u * __popc(lo32(B1)) + 4u * __popc(lo32(B2)) + 8u * __popc(lo32(B3)));
                src = (best >> 3) & 63u;
                kind = best & 7u;
                have = true;
            } else {
                const uint32_t stride = MR_STRIDE[n][r & 15u];
                uint32_t i = i0;
                for (uint32_t j = 0; j < n; ++j) {
                    uint32_t kq;
                    const uint32_t s = select_move(i, kq);
                    uint64_t ck = cand[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) ck = kq == (uint32_t)q ? cand[q] : ck;
                    if ((ck >> s) & 1ull) {
                        src = s;
                        kind = kq;
                        break;
                    }
                    i += stride;
                    if (i >= n) i -= n;
                }
                have = true;
            }
        }
    }
    if (!have && !idx_set) idx = mulhi(x0, n);
    if (!have) src = select_move(idx, kind);
    const uint32_t dst = (uint32_t)((int)src + delta(kind));
    action = (src << 8) | dst;
    slot = (src << 3) | kind;
    const bool won = apply_move(is_white, src, dst);
    if (won) return ST_MOVED | (is_white ? ST_WIN_P1 : ST_WIN_P0);
    return ST_MOVED;
}

// ==========================================================================================
// Othello 8x8 -- games/othello/src/lib.rs
//   movegen (:132-180) and flips (:188-234) by directional fills; actions = legal squares ascending, or
//   the single PASS (=64) when there are none (:484-499); apply (:502-541); terminal = 64 discs, or
//   last_pass and the mover has no move (:565-579); winner by disc count (:585-596).
// ==========================================================================================
struct Othello {
    static constexpr int NA = 65;
    static constexpr int NSLOTS = 65;
    static constexpr int KINDS = 1;
    static constexpr bool MASK_AMAF = false;
    __device__ static uint32_t slot_action_id(uint32_t pl, uint32_t slot) { return slot; }
    __device__ static uint32_t code_to_slot(uint32_t pl, uint32_t code) { return code < 65u ? code : 0u; }
    // The per-ply code (8 fills for moves + 8 for flips) is large: the four plies of a window run as a LOOP,
    // with the ply's parity a run-time value, so the kernel fits the instruction cache.
    static constexpr bool ROLLED = true;
    __host__ __device__ static constexpr bool rolled(int) { return ROLLED; }
    static constexpr int MIN_BLOCKS = 1; // a register cap (6 blocks/SM, 80 regs) measured no faster on B200
    static constexpr int min_blocks(int policy) { return MIN_BLOCKS; }

    struct Shared {
        uint64_t black, white;
        uint32_t turn, last_pass;
        uint64_t ray[3][64]; // squares beyond x towards +8 / +9 / +7 up to the board edge (flips_ray)
    };
    __device__ static void init_shared(Shared &s, uint32_t lane) { // once per warp, before the first task
        for (uint32_t e = lane; e < 192u; e += 32u) {
            const int k = (int)(e >> 6), r0 = (int)((e & 63u) >> 3), c0 = (int)(e & 7u), dc = k == 0 ? 0 : (k == 1 ? 1 : -1);
            uint64_t m = 0;
            for (int i = 1; i < 8; ++i) {
                const int r = r0 + i, c = c0 + dc * i;
                if (r <= 7 && c >= 0 && c <= 7) m |= 1ull << (r * 8 + c);
            }
            s.ray[k][e & 63u] = m;
        }
    }
    uint64_t black, white;
    uint32_t last_pass;
    uint32_t rt_parity; // parity of the current ply inside its window (PARITY == 2 call sites)
    uint32_t turn_u; // the leaf's mover as a warp-uniform value (see Connect4::turn_u)
    uint32_t lgr_reply; // MR_EPS_LGR: 1 + square of the recorded reply (65 = PASS), 0 = none / exploring
    uint32_t lgr_reply2; // MR_EPS_LGR2: the 2-ply table's reply, tried before lgr_reply
    const uint64_t (*nst_planes)[MAST_BITS]; // MR_EPS_NST: this ply's context rank planes, set by the kernel
    uint32_t nst_nbits;
    const uint16_t *nst_ranks;

    // (measured: writing the high word of right shifts as a multiply-high to move it to the FMA pipe is SLOWER,
    // 13.9 ms vs 13.2 ms per 2^24 playouts: IMAD.HI is half rate)
    // (also measured: a 64-bit left shift as IMAD.WIDE + IMAD, no ALU instruction at all, for one to four of the eight
    // directions: 12.5 - 13.0 ms vs 12.3 ms -- the half-rate wide multiply costs more than the funnel shift it replaces)
    template <int S, bool L>
    __device__ static uint64_t sh(uint64_t v) { return L ? (v << S) : (v >> S); }

    // discs of `om` reachable from `g` by repeated steps of S (left when L): parallel prefix, runs of <= 6 discs
    // (the same Kogge-Stone scheme as othello/lib.rs:101-124, with the wall mask folded into `om`)
    template <int S, bool L>
    __device__ static uint64_t ray(uint64_t g, uint64_t om) {
        uint64_t f = om & sh<S, L>(g);
        f |= om & sh<S, L>(f);
        const uint64_t pre = om & sh<S, L>(om);
        f |= pre & sh<2 * S, L>(f);
        f |= pre & sh<2 * S, L>(f);
        return f;
    }

    // othello/lib.rs:132-180.  An opponent disc on column 0 or 7 can never be outflanked along a row or a
    // diagonal, so those axes flood through `O & INNER` and need no wrap masks on the final step.
    static constexpr uint64_t INNER = 0x7E7E7E7E7E7E7E7Eull;
    // The four "toward lower bits" directions are run as LEFT-shift rays on the bit-reversed boards (bit i <-> 63-i
    // maps a right shift by s onto a left shift by s, and keeps the INNER column mask).  Left shifts put their low
    // word on the FMA pipe (IMAD.SHL); right shifts are two ALU funnel shifts, and the ALU pipe is 95% busy here.
    // The +1 direction (next cell of the same row) needs no parallel prefix: adding the cells right after the own discs
    // to the inner opponent discs ripples a carry through every contiguous run that starts there and drops it on the
    // first cell after the run (`mO` holds columns 1..6 only, so a carry never leaves its row).  6 instructions instead
    // of a 22-instruction ray; with the bit-reversed boards this covers both row directions.
    __device__ static uint64_t moves_carry1(uint64_t P, uint64_t mO) {
        const uint64_t x = P << 1;
        return (mO + x) & ~mO & ~x; // cells where a carry landed (an own disc's neighbour without a run is `x` itself)
    }
    __device__ static uint64_t flips_carry1(uint64_t mv, uint64_t P, uint64_t mO) {
        const uint64_t x = mv << 1, a = mO + x;
        return ((a & ~mO & ~x) & P) ? (mO & ~a) : 0ull; // the run the carry cleared, if it landed on an own disc
    }
    __device__ static uint64_t brev64(uint64_t v) { return mk64(__brev(hi32(v)), __brev(lo32(v))); }
    __device__ static uint64_t moves_left(uint64_t P, uint64_t O) { // the four left-shift directions
        const uint64_t mO = O & INNER;
        return sh<8, true>(ray<8, true>(P, O)) | moves_carry1(P, mO) | sh<9, true>(ray<9, true>(P, mO)) |
               sh<7, true>(ray<7, true>(P, mO));
    }
    __device__ static uint64_t legal_moves(uint64_t P, uint64_t O) {
        const uint64_t m = moves_left(P, O) | brev64(moves_left(brev64(P), brev64(O)));
        return m & ~(P | O);
    }
    // othello/lib.rs:188-234: a direction flips its run iff the cell after the run holds an own disc
    template <int S>
    __device__ static uint64_t flips_dir(uint64_t mv, uint64_t P, uint64_t om) {
        const uint64_t f = ray<S, true>(mv, om);
        return (sh<S, true>(f) & P) ? f : 0ull;
    }
    __device__ static uint64_t flips_left(uint64_t mv, uint64_t P, uint64_t O) {
        const uint64_t mO = O & INNER;
        return flips_dir<8>(mv, P, O) | flips_carry1(mv, P, mO) | flips_dir<9>(mv, P, mO) | flips_dir<7>(mv, P, mO);
    }
    __device__ static uint64_t flips(uint64_t mv, uint64_t P, uint64_t O) {
        return flips_left(mv, P, O) | brev64(flips_left(brev64(mv), brev64(P), brev64(O)));
    }
    // The flips of ONE known square need no prefix network in the other directions either: with every bit off the ray set
    // (`O | ~ray`), adding the first ray cell ripples a carry along the ray through the opponent's run -- across the gaps
    // between ray cells, which are all ones -- and drops it on the first ray cell that is not an opponent disc.  The run
    // flips iff that cell holds an own disc; a run that reaches the edge carries out of the word.  (The classic
    // "outflank by carry" of Othello engines; the reference's Kogge-Stone fills are othello/lib.rs:188-234.)
    template <int S, int K>
    __device__ static uint64_t flips_ray(const Shared &s, uint32_t sq, uint64_t mv, uint64_t P, uint64_t O) {
        const uint64_t m = s.ray[K][sq];
        const uint64_t u = O | ~m, t = u + (mv << S);
        return (t & ~u & m & P) ? (u & ~t & m) : 0ull;
    }
    __device__ static uint64_t flips_left_sq(const Shared &s, uint32_t sq, uint64_t mv, uint64_t P, uint64_t O) {
        return flips_ray<8, 0>(s, sq, mv, P, O) | flips_carry1(mv, P, O & INNER) | flips_ray<9, 1>(s, sq, mv, P, O) |
               flips_ray<7, 2>(s, sq, mv, P, O);
    }
    __device__ static uint64_t flips_sq(const Shared &s, uint32_t sq, uint64_t P, uint64_t O) {
        const uint64_t mv = bit64(sq);
        return flips_left_sq(s, sq, mv, P, O) | brev64(flips_left_sq(s, 63u - sq, brev64(mv), brev64(P), brev64(O)));
    }
    // position of the idx-th set bit of a 64-bit mask
    __device__ static uint32_t select64(uint64_t w, uint32_t idx) {
        const uint32_t nlo = __popc(lo32(w));
        const bool in_hi = idx >= nlo;
        return select32(in_hi ? hi32(w) : lo32(w), in_hi ? idx - nlo : idx) + (in_hi ? 32u : 0u);
    }
    __device__ static int count_outcome(uint64_t b, uint64_t w) {
        const int nb = __popcll(b), nw = __popcll(w);
        return nb > nw ? MR_WINNER_P0 : (nb < nw ? MR_WINNER_P1 : MR_DRAW);
    }

    __device__ static int prepare(Shared &s, const mr_leaf &leaf, uint32_t lane) {
        if (lane == 0) {
            s.black = leaf.board[0];
            s.white = leaf.board[1];
            s.turn = leaf.turn;
            s.last_pass = leaf.flag;
        }
        const uint64_t b = leaf.board[0], w = leaf.board[1];
        bool term = __popcll(b | w) == 64;
        if (!term && leaf.flag) term = legal_moves(leaf.turn ? w : b, leaf.turn ? b : w) == 0;
        return term ? count_outcome(b, w) : MR_NOT_TERMINAL;
    }
    __device__ void reset(const Shared &s) {
        black = s.black;
        white = s.white;
        last_pass = s.last_pass;
    }

    template <int POLICY, int PARITY>
    __device__ int step(const Shared &s, uint32_t x0, uint32_t x1, const KernelParams &p, uint32_t &action, uint32_t &slot) {
        const bool is_white = ((turn_u ^ (PARITY == 2 ? rt_parity : (uint32_t)PARITY)) & 1u) != 0;
        const uint64_t P = is_white ? white : black, O = is_white ? black : white;
        const uint64_t moves = legal_moves(P, O);
        if (moves == 0) {
            // the only action is PASS.  (A double pass is terminal and was caught after the first pass.)
            action = 64;
            slot = 64;
            last_pass = 1;
            // terminal next iff the opponent has no move either (:565-579)
            if (legal_moves(O, P) == 0) return ST_MOVED | terminal_code(black, white);
            return ST_MOVED;
        }
        // DecisiveMove<Win> (simulate.rs:662-687) can only find a terminal child when a single action
        // exists (the 64th disc, or a PASS), so it never changes the pick: see DESIGN.md "Decisive move".
        const uint32_t n = __popcll(moves);
        uint32_t sq = 0;
        uint64_t mv = 0, fl = 0;
        bool have = false;
        if (POLICY == MR_DECISIVE_ANTI) {
            // DecisiveMove<AntiDecisive> (simulate.rs:689-735).  Pass 1 can only fire on the single 64th-disc action
            // (the pick either way).  Pass 2: the first square whose child is not terminal and after which no
            // opponent reply ends the game with a winner -- a reply ends the game only as the 64th disc, or as a
            // PASS answered by a position where we cannot move either.
            const uint32_t discs = __popcll(P | O);
            for (uint64_t m = moves; m && !have; m &= m - 1) {
                const uint64_t c_mv = m & (0 - m);
                const uint64_t c_fl = flips(c_mv, P, O);
                const uint64_t cP = P | c_mv | c_fl, cO = O & ~c_fl; // child: cO to move
                if (discs == 63) break; // the child is terminal (full board): never preferred
                const uint64_t replies = legal_moves(cO, cP);
                bool opponent_can_win = false;
                if (replies == 0) { // the only reply is PASS
                    if (legal_moves(cP, cO) == 0) opponent_can_win = __popcll(cP) != __popcll(cO);
                } else if (discs == 62) { // a reply places the 64th disc
                    for (uint64_t r = replies; r; r &= r - 1) {
                        const uint64_t r_mv = r & (0 - r);
                        const uint64_t r_fl = flips(r_mv, cO, cP);
                        if (__popcll(cO | r_mv | r_fl) != __popcll(cP & ~r_fl)) opponent_can_win = true;
                    }
                }
                if (!opponent_can_win) {
                    mv = c_mv;
                    fl = c_fl;
                    sq = (uint32_t)__ffsll((long long)c_mv) - 1u;
                    have = true;
                }
            }
        }
        if (policy_is_lgr(POLICY)) {
            // the recorded reply square, if legal now (a recorded PASS is legal exactly when it is the only action,
            // handled above); LGRF-2 tries its 2-ply reply first
            uint32_t rep = (lgr_reply && lgr_reply <= 64u && ((moves >> (lgr_reply - 1u)) & 1ull)) ? lgr_reply : 0u;
            if (policy_is_lgr2(POLICY) && lgr_reply2 && lgr_reply2 <= 64u && ((moves >> (lgr_reply2 - 1u)) & 1ull)) rep = lgr_reply2;
            if (rep) {
                sq = rep - 1u;
                mv = bit64(sq);
                fl = flips_sq(s, sq, P, O);
                have = true;
            }
        }
        if (policy_mast_inner(POLICY) && !have && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
            // Mast::select_move (simulate.rs:824-848): bit-sliced arg-max of the rank planes over the legal squares,
            // then random_best's visiting order among the maximal ones (as for Breakthrough, with one move kind)
            const uint32_t pl = is_white ? 1u : 0u;
            uint64_t cand[1] = {moves};
            if (POLICY == MR_EPS_NST) mast_argmax<1>(cand, nst_planes, nst_nbits);
            else mast_argmax<1>(cand, p.mast_planes[pl], p.mast_nbits[pl]);
            const uint32_t m = __popcll(cand[0]);
            const uint32_t r = mulhi(x0, 16u * n);
            const uint32_t i0 = r >> 4;
            if (m == 1) {
                sq = (uint32_t)__ffsll((long long)cand[0]) - 1u;
            } else if (m == n) { // every legal square is maximal (e.g. an empty table): list position i0
                sq = select64(moves, i0);
            } else {
                const uint32_t inv = MR_STRIDE_INV[n][r & 15u], rcp = MR_RCP32[n];
                uint32_t best_j = 0xffffffffu;
                for (uint64_t cells = cand[0]; cells; cells &= cells - 1) {
                    const uint32_t s0 = (uint32_t)__ffsll((long long)cells) - 1u;
                    const uint32_t j = visiting_rank((uint32_t)__popcll(moves & ((1ull << s0) - 1ull)), i0, n, inv, rcp);
                    if (j < best_j) {
                        best_j = j;
                        sq = s0;
                    }
                }
            }
            mv = bit64(sq);
            fl = flips_sq(s, sq, P, O);
            have = true;
        }
        if (!have) {
            sq = select64(moves, mulhi(x0, n));
            mv = bit64(sq);
            fl = flips_sq(s, sq, P, O);
        }
        const uint64_t nP = P | mv | fl, nO = O & ~fl;
        black = is_white ? nO : nP;
        white = is_white ? nP : nO;
        last_pass = 0;
        action = sq;
        slot = sq;
        if (__popcll(black | white) == 64) return ST_MOVED | terminal_code(black, white);
        return ST_MOVED;
    }
    __device__ static int terminal_code(uint64_t b, uint64_t w) {
        const int o = count_outcome(b, w);
        return o == MR_WINNER_P0 ? ST_WIN_P0 : (o == MR_WINNER_P1 ? ST_WIN_P1 : ST_DRAW);
    }
    __device__ void store_final(const Shared &s, mr_leaf *out, uint32_t plies, int end_code) const {
        out->board[0] = black;
        out->board[1] = white;
        out->board[2] = out->board[3] = 0;
        out->turn = (uint8_t)((s.turn + plies) & 1);
        out->flag = (uint8_t)last_pass;
        out->prev_action_plus1 = 0;
        out->prev2_action_plus1 = 0;
        out->pad[0] = out->pad[1] = 0;
    }
};

// ==========================================================================================
// Hex 11x11 -- games/hex-gen/src/lib.rs (Position<11,2>)
//   legal = empty cells ascending (:115-120); apply sets the cell, toggles the turn (:122-125); only the
//   last mover is checked (:130-144): P0 connects north (row 10) <-> south (row 0), P1 west (col 0) <-> east
//   (col 10) over the six hex neighbours N,S,E,W,NE(+12),SW(-12) (bitboard/src/board.rs:509-531).
//   device: boards as 4 x u32; this struct only holds the board primitives (masks, six-neighbourhood, flood).
//   The playout itself is hex_kernel.cuh's fill-then-search, which replaces the per-ply flood6 fix-point by
//   one winner test at the end plus a binary search for the first connecting move -- same stop ply.
// ==========================================================================================
struct Hex11 {
    static constexpr int NA = 121;
    static constexpr int NSLOTS = 121;
    static constexpr int KINDS = 2; // the two 64-bit halves of the board
    // a Hex playout plays every cell at most once, so its (action, player) occurrences are exactly
    // final stones minus leaf stones: AMAF is accumulated from board masks, no move trace needed
    static constexpr bool MASK_AMAF = true;
    __device__ static uint32_t slot_action_id(uint32_t pl, uint32_t slot) { return slot; }
    struct B128 {
        uint32_t w[4];
    };
    __device__ static B128 zero() { return B128{{0, 0, 0, 0}}; }
    __device__ static B128 band(B128 a, B128 b) { return B128{{a.w[0] & b.w[0], a.w[1] & b.w[1], a.w[2] & b.w[2], a.w[3] & b.w[3]}}; }
    __device__ static B128 bor(B128 a, B128 b) { return B128{{a.w[0] | b.w[0], a.w[1] | b.w[1], a.w[2] | b.w[2], a.w[3] | b.w[3]}}; }
    __device__ static B128 bandn(B128 a, B128 b) { return B128{{a.w[0] & ~b.w[0], a.w[1] & ~b.w[1], a.w[2] & ~b.w[2], a.w[3] & ~b.w[3]}}; }
    __device__ static bool any(B128 a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) != 0; }
    template <int S>
    __device__ static B128 shl(B128 a) { // S < 32
        return B128{{a.w[0] << S, __funnelshift_l(a.w[0], a.w[1], S), __funnelshift_l(a.w[1], a.w[2], S),
                     __funnelshift_l(a.w[2], a.w[3], S)}};
    }
    template <int S>
    __device__ static B128 shr(B128 a) {
        return B128{{__funnelshift_r(a.w[0], a.w[1], S), __funnelshift_r(a.w[1], a.w[2], S),
                     __funnelshift_r(a.w[2], a.w[3], S), a.w[3] >> S}};
    }
    // column / row masks of the 121-cell board (idx = row*11 + col), folded at compile time
    __host__ __device__ static constexpr B128 col_mask(int col) {
        B128 m{{0, 0, 0, 0}};
        for (int r = 0; r < 11; ++r) {
            const int i = r * 11 + col;
            m.w[i >> 5] |= 1u << (i & 31);
        }
        return m;
    }
    __host__ __device__ static constexpr B128 row_mask(int row) {
        B128 m{{0, 0, 0, 0}};
        for (int c = 0; c < 11; ++c) {
            const int i = row * 11 + c;
            m.w[i >> 5] |= 1u << (i & 31);
        }
        return m;
    }
    __host__ __device__ static constexpr B128 inv(B128 c) { return B128{{~c.w[0], ~c.w[1], ~c.w[2], ~c.w[3]}}; }
    // six-neighbourhood of a set, restricted to the board: N(+11) S(-11) E(+1) W(-1) NE(+12) SW(-12)
    __device__ static B128 neighbours(B128 f, const B128 &NOT_COL0, const B128 &NOT_COL10, const B128 &BOARDM) {
        const B128 fe = band(f, NOT_COL10); // may step east
        const B128 fw = band(f, NOT_COL0);  // may step west
        B128 n = shl<11>(f);
        n = bor(n, shr<11>(f));
        n = bor(n, shl<1>(fe));
        n = bor(n, shr<1>(fw));
        n = bor(n, shl<12>(fe));
        n = bor(n, shr<12>(fw));
        return band(n, BOARDM);
    }
    // flood from `seed` through `own` to a fix-point (bitboard/src/board.rs:509-531)
    __device__ static B128 flood(B128 seed, B128 own, const B128 &NC0, const B128 &NC10, const B128 &BM) {
        B128 f = band(seed, own);
        for (;;) {
            const B128 g = bandn(band(neighbours(f, NC0, NC10, BM), own), f);
            if (!any(g)) break;
            f = bor(f, g);
        }
        return f;
    }
    __host__ __device__ static constexpr B128 board_mask() { return B128{{0xffffffffu, 0xffffffffu, 0xffffffffu, 0x01ffffffu}}; }
    // first / second edge of player p: P0 north(row 10) -> south(row 0); P1 west(col 0) -> east(col 10)
    __device__ static B128 first_edge(int p) {
        constexpr B128 n = row_mask(10), w = col_mask(0);
        return p == 0 ? n : w;
    }
    __device__ static B128 second_edge(int p) {
        constexpr B128 so = row_mask(0), e = col_mask(10);
        return p == 0 ? so : e;
    }
};

} // namespace mr
