/*
 * oracle/games.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the five `Game` impls on the playout path of
 * thomasmarsh/mcts.  Every function cites the reference lines it follows
 * (paths relative to the reference root).  Data flow is kept faithful to the
 * reference: action lists are materialised in the reference's order, the terminal
 * check is the default `terminal_status` = `is_terminal` then `winner`
 * (mcts/src/game.rs:202-211), walls are recomputed per call as the reference's
 * Const-dimension boards do (bitboard/src/board.rs:329-333).
 */
#include "oracle.h"
#include "bitboard.h"

#include <stdlib.h>
#include <string.h>

bb_t bb_wall_table[BB_WALL_MAX_DIM + 1][BB_WALL_MAX_DIM + 1][4];
__attribute__((constructor)) static void bb_wall_table_init(void) {
    for (int r = 1; r <= BB_WALL_MAX_DIM; ++r)
        for (int c = 1; c <= BB_WALL_MAX_DIM; ++c) {
            if (r * c > 128) continue;
            for (int d = 0; d < 4; ++d) bb_wall_table[r][c][d] = bb_wall_uncached(r, c, d);
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Connect4 -- games/connect4/src/lib.rs                                                       */
/* ------------------------------------------------------------------------------------------ */

/* lib.rs:195-207 four_in_a_row: m1 = b & step(b); m2 = m1 & step(step(m1)) */
static int c4_run(bb_t b, bb_t (*step)(bb_t)) {
    bb_t m1 = bb_and(b, step(b));
    bb_t m2 = bb_and(m1, step(step(m1)));
    return !bb_none_set(m2);
}
static int c4_four_in_a_row(bb_t b) {
    return c4_run(b, bb_shift_north) || c4_run(b, bb_shift_east) || c4_run(b, bb_shift_northeast) ||
           c4_run(b, bb_shift_northwest);
}

/* lib.rs:290-296 height: how far up the column stays occupied from row 0 */
static int c4_height(bb_t occupied, int col) {
    int row = 0;
    while (row < occupied.rows && bb_get_index(occupied, row * occupied.cols + col)) ++row;
    return row;
}

typedef struct {
    bb_t black, white;
    int turn, winner;
} c4_state;

static c4_state c4_load(const orc_state *s, int R, int C) {
    c4_state t;
    t.black = bb_empty(R, C);
    t.white = bb_empty(R, C);
    t.black.w[0] = s->board[0];
    t.white.w[0] = s->board[1];
    t.turn = s->turn;
    t.winner = s->flag;
    return t;
}
static void c4_store(orc_state *s, const c4_state *t) {
    s->board[0] = t->black.w[0];
    s->board[1] = t->white.w[0];
    s->turn = (uint8_t)t->turn;
    s->flag = (uint8_t)t->winner;
}

/* lib.rs:303-312 moves: columns ascending whose height < R; nothing once won */
static int c4_moves(const c4_state *t, uint16_t *out) {
    int n = 0;
    if (t->winner) return 0;
    bb_t occ = bb_or(t->black, t->white);
    for (int col = 0; col < occ.cols; ++col)
        if (c4_height(occ, col) < occ.rows) BB_OP(ORC_OP_LIST_STORE, 1), out[n++] = (uint16_t)col;
    return n;
}

/* lib.rs:314-339 apply: drop at row=height(col); win => winner=true and the turn is NOT advanced */
static void c4_apply(c4_state *t, int col) {
    bb_t occ = bb_or(t->black, t->white);
    int row = c4_height(occ, col);
    int index = row * occ.cols + col;
    bb_t player = t->turn == 0 ? t->black : t->white;
    player = bb_set_index(player, index);
    if (t->turn == 0)
        t->black = player;
    else
        t->white = player;
    if (c4_four_in_a_row(player))
        t->winner = 1;
    else
        t->turn ^= 1;
}

/* lib.rs:298-301 is_full, :372-374 is_terminal, :380-386 winner */
static int c4_is_terminal(const c4_state *t) {
    bb_t occ = bb_or(t->black, t->white);
    return t->winner || bb_count_ones(occ) == bb_len(occ);
}
static int c4_winner(const c4_state *t) { return t->winner ? t->turn : -1; }

/* ------------------------------------------------------------------------------------------ */
/* Breakthrough / Knightthrough -- games/breakthrough/src/lib.rs, games/knightthrough/src/lib.rs */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    bb_t black, white;
    int turn, winner;
} bt_state;

static bt_state bt_load(const orc_state *s) {
    bt_state t;
    t.black = bb_empty(8, 8);
    t.white = bb_empty(8, 8);
    t.black.w[0] = s->board[0];
    t.white.w[0] = s->board[1];
    t.turn = s->turn;
    t.winner = s->flag;
    return t;
}
static void bt_store(orc_state *s, const bt_state *t) {
    s->board[0] = t->black.w[0];
    s->board[1] = t->white.w[0];
    s->turn = (uint8_t)t->turn;
    s->flag = (uint8_t)t->winner;
}

/* breakthrough/lib.rs:64-82 (and knightthrough/lib.rs:59-77): black = north wall + the row south
 * of it, white = south wall + the row north of it, Black to move */
static void bt_default(orc_state *s) {
    bb_t n = bb_wall(8, 8, BB_NORTH);
    bb_t so = bb_wall(8, 8, BB_SOUTH);
    bb_t black = bb_or(n, bb_shift_south(n));
    bb_t white = bb_or(so, bb_shift_north(so));
    memset(s, 0, sizeof *s);
    s->board[0] = black.w[0];
    s->board[1] = white.w[0];
}

/* breakthrough/lib.rs:131-156 moves */
static int bt_moves(const bt_state *t, uint16_t *out) {
    int n = 0;
    if (t->winner) return 0;
    bb_t player = t->turn == 0 ? t->black : t->white;
    bb_t opponent = t->turn == 0 ? t->white : t->black;
    bb_t occupied = bb_or(player, opponent);
    bb_t it = player;
    int src;
    while ((src = bb_pop_lowest(&it)) >= 0) {
        bb_t from = bb_from_index(8, 8, src);
        bb_t forward = t->turn == 0 ? bb_shift_south(from) : bb_shift_north(from);
        bb_t w = bb_shift_west(bb_and(forward, bb_not(bb_wall_of(forward, BB_WEST))));
        bb_t e = bb_shift_east(bb_and(forward, bb_not(bb_wall_of(forward, BB_EAST))));
        bb_t available =
            bb_or(bb_or(bb_and(w, bb_not(player)), bb_and(e, bb_not(player))), bb_and(forward, bb_not(occupied)));
        int dst;
        while ((dst = bb_pop_lowest(&available)) >= 0) BB_OP(ORC_OP_LIST_STORE, 1), out[n++] = (uint16_t)((src << 8) | dst);
    }
    return n;
}

/* knightthrough/lib.rs:126-157 moves: all eight knight jumps that stay on the board, minus own
 * pieces, emitted in ascending dst order (bitboard iteration, not offset order) */
static int kt_moves(const bt_state *t, uint16_t *out) {
    static const int off[8][2] = {{2, 1}, {2, -1}, {-2, 1}, {-2, -1}, {1, 2}, {1, -2}, {-1, 2}, {-1, -2}};
    int n = 0;
    if (t->winner) return 0;
    bb_t player = t->turn == 0 ? t->black : t->white;
    bb_t it = player;
    int src;
    while ((src = bb_pop_lowest(&it)) >= 0) {
        bb_t knight = bb_empty(8, 8);
        int row = src / 8, col = src % 8;
        for (int k = 0; k < 8; ++k) {
            int r = row + off[k][0], c = col + off[k][1];
            if (r >= 0 && r < 8 && c >= 0 && c < 8) knight = bb_set_index(knight, r * 8 + c);
        }
        bb_t available = bb_and(knight, bb_not(player));
        int dst;
        while ((dst = bb_pop_lowest(&available)) >= 0) BB_OP(ORC_OP_LIST_STORE, 1), out[n++] = (uint16_t)((src << 8) | dst);
    }
    return n;
}

/* breakthrough/lib.rs:158-188, knightthrough/lib.rs:159-190 apply: move the piece, clear dst from the
 * opponent; touching the far wall wins and leaves the turn on the winner */
static void bt_apply(bt_state *t, uint16_t action) {
    bb_t src = bb_from_index(8, 8, action >> 8);
    bb_t dst = bb_from_index(8, 8, action & 0xff);
    bb_t player = t->turn == 0 ? t->black : t->white;
    bb_t other = t->turn == 0 ? t->white : t->black;
    player = bb_or(player, dst);
    player = bb_and(player, bb_not(src));
    bb_t opponent = bb_and(other, bb_not(dst));
    bb_t goal;
    if (t->turn == 0) {
        t->black = player;
        t->white = opponent;
        goal = bb_wall(8, 8, BB_SOUTH);
    } else {
        t->white = player;
        t->black = opponent;
        goal = bb_wall(8, 8, BB_NORTH);
    }
    if (bb_intersects(player, goal))
        t->winner = 1;
    else
        t->turn ^= 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Othello -- games/othello/src/lib.rs                                                         */
/* ------------------------------------------------------------------------------------------ */

#define OTH_INITIAL_BLACK ((((uint64_t)1) << 28) | (((uint64_t)1) << 35)) /* lib.rs:10 */
#define OTH_INITIAL_WHITE ((((uint64_t)1) << 27) | (((uint64_t)1) << 36)) /* lib.rs:11 */

static bb_t oth_bb(uint64_t bits) {
    bb_t b = bb_empty(8, 8);
    b.w[0] = bits;
    return b;
}
static bb_t oth_ones(void) { return bb_not(bb_empty(8, 8)); }

/* lib.rs:101-111 flood_left */
static bb_t oth_flood_left(bb_t p, bb_t o, int shift, bb_t mask) {
    bb_t gen = bb_and(bb_and(bb_shl(p, shift), mask), o);
    bb_t pro = bb_and(o, mask);
    gen = bb_or(gen, bb_and(bb_shl(gen, shift), pro));
    pro = bb_and(pro, bb_shl(pro, shift));
    gen = bb_or(gen, bb_and(bb_shl(gen, shift * 2), pro));
    pro = bb_and(pro, bb_shl(pro, shift * 2));
    gen = bb_or(gen, bb_and(bb_shl(gen, shift * 4), pro));
    return gen;
}
/* lib.rs:114-124 flood_right */
static bb_t oth_flood_right(bb_t p, bb_t o, int shift, bb_t mask) {
    bb_t gen = bb_and(bb_and(bb_shr(p, shift), mask), o);
    bb_t pro = bb_and(o, mask);
    gen = bb_or(gen, bb_and(bb_shr(gen, shift), pro));
    pro = bb_and(pro, bb_shr(pro, shift));
    gen = bb_or(gen, bb_and(bb_shr(gen, shift * 2), pro));
    pro = bb_and(pro, bb_shr(pro, shift * 2));
    gen = bb_or(gen, bb_and(bb_shr(gen, shift * 4), pro));
    return gen;
}

/* lib.rs:132-180 generate_moves */
static bb_t oth_generate_moves(bb_t player, bb_t opponent) {
    bb_t empty = bb_not(bb_or(player, opponent));
    bb_t legal = bb_empty(8, 8);
    bb_t not_east = bb_not(bb_wall(8, 8, BB_EAST));
    bb_t not_west = bb_not(bb_wall(8, 8, BB_WEST));
    bb_t ones = oth_ones();

    bb_t n = oth_flood_left(player, opponent, 8, ones);
    legal = bb_or(legal, bb_and(bb_shl(n, 8), empty));
    bb_t nw = oth_flood_left(player, opponent, 7, not_east);
    legal = bb_or(legal, bb_and(bb_and(bb_shl(nw, 7), not_east), empty));
    bb_t ne = oth_flood_left(player, opponent, 9, not_west);
    legal = bb_or(legal, bb_and(bb_and(bb_shl(ne, 9), not_west), empty));
    bb_t e = oth_flood_left(player, opponent, 1, not_west);
    legal = bb_or(legal, bb_and(bb_and(bb_shl(e, 1), not_west), empty));

    bb_t s = oth_flood_right(player, opponent, 8, ones);
    legal = bb_or(legal, bb_and(bb_shr(s, 8), empty));
    bb_t sw = oth_flood_right(player, opponent, 9, not_east);
    legal = bb_or(legal, bb_and(bb_and(bb_shr(sw, 9), not_east), empty));
    bb_t se = oth_flood_right(player, opponent, 7, not_west);
    legal = bb_or(legal, bb_and(bb_and(bb_shr(se, 7), not_west), empty));
    bb_t w = oth_flood_right(player, opponent, 1, not_east);
    legal = bb_or(legal, bb_and(bb_and(bb_shr(w, 1), not_east), empty));
    return legal;
}

/* lib.rs:188-234 get_flips */
static bb_t oth_get_flips(bb_t player, bb_t opponent, bb_t mv) {
    bb_t flips = bb_empty(8, 8);
    bb_t not_east = bb_not(bb_wall(8, 8, BB_EAST));
    bb_t not_west = bb_not(bb_wall(8, 8, BB_WEST));
    bb_t ones = oth_ones();

    bb_t n = oth_flood_left(mv, opponent, 8, ones);
    if (bb_intersects(bb_and(bb_shl(n, 8), player), player)) flips = bb_or(flips, n);
    bb_t ne = oth_flood_left(mv, opponent, 9, not_west);
    if (bb_intersects(bb_and(bb_and(bb_shl(ne, 9), not_west), player), player)) flips = bb_or(flips, ne);
    bb_t nw = oth_flood_left(mv, opponent, 7, not_east);
    if (bb_intersects(bb_and(bb_and(bb_shl(nw, 7), not_east), player), player)) flips = bb_or(flips, nw);
    bb_t e = oth_flood_left(mv, opponent, 1, not_west);
    if (bb_intersects(bb_and(bb_and(bb_shl(e, 1), not_west), player), player)) flips = bb_or(flips, e);

    bb_t s = oth_flood_right(mv, opponent, 8, ones);
    if (bb_intersects(bb_and(bb_shr(s, 8), player), player)) flips = bb_or(flips, s);
    bb_t sw = oth_flood_right(mv, opponent, 9, not_east);
    if (bb_intersects(bb_and(bb_and(bb_shr(sw, 9), not_east), player), player)) flips = bb_or(flips, sw);
    bb_t se = oth_flood_right(mv, opponent, 7, not_west);
    if (bb_intersects(bb_and(bb_and(bb_shr(se, 7), not_west), player), player)) flips = bb_or(flips, se);
    bb_t w = oth_flood_right(mv, opponent, 1, not_east);
    if (bb_intersects(bb_and(bb_and(bb_shr(w, 1), not_east), player), player)) flips = bb_or(flips, w);
    return flips;
}

uint64_t orc_othello_generate_moves(uint64_t player, uint64_t opponent) {
    return oth_generate_moves(oth_bb(player), oth_bb(opponent)).w[0];
}
uint64_t orc_othello_get_flips(uint64_t player, uint64_t opponent, uint64_t move_mask) {
    return oth_get_flips(oth_bb(player), oth_bb(opponent), oth_bb(move_mask)).w[0];
}

/* lib.rs:484-499 generate_actions: set bits ascending, or a single PASS */
static int oth_actions(const orc_state *s, uint16_t *out) {
    bb_t player = oth_bb(s->turn == 0 ? s->board[0] : s->board[1]);
    bb_t opponent = oth_bb(s->turn == 0 ? s->board[1] : s->board[0]);
    bb_t moves = oth_generate_moves(player, opponent);
    int n = 0;
    if (bb_count_ones(moves) == 0) {
        BB_OP(ORC_OP_LIST_STORE, 1), out[n++] = ORC_OTHELLO_PASS;
    } else {
        int idx;
        while ((idx = bb_pop_lowest(&moves)) >= 0) BB_OP(ORC_OP_LIST_STORE, 1), out[n++] = (uint16_t)idx;
    }
    return n;
}

/* lib.rs:502-541 apply (the eight incremental Zobrist hashes are tree-side state and are not
 * carried by the playout record) */
static void oth_apply(orc_state *s, uint16_t action) {
    if (action == ORC_OTHELLO_PASS) {
        s->flag = 1; /* last_pass */
        s->turn ^= 1;
        return;
    }
    bb_t piece = bb_from_index(8, 8, action);
    bb_t player = oth_bb(s->turn == 0 ? s->board[0] : s->board[1]);
    bb_t opponent = oth_bb(s->turn == 0 ? s->board[1] : s->board[0]);
    bb_t flips = oth_get_flips(player, opponent, piece);
    player = bb_or(bb_or(player, piece), flips);
    opponent = bb_and(opponent, bb_not(flips));
    if (s->turn == 0) {
        s->board[0] = player.w[0];
        s->board[1] = opponent.w[0];
    } else {
        s->board[1] = player.w[0];
        s->board[0] = opponent.w[0];
    }
    s->flag = 0;
    s->turn ^= 1;
}

/* lib.rs:565-579 is_terminal */
static int oth_is_terminal(const orc_state *s) {
    if (__builtin_popcountll(s->board[0] | s->board[1]) == 64) return 1;
    if (!s->flag) return 0;
    bb_t player = oth_bb(s->turn == 0 ? s->board[0] : s->board[1]);
    bb_t opponent = oth_bb(s->turn == 0 ? s->board[1] : s->board[0]);
    return bb_count_ones(oth_generate_moves(player, opponent)) == 0;
}

/* lib.rs:585-596 winner: re-runs is_terminal, then compares disc counts */
static int oth_winner(const orc_state *s) {
    if (!oth_is_terminal(s)) return -1;
    int black = __builtin_popcountll(s->board[0]);
    int white = __builtin_popcountll(s->board[1]);
    if (black > white) return 0;
    if (black < white) return 1;
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* Hex 11x11 -- games/hex-gen/src/lib.rs (Position<11,2>)                                       */
/* ------------------------------------------------------------------------------------------ */

#define HEX_N 11
static uint64_t g_hex_flood_iterations;
uint64_t orc_hex_flood_iterations(void) { return g_hex_flood_iterations; }
void orc_hex_flood_iterations_reset(void) { g_hex_flood_iterations = 0; }

static bb_t hex_bb(const uint64_t *w) {
    bb_t b = bb_empty(HEX_N, HEX_N);
    b.w[0] = w[0];
    b.w[1] = w[1];
    return b;
}

/* lib.rs:63-90 side_east / side_north / side_south / side_west */
static bb_t hex_side(int which) {
    bb_t b = bb_empty(HEX_N, HEX_N);
    for (int c = 0; c < HEX_N; ++c) {
        int idx;
        switch (which) {
        case BB_EAST: idx = c * HEX_N + (HEX_N - 1); break;
        case BB_NORTH: idx = (HEX_N - 1) * HEX_N + c; break;
        case BB_SOUTH: idx = c; break;
        default: idx = c * HEX_N; break; /* West */
        }
        b = bb_set_index(b, idx);
    }
    return b;
}

/* lib.rs:52-61 hex_connects: flood from edges[0] through the mover's stones, must touch edges[1] */
static int hex_connects(bb_t board, bb_t first, bb_t second) {
    int it = 0;
    bb_t flooded = bb_flood6(board, first, &it);
    __atomic_fetch_add(&g_hex_flood_iterations, (uint64_t)it, __ATOMIC_RELAXED);
    return bb_intersects(flooded, second);
}

/* lib.rs:130-144 winner: only the player who moved last is checked;
 * P0 connects north<->south, P1 connects west<->east */
static int hex_winner(const orc_state *s) {
    int last_mover = (s->turn + 2 - 1) % 2;
    bb_t board = hex_bb(&s->board[2 * last_mover]);
    int won = last_mover == 0 ? hex_connects(board, hex_side(BB_NORTH), hex_side(BB_SOUTH))
                              : hex_connects(board, hex_side(BB_WEST), hex_side(BB_EAST));
    return won ? last_mover : -1;
}

/* lib.rs:115-120 gen_moves: every empty cell, ascending */
static int hex_actions(const orc_state *s, uint16_t *out) {
    bb_t legal = bb_not(bb_or(hex_bb(&s->board[0]), hex_bb(&s->board[2])));
    int n = 0, idx;
    while ((idx = bb_pop_lowest(&legal)) >= 0) BB_OP(ORC_OP_LIST_STORE, 1), out[n++] = (uint16_t)idx;
    return n;
}

/* lib.rs:122-125 apply */
static void hex_apply(orc_state *s, uint16_t action) {
    s->board[2 * s->turn + action / 64] |= (uint64_t)1 << (action % 64);
    s->turn ^= 1;
}

/* lib.rs:208-215 is_terminal: winner().is_some(), else a full generate_actions that is empty */
static int hex_is_terminal(const orc_state *s) {
    if (hex_winner(s) >= 0) return 1;
    uint16_t tmp[ORC_MAX_ACTIONS];
    return hex_actions(s, tmp) == 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Dispatch                                                                                    */
/* ------------------------------------------------------------------------------------------ */

void orc_initial_state(int game, orc_state *out) {
    memset(out, 0, sizeof *out);
    switch (game) {
    case ORC_CONNECT4: break; /* connect4/lib.rs:223-235: empty boards, Black to move */
    case ORC_BREAKTHROUGH:
    case ORC_KNIGHTTHROUGH: bt_default(out); break;
    case ORC_OTHELLO: /* othello/lib.rs:372-387 */
        out->board[0] = OTH_INITIAL_BLACK;
        out->board[1] = OTH_INITIAL_WHITE;
        break;
    case ORC_HEX11: break; /* hex-gen/lib.rs:107-113: empty, P0 to move */
    }
}

/* Compact per-player action index ("slot"): the key space of the bigram (NST) tables, where the dense id space of
 * Breakthrough / Knightthrough (src*64 + dst) would be 4096 x 4096 per player.
 *   Connect4 column (7), Othello square or PASS (65), Hex cell (121),
 *   Breakthrough src*3 + kind with kind = 0,1,2 for the west-diagonal, straight, east-diagonal move of that player
 *   (dst - src = -9,-8,-7 for Black, +7,+8,+9 for White; breakthrough/lib.rs:131-156 order),
 *   Knightthrough src*8 + kind with kind indexing the jumps in ascending-dst order (-17,-15,-10,-6,+6,+10,+15,+17). */
int orc_num_slots(int game) {
    switch (game) {
    case ORC_CONNECT4: return 7;
    case ORC_BREAKTHROUGH: return 192;
    case ORC_KNIGHTTHROUGH: return 512;
    case ORC_OTHELLO: return 65;
    case ORC_HEX11: return 121;
    }
    return 0;
}

int orc_action_slot(int game, int player, uint16_t code) {
    static const int kt_delta[8] = {-17, -15, -10, -6, 6, 10, 15, 17};
    if (game == ORC_BREAKTHROUGH || game == ORC_KNIGHTTHROUGH) {
        int src = code >> 8, dst = code & 0xff;
        if (src >= 64 || dst >= 64) return -1;
        int dr = dst / 8 - src / 8, dc = dst % 8 - src % 8;
        if (game == ORC_BREAKTHROUGH) { /* one row toward the far wall, at most one column sideways */
            if (dr != (player ? 1 : -1) || dc < -1 || dc > 1) return -1;
            return src * 3 + dc + 1;
        }
        if (!((abs(dr) == 2 && abs(dc) == 1) || (abs(dr) == 1 && abs(dc) == 2))) return -1;
        for (int q = 0; q < 8; ++q)
            if (dst - src == kt_delta[q]) return src * 8 + q;
        return -1;
    }
    return (int)code < orc_num_slots(game) ? (int)code : -1;
}

int orc_num_actions(int game) {
    switch (game) {
    case ORC_CONNECT4: return 7;
    case ORC_BREAKTHROUGH:
    case ORC_KNIGHTTHROUGH: return 4096; /* src*64 + dst */
    case ORC_OTHELLO: return 65;         /* 64 squares + PASS */
    case ORC_HEX11: return 121;
    }
    return 0;
}

int orc_generate_actions(int game, const orc_state *s, uint16_t *out) {
    switch (game) {
    case ORC_CONNECT4: {
        c4_state t = c4_load(s, 6, 7);
        return c4_moves(&t, out);
    }
    case ORC_BREAKTHROUGH: {
        bt_state t = bt_load(s);
        return bt_moves(&t, out);
    }
    case ORC_KNIGHTTHROUGH: {
        bt_state t = bt_load(s);
        return kt_moves(&t, out);
    }
    case ORC_OTHELLO: return oth_actions(s, out);
    case ORC_HEX11: return hex_actions(s, out);
    }
    return 0;
}

void orc_apply(int game, orc_state *s, uint16_t action) {
    switch (game) {
    case ORC_CONNECT4: {
        c4_state t = c4_load(s, 6, 7);
        c4_apply(&t, action);
        c4_store(s, &t);
        break;
    }
    case ORC_BREAKTHROUGH:
    case ORC_KNIGHTTHROUGH: {
        bt_state t = bt_load(s);
        bt_apply(&t, action);
        bt_store(s, &t);
        break;
    }
    case ORC_OTHELLO: oth_apply(s, action); break;
    case ORC_HEX11: hex_apply(s, action); break;
    }
}

int orc_is_terminal(int game, const orc_state *s) {
    switch (game) {
    case ORC_CONNECT4: {
        c4_state t = c4_load(s, 6, 7);
        return c4_is_terminal(&t);
    }
    case ORC_BREAKTHROUGH:
    case ORC_KNIGHTTHROUGH: return s->flag != 0; /* breakthrough/lib.rs:207-209 */
    case ORC_OTHELLO: return oth_is_terminal(s);
    case ORC_HEX11: return hex_is_terminal(s);
    }
    return 1;
}

int orc_winner(int game, const orc_state *s) {
    switch (game) {
    case ORC_CONNECT4:
    case ORC_BREAKTHROUGH:
    case ORC_KNIGHTTHROUGH: return s->flag ? s->turn : -1; /* connect4/lib.rs:380-386, breakthrough/lib.rs:215-221 */
    case ORC_OTHELLO: return oth_winner(s);
    case ORC_HEX11: return hex_winner(s);
    }
    return -1;
}

/* mcts/src/game.rs:202-211 default terminal_status: is_terminal, then winner */
int orc_terminal_status(int game, const orc_state *s) {
    if (orc_is_terminal(game, s)) {
        int w = orc_winner(game, s);
        return w < 0 ? ORC_DRAW : (w == 0 ? ORC_WINNER_P0 : ORC_WINNER_P1);
    }
    return ORC_NOT_TERMINAL;
}

int orc_player_to_move(int game, const orc_state *s) {
    (void)game;
    return s->turn;
}

/* connect4/lib.rs:534-553 test_draw: random self-play on a small board until a game fills the
 * board with no winner.  The reference draws with SmallRng (not reproducible here); this plays the
 * same loop with a xorshift64* stream so that the property "a drawn 4x5 game exists and reports
 * winner() == None" can be checked through the restated rules. */
int orc_c4_generic_random_game(int rows, int cols, uint64_t seed, int *out_plies, int *out_has_winner) {
    uint64_t x = seed * 0x9e3779b97f4a7c15ull + 0x2545f4914f6cdd1dull;
    c4_state t;
    t.black = bb_empty(rows, cols);
    t.white = bb_empty(rows, cols);
    t.turn = 0;
    t.winner = 0;
    int plies = 0;
    while (!c4_is_terminal(&t)) {
        uint16_t acts[64];
        int n = c4_moves(&t, acts);
        if (n == 0) return -1;
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        uint64_t r = x * 0x9e3779b97f4a7c15ull;
        c4_apply(&t, acts[(r >> 33) % (uint64_t)n]);
        ++plies;
    }
    *out_plies = plies;
    *out_has_winner = t.winner;
    return c4_winner(&t);
}
