/*
 * oracle/naive_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restatements of the loop-and-array oracles the REFERENCE'S OWN TESTS use to pin its
 * bitboard rules, and of the stress sweeps that drive them.  They pin oracle/games.c the
 * same way the reference's tests pin the reference:
 *   - Othello naive_generate_moves / naive_get_flips ...... games/othello/src/lib.rs:239-318
 *   - Othello 100 000-game symmetry sweep ................. mcts-tests/tests/stress.rs:249-454
 *   - Breakthrough / Knightthrough array-board oracle ..... mcts-tests/tests/stress.rs:485-626
 *   - Breakthrough / Knightthrough 5 000-game sweep ....... mcts-tests/tests/stress.rs:638-818
 *   - Hex DFS oracle ...................................... gdl/tests/hex_gen_oracle.rs:34-124
 *   - D4 index symmetries ................................. games/game-core/src/symmetry.rs:88-127
 */
#include "oracle.h"

#include <string.h>

/* games/othello/src/lib.rs:239-248 */
static const int DIRS[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};

static inline int on_board(int r, int c) { return r >= 0 && r < 8 && c >= 0 && c < 8; }
static inline uint64_t coord_bit(int r, int c) { return (uint64_t)1 << (r * 8 + c); }

/* games/othello/src/lib.rs:250-289 */
uint64_t orc_naive_othello_generate_moves(uint64_t player, uint64_t opponent) {
    if (player == 0 || opponent == 0) return 0;
    uint64_t occupied = player | opponent, legal = 0;
    for (int idx = 0; idx < 64; ++idx) {
        if ((occupied >> idx) & 1) continue;
        int sr = idx / 8, sc = idx % 8;
        for (int d = 0; d < 8; ++d) {
            int dr = DIRS[d][0], dc = DIRS[d][1];
            int r = sr + dr, c = sc + dc;
            if (!on_board(r, c)) continue;
            if (!(opponent & coord_bit(r, c))) continue;
            int found = 0;
            for (;;) {
                r += dr;
                c += dc;
                if (!on_board(r, c)) break;
                uint64_t pos = coord_bit(r, c);
                if (player & pos) {
                    found = 1;
                    break;
                }
                if (!(opponent & pos)) break;
            }
            if (found) {
                legal |= coord_bit(sr, sc);
                break; /* break 'dirs */
            }
        }
    }
    return legal;
}

/* games/othello/src/lib.rs:291-321 */
uint64_t orc_naive_othello_get_flips(uint64_t player, uint64_t opponent, uint64_t move_mask) {
    int sq = __builtin_ctzll(move_mask);
    int sr = sq / 8, sc = sq % 8;
    uint64_t flips = 0;
    for (int d = 0; d < 8; ++d) {
        int dr = DIRS[d][0], dc = DIRS[d][1];
        int r = sr + dr, c = sc + dc;
        if (!on_board(r, c)) continue;
        uint64_t line = 0;
        for (;;) {
            uint64_t pos = coord_bit(r, c);
            if (opponent & pos) {
                line |= pos;
            } else if (player & pos) {
                flips |= line;
                break;
            } else {
                break;
            }
            r += dr;
            c += dc;
            if (!on_board(r, c)) break;
        }
    }
    return flips;
}

/* games/game-core/src/symmetry.rs:88-127: identity, flip_cols, flip_rows, transpose,
 * flip_rows.flip_cols, transpose.flip_cols, transpose.flip_rows, transpose.flip_rows.flip_cols */
static int d4_index(int i, int sym) {
    int row = i / 8, col = i % 8;
#define FC(r, c) ((r) * 8 + (7 - (c)))
#define FR(r, c) ((7 - (r)) * 8 + (c))
#define TR(r, c) ((c) * 8 + (r))
    int fc = FC(row, col), fr = FR(row, col);
    switch (sym) {
    case 0: return i;
    case 1: return fc;
    case 2: return fr;
    case 3: return TR(row, col);
    case 4: return FR(fc / 8, fc % 8);
    case 5: return TR(fc / 8, fc % 8);
    case 6: return TR(fr / 8, fr % 8);
    default: {
        int t = FR(fc / 8, fc % 8);
        return TR(t / 8, t % 8);
    }
    }
#undef FC
#undef FR
#undef TR
}
static uint64_t d4_bits(uint64_t b, int sym) {
    uint64_t out = 0;
    for (int i = 0; i < 64; ++i)
        if ((b >> i) & 1) out |= (uint64_t)1 << d4_index(i, sym);
    return out;
}

static uint64_t next_rand(uint64_t *rng) { /* xorshift64*, stress.rs:262-270 */
    uint64_t x = *rng;
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    *rng = x;
    return x * 0x9e3779b97f4a7c15ull;
}

/* mcts-tests/tests/stress.rs:249-454 test_othello_oracle_symmetry_stress */
int orc_stress_othello(uint64_t num_games, uint64_t *games_done, uint64_t *plies_done) {
    uint64_t rng = 0xdeadbeefcafebabeull;
    uint64_t plies_total = 0;
    for (uint64_t game = 0; game < num_games; ++game) {
        orc_state state;
        orc_initial_state(ORC_OTHELLO, &state);
        for (int ply = 0; ply < 128; ++ply) {
            uint64_t player = state.turn == 0 ? state.board[0] : state.board[1];
            uint64_t opponent = state.turn == 0 ? state.board[1] : state.board[0];
            if (state.board[0] & state.board[1]) return 1;
            for (int sym = 0; sym < 8; ++sym) {
                uint64_t bs = d4_bits(state.board[0], sym), ws = d4_bits(state.board[1], sym);
                uint64_t ps = state.turn == 0 ? bs : ws, os = state.turn == 0 ? ws : bs;
                if (orc_othello_generate_moves(ps, os) != orc_naive_othello_generate_moves(ps, os)) return 2;
            }
            uint64_t legal = orc_othello_generate_moves(player, opponent);
            if (legal != orc_naive_othello_generate_moves(player, opponent)) return 3;
            if (legal == 0) {
                uint64_t after = orc_othello_generate_moves(opponent, player);
                if (after != orc_naive_othello_generate_moves(opponent, player)) return 4;
                if (after == 0) break; /* double pass */
                orc_state prod = state;
                orc_apply(ORC_OTHELLO, &prod, ORC_OTHELLO_PASS);
                /* naive_apply(PASS): boards unchanged, turn flipped, last_pass = true (lib.rs:323-332) */
                if (prod.board[0] != state.board[0] || prod.board[1] != state.board[1] || prod.turn != (state.turn ^ 1) ||
                    prod.flag != 1)
                    return 5;
                /* the generated action list must be exactly [PASS] */
                uint16_t acts[ORC_MAX_ACTIONS];
                if (orc_generate_actions(ORC_OTHELLO, &state, acts) != 1 || acts[0] != ORC_OTHELLO_PASS) return 6;
                state = prod;
                ++plies_total;
                continue;
            }
            int num_legal = __builtin_popcountll(legal);
            int choice = (int)(next_rand(&rng) % (uint64_t)num_legal);
            int mv = 0, found = 0;
            for (int i = 0; i < 64; ++i) {
                if ((legal >> i) & 1) {
                    if (found == choice) {
                        mv = i;
                        break;
                    }
                    ++found;
                }
            }
            /* the ascending action list must index the same square */
            uint16_t acts[ORC_MAX_ACTIONS];
            int n = orc_generate_actions(ORC_OTHELLO, &state, acts);
            if (n != num_legal || acts[choice] != mv) return 7;
            uint64_t mv_bb = (uint64_t)1 << mv;
            uint64_t pf = orc_othello_get_flips(player, opponent, mv_bb);
            uint64_t nf = orc_naive_othello_get_flips(player, opponent, mv_bb);
            if (pf != nf) return 8;
            orc_state prod = state;
            orc_apply(ORC_OTHELLO, &prod, (uint16_t)mv);
            /* naive_apply (lib.rs:333-352): xor in the move and the flips */
            uint64_t np = player ^ mv_bb ^ nf, no = opponent ^ nf;
            uint64_t nb = state.turn == 0 ? np : no, nw = state.turn == 0 ? no : np;
            if (prod.board[0] != nb || prod.board[1] != nw || prod.turn != (state.turn ^ 1) || prod.flag != 0) return 9;
            state = prod;
            ++plies_total;
        }
        if (games_done) *games_done = game + 1;
    }
    if (plies_done) *plies_done = plies_total;
    return 0;
}

/* ---- Breakthrough / Knightthrough array-board oracle (stress.rs:485-626) ---- */
/* board cells: -1 empty, 0 Black, 1 White; row 0 = south wall */

int orc_naive_breakthrough_moves(const int8_t board[64], int turn, uint16_t *out) {
    int n = 0;
    for (int src = 0; src < 64; ++src) {
        if (board[src] != turn) continue;
        int row = src / 8, col = src % 8;
        int straight, sw, se;
        if (turn == 0) {
            if (row == 0) continue;
            straight = src - 8;
            sw = col > 0 ? src - 9 : -1;
            se = col < 7 ? src - 7 : -1;
        } else {
            if (row == 7) continue;
            straight = src + 8;
            sw = col > 0 ? src + 7 : -1;
            se = col < 7 ? src + 9 : -1;
        }
        if (board[straight] < 0) out[n++] = (uint16_t)((src << 8) | straight);
        if (sw >= 0 && board[sw] != turn) out[n++] = (uint16_t)((src << 8) | sw);
        if (se >= 0 && board[se] != turn) out[n++] = (uint16_t)((src << 8) | se);
    }
    return n;
}

int orc_naive_knightthrough_moves(const int8_t board[64], int turn, uint16_t *out) {
    static const int off[8][2] = {{2, 1}, {2, -1}, {-2, 1}, {-2, -1}, {1, 2}, {1, -2}, {-1, 2}, {-1, -2}};
    int n = 0;
    for (int src = 0; src < 64; ++src) {
        if (board[src] != turn) continue;
        int row = src / 8, col = src % 8;
        for (int k = 0; k < 8; ++k) {
            int r = row + off[k][0], c = col + off[k][1];
            if (on_board(r, c)) {
                int dst = r * 8 + c;
                if (board[dst] != turn) out[n++] = (uint16_t)((src << 8) | dst);
            }
        }
    }
    return n;
}

static void naive_from_bitboards(uint64_t black, uint64_t white, int8_t board[64]) {
    for (int i = 0; i < 64; ++i) board[i] = (black >> i) & 1 ? 0 : ((white >> i) & 1 ? 1 : -1);
}

static void sort_u16(uint16_t *a, int n) {
    for (int i = 1; i < n; ++i) {
        uint16_t v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            --j;
        }
        a[j + 1] = v;
    }
}

/* stress.rs:638-818 stress_oracle_test!: 5 000 games, seed 0xcafe_babe_dead_beef, choice = rand % n over
 * the NAIVE list, move SETS compared sorted, states compared after every apply */
int orc_stress_through(int game, uint64_t num_games, uint64_t *goal_wins, uint64_t *stuck_games) {
    uint64_t rng = 0xcafebabedeadbeefull;
    uint64_t goals = 0, stuck = 0;
    for (uint64_t g = 0; g < num_games; ++g) {
        orc_state bb;
        orc_initial_state(game, &bb);
        int8_t naive[64];
        naive_from_bitboards(bb.board[0], bb.board[1], naive);
        int naive_turn = 0;
        for (int ply = 0; ply < 512; ++ply) {
            int8_t cur[64];
            naive_from_bitboards(bb.board[0], bb.board[1], cur);
            if (memcmp(cur, naive, 64) != 0) return 1;
            uint16_t a[ORC_MAX_ACTIONS], b[ORC_MAX_ACTIONS], b_unsorted[ORC_MAX_ACTIONS];
            int na = orc_generate_actions(game, &bb, a);
            int nb = game == ORC_BREAKTHROUGH ? orc_naive_breakthrough_moves(naive, naive_turn, b)
                                              : orc_naive_knightthrough_moves(naive, naive_turn, b);
            memcpy(b_unsorted, b, sizeof(uint16_t) * (size_t)nb);
            /* the bitboard list must already be in ascending (src, dst) order: that ORDER is what
             * the playout's index pick observes (bitboard/src/board.rs:864-879) */
            for (int i = 1; i < na; ++i)
                if (a[i - 1] >= a[i]) return 2;
            sort_u16(b, nb);
            if (na != nb || memcmp(a, b, sizeof(uint16_t) * (size_t)na) != 0) return 3;
            int terminal = orc_is_terminal(game, &bb);
            if (na == 0 && !terminal) {
                ++stuck;
                break;
            }
            if (na == 0) break;
            int pick = (int)(next_rand(&rng) % (uint64_t)nb);
            uint16_t mv = b_unsorted[pick];
            orc_state next = bb;
            orc_apply(game, &next, mv);
            /* naive_apply (stress.rs:590-616) */
            int src = mv >> 8, dst = mv & 0xff;
            int8_t piece = naive[src];
            naive[src] = -1;
            naive[dst] = piece;
            int ended = piece == 0 ? dst < 8 : dst >= 56;
            if (!ended) {
                int opp_left = 0;
                for (int i = 0; i < 64; ++i) opp_left |= naive[i] == (piece ^ 1);
                if (!opp_left) ended = 1;
            }
            if (!ended) naive_turn ^= 1;
            naive_from_bitboards(next.board[0], next.board[1], cur);
            if (memcmp(cur, naive, 64) != 0) return 4;
            if (ended) {
                if (!next.flag) return 5; /* reference asserts has_winner() here */
                if (next.turn != naive_turn) return 6;
                ++goals;
                break;
            }
            if (next.turn != naive_turn) return 7;
            bb = next;
        }
    }
    if (goal_wins) *goal_wins = goals;
    if (stuck_games) *stuck_games = stuck;
    return 0;
}

/* gdl/tests/hex_gen_oracle.rs:66-124 HexOracle::winner: DFS components of the last mover's stones over
 * the neighbour set (1,0),(-1,0),(0,1),(0,-1),(1,1),(-1,-1); P0 needs rows 0 and 10, P1 cols 0 and 10 */
int orc_naive_hex_winner(const uint8_t occ0[121], const uint8_t occ1[121], int last_mover) {
    static const int nb[6][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, -1}};
    const uint8_t *stones = last_mover == 0 ? occ0 : occ1;
    uint8_t visited[121];
    memset(visited, 0, sizeof visited);
    for (int start = 0; start < 121; ++start) {
        if (!stones[start] || visited[start]) continue;
        int stack[121], sp = 0, ta = 0, tb = 0;
        stack[sp++] = start;
        visited[start] = 1;
        while (sp) {
            int site = stack[--sp];
            int row = site / 11, col = site % 11;
            if (last_mover == 0) {
                ta |= row == 0;
                tb |= row == 10;
            } else {
                ta |= col == 0;
                tb |= col == 10;
            }
            for (int k = 0; k < 6; ++k) {
                int r = row + nb[k][0], c = col + nb[k][1];
                if (r < 0 || r >= 11 || c < 0 || c >= 11) continue;
                int n = r * 11 + c;
                if (stones[n] && !visited[n]) {
                    visited[n] = 1;
                    stack[sp++] = n;
                }
            }
        }
        if (ta && tb) return last_mover;
    }
    return -1;
}
