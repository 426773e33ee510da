// host_games.h -- host-side rules of the five games, used ONLY by the tree driver (tree_search.cu) to
// expand nodes and step leaf states.  In the reference these calls are the crate's own `Game` impls
// (`generate_actions`, `apply`, `terminal_status`); a Rust host would keep using them.  The C++ driver
// needs its own copy because no Rust toolchain exists in this image.  Action ORDER follows the reference
// (child order == generate_actions order, search/shared.rs:220-289).
//
// Reference lines: connect4/lib.rs:195-339,372-386; breakthrough/lib.rs:131-188; knightthrough/lib.rs:126-190;
// othello/lib.rs:132-234,484-596; hex-gen/lib.rs:52-145,199-223; default terminal_status game.rs:202-211.
#pragma once

#include <stdint.h>
#include <string.h>

#include "mcts_rollout.h"

namespace mr_host {

inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

// ---------------------------------------------------------------- Connect4 (row-major 6x7, idx = row*7 + col)
inline int c4_height(uint64_t occ, int col) {
    int row = 0;
    while (row < 6 && ((occ >> (row * 7 + col)) & 1)) ++row;
    return row;
}
inline bool c4_line(uint64_t b, int idx, int dr, int dc) {
    // stones of `b` through cell idx along (dr, dc), counting both directions
    int row = idx / 7, col = idx % 7, cnt = 1;
    for (int s = -1; s <= 1; s += 2) {
        int r = row + s * dr, c = col + s * dc;
        while (r >= 0 && r < 6 && c >= 0 && c < 7 && ((b >> (r * 7 + c)) & 1)) {
            ++cnt;
            r += s * dr;
            c += s * dc;
        }
    }
    return cnt >= 4;
}

// ---------------------------------------------------------------- Othello
inline uint64_t oth_ray_flips(uint64_t P, uint64_t O, int sq, int dr, int dc) {
    int r = sq / 8 + dr, c = sq % 8 + dc;
    uint64_t line = 0;
    while (r >= 0 && r < 8 && c >= 0 && c < 8) {
        const uint64_t b = 1ull << (r * 8 + c);
        if (O & b) line |= b;
        else if (P & b) return line;
        else return 0;
        r += dr;
        c += dc;
    }
    return 0;
}
inline uint64_t oth_flips(uint64_t P, uint64_t O, int sq) {
    uint64_t f = 0;
    for (int dr = -1; dr <= 1; ++dr)
        for (int dc = -1; dc <= 1; ++dc)
            if (dr || dc) f |= oth_ray_flips(P, O, sq, dr, dc);
    return f;
}
// legal squares of the player P: an empty square at the far end of a run of opponent discs that starts next to an own
// disc, in any of the eight directions (the move SET of othello/lib.rs:132-180; the set has no order).  Fill by repeated
// shifts; the opponent board is cut to columns 1..6 for the directions that change column, so nothing wraps around a row.
inline uint64_t oth_moves(uint64_t P, uint64_t O) {
    const uint64_t inner = O & 0x7e7e7e7e7e7e7e7eull;
    uint64_t m = 0;
    const int shifts[4] = {1, 7, 9, 8};
    for (int d = 0; d < 4; ++d) {
        const int s = shifts[d];
        const uint64_t M = s == 8 ? O : inner;
        uint64_t up = M & (P << s), dn = M & (P >> s);
        for (int i = 0; i < 5; ++i) {
            up |= M & (up << s);
            dn |= M & (dn >> s);
        }
        m |= (up << s) | (dn >> s);
    }
    return m & ~(P | O);
}
inline uint64_t oth_moves_by_scan(uint64_t P, uint64_t O) { // the definition, square by square (checked against the fill)
    uint64_t m = 0;
    const uint64_t empty = ~(P | O);
    for (int sq = 0; sq < 64; ++sq)
        if (((empty >> sq) & 1) && oth_flips(P, O, sq)) m |= 1ull << sq;
    return m;
}

// ---------------------------------------------------------------- Hex 11x11
inline bool hex_get(const mr_leaf &s, int pl, int idx) { return (s.board[2 * pl + idx / 64] >> (idx % 64)) & 1; }
inline bool hex_connected(const mr_leaf &s, int pl) {
    // P0: row 10 <-> row 0; P1: col 0 <-> col 10; neighbours N,S,E,W,NE(+1,+1),SW(-1,-1)
    static const int nb[6][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, -1}};
    bool seen[121];
    int stack[121], sp = 0;
    memset(seen, 0, sizeof seen);
    for (int i = 0; i < 11; ++i) {
        const int idx = pl == 0 ? 10 * 11 + i : i * 11;
        if (hex_get(s, pl, idx)) {
            seen[idx] = true;
            stack[sp++] = idx;
        }
    }
    while (sp) {
        const int idx = stack[--sp], row = idx / 11, col = idx % 11;
        if (pl == 0 ? row == 0 : col == 10) return true;
        for (int k = 0; k < 6; ++k) {
            const int r = row + nb[k][0], c = col + nb[k][1];
            if (r < 0 || r > 10 || c < 0 || c > 10) continue;
            const int n = r * 11 + c;
            if (!seen[n] && hex_get(s, pl, n)) {
                seen[n] = true;
                stack[sp++] = n;
            }
        }
    }
    return false;
}

// ---------------------------------------------------------------- the Game surface
// generate_actions: codes as in the C ABI (column / (src<<8)|dst / square, 64 = PASS / cell)
inline int generate_actions(int game, const mr_leaf &s, uint16_t *out) {
    int n = 0;
    switch (game) {
    case MR_CONNECT4: {
        if (s.flag) return 0;
        const uint64_t occ = s.board[0] | s.board[1];
        for (int col = 0; col < 7; ++col)
            if (c4_height(occ, col) < 6) out[n++] = (uint16_t)col;
        return n;
    }
    case MR_BREAKTHROUGH: {
        if (s.flag) return 0;
        const uint64_t own = s.turn ? s.board[1] : s.board[0], occ = s.board[0] | s.board[1];
        const int fwd = s.turn ? 8 : -8;
        for (int src = 0; src < 64; ++src) {
            if (!((own >> src) & 1)) continue;
            const int row = src / 8, col = src % 8;
            if (s.turn ? row == 7 : row == 0) continue;
            const int f = src + fwd; // ascending dst: west diagonal, straight, east diagonal
            if (col > 0 && !((own >> (f - 1)) & 1)) out[n++] = (uint16_t)((src << 8) | (f - 1));
            if (!((occ >> f) & 1)) out[n++] = (uint16_t)((src << 8) | f);
            if (col < 7 && !((own >> (f + 1)) & 1)) out[n++] = (uint16_t)((src << 8) | (f + 1));
        }
        return n;
    }
    case MR_KNIGHTTHROUGH: {
        if (s.flag) return 0;
        static const int off[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}}; // ascending dst
        const uint64_t own = s.turn ? s.board[1] : s.board[0];
        for (int src = 0; src < 64; ++src) {
            if (!((own >> src) & 1)) continue;
            const int row = src / 8, col = src % 8;
            for (int k = 0; k < 8; ++k) {
                const int r = row + off[k][0], c = col + off[k][1];
                if (r < 0 || r > 7 || c < 0 || c > 7) continue;
                const int dst = r * 8 + c;
                if (!((own >> dst) & 1)) out[n++] = (uint16_t)((src << 8) | dst);
            }
        }
        return n;
    }
    case MR_OTHELLO: {
        const uint64_t P = s.turn ? s.board[1] : s.board[0], O = s.turn ? s.board[0] : s.board[1];
        const uint64_t m = oth_moves(P, O);
        if (!m) {
            out[n++] = 64;
            return n;
        }
        for (int sq = 0; sq < 64; ++sq)
            if ((m >> sq) & 1) out[n++] = (uint16_t)sq;
        return n;
    }
    case MR_HEX11: {
        for (int idx = 0; idx < 121; ++idx)
            if (!hex_get(s, 0, idx) && !hex_get(s, 1, idx)) out[n++] = (uint16_t)idx;
        return n;
    }
    }
    return 0;
}

inline void apply(int game, mr_leaf &s, uint16_t a) {
    switch (game) {
    case MR_CONNECT4: {
        const uint64_t occ = s.board[0] | s.board[1];
        const int idx = c4_height(occ, a) * 7 + a;
        uint64_t &me = s.board[s.turn];
        me |= 1ull << idx;
        if (c4_line(me, idx, 1, 0) || c4_line(me, idx, 0, 1) || c4_line(me, idx, 1, 1) || c4_line(me, idx, 1, -1)) s.flag = 1;
        else s.turn ^= 1;
        return;
    }
    case MR_BREAKTHROUGH:
    case MR_KNIGHTTHROUGH: {
        const int src = a >> 8, dst = a & 0xff;
        uint64_t &me = s.board[s.turn], &op = s.board[s.turn ^ 1];
        me = (me | (1ull << dst)) & ~(1ull << src);
        op &= ~(1ull << dst);
        if (s.turn ? dst >= 56 : dst < 8) s.flag = 1;
        else s.turn ^= 1;
        return;
    }
    case MR_OTHELLO: {
        if (a == 64) {
            s.flag = 1;
            s.turn ^= 1;
            return;
        }
        uint64_t &P = s.board[s.turn], &O = s.board[s.turn ^ 1];
        const uint64_t f = oth_flips(P, O, a);
        P |= f | (1ull << a);
        O &= ~f;
        s.flag = 0;
        s.turn ^= 1;
        return;
    }
    case MR_HEX11:
        s.board[2 * s.turn + a / 64] |= 1ull << (a % 64);
        s.turn ^= 1;
        return;
    }
}

// TerminalStatus as mr_outcome
inline int terminal_status(int game, const mr_leaf &s) {
    switch (game) {
    case MR_CONNECT4:
        if (s.flag) return s.turn ? MR_WINNER_P1 : MR_WINNER_P0;
        return popc64(s.board[0] | s.board[1]) == 42 ? MR_DRAW : MR_NOT_TERMINAL;
    case MR_BREAKTHROUGH:
    case MR_KNIGHTTHROUGH:
        return s.flag ? (s.turn ? MR_WINNER_P1 : MR_WINNER_P0) : MR_NOT_TERMINAL;
    case MR_OTHELLO: {
        bool term = popc64(s.board[0] | s.board[1]) == 64;
        if (!term && s.flag) {
            const uint64_t P = s.turn ? s.board[1] : s.board[0], O = s.turn ? s.board[0] : s.board[1];
            term = oth_moves(P, O) == 0;
        }
        if (!term) return MR_NOT_TERMINAL;
        const int b = popc64(s.board[0]), w = popc64(s.board[1]);
        return b > w ? MR_WINNER_P0 : (b < w ? MR_WINNER_P1 : MR_DRAW);
    }
    case MR_HEX11: {
        const int last = s.turn ^ 1;
        if (hex_connected(s, last)) return last ? MR_WINNER_P1 : MR_WINNER_P0;
        const uint64_t full0 = ~0ull, full1 = (1ull << 57) - 1;
        const bool full = (s.board[0] | s.board[2]) == full0 && ((s.board[1] | s.board[3]) & full1) == full1;
        return full ? MR_DRAW : MR_NOT_TERMINAL;
    }
    }
    return MR_DRAW;
}

// dense action id (MAST / AMAF tables)
inline int action_id(int game, uint16_t a) {
    if (game == MR_BREAKTHROUGH || game == MR_KNIGHTTHROUGH) return (a >> 8) * 64 + (a & 0xff);
    return a;
}

// compact per-player action slot (the index of the slot-space tables, include/mcts_rollout.h "N-gram Selection
// Technique") of an action that IS legal for `player` -- the unchecked inline form of mr_action_slot for the driver's
// expansion loop
inline int action_slot(int game, int player, uint16_t a) {
    if (game == MR_BREAKTHROUGH) {
        const int src = a >> 8, dst = a & 0xff;
        return src * 3 + (dst - src - (player ? 7 : -9));
    }
    if (game == MR_KNIGHTTHROUGH) {
        static const int8_t kind_of_delta[35] = {0, -1, 1, -1, -1, -1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1,
                                                 -1, -1, -1, -1, -1, 4, -1, -1, -1, 5, -1, -1, -1, -1, 6, -1, 7}; // delta + 17
        const int src = a >> 8, dst = a & 0xff;
        return src * 8 + kind_of_delta[dst - src + 17];
    }
    return a;
}

// ---- D4 symmetry of the 8x8 board (games/game-core/src/symmetry.rs:88-147, :357-374), word-parallel ----
// Image k of a bitboard in the reference's order: identity, flip_cols, flip_rows, transpose, flip_rows o flip_cols,
// transpose o flip_cols, transpose o flip_rows, transpose o flip_rows o flip_cols (a set square i moves to
// index_symmetries(i)[k]; square = row * 8 + col).
inline uint64_t flip_cols8(uint64_t x) { // col -> 7 - col: reverse the bits of every byte
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    return ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
}
inline uint64_t flip_rows8(uint64_t x) { return __builtin_bswap64(x); } // row -> 7 - row: reverse the bytes
inline uint64_t transpose8(uint64_t x) { // (row, col) -> (col, row): three delta swaps across the main diagonal
    uint64_t t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;
    x ^= t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull;
    x ^= t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull;
    return x ^ t ^ (t << 28);
}
inline uint64_t board_symmetry8(uint64_t x, int k) {
    switch (k) {
    case 0: return x;
    case 1: return flip_cols8(x);
    case 2: return flip_rows8(x);
    case 3: return transpose8(x);
    case 4: return flip_rows8(flip_cols8(x));
    case 5: return transpose8(flip_cols8(x));
    case 6: return transpose8(flip_rows8(x));
    default: return transpose8(flip_rows8(flip_cols8(x)));
    }
}
// the square whose image under symmetry k is `i` (symmetry.rs:134-147 invert_symmetry)
inline int invert_symmetry8(int i, int k) {
    auto fc = [](int v) { return (v & 56) | (7 - (v & 7)); };
    auto fr = [](int v) { return (56 - (v & 56)) | (v & 7); };
    auto tr = [](int v) { return ((v & 7) << 3) | (v >> 3); };
    switch (k) {
    case 0: return i;
    case 1: return fc(i);
    case 2: return fr(i);
    case 3: return tr(i);
    case 4: return fc(fr(i));
    case 5: return fc(tr(i));
    case 6: return fr(tr(i));
    default: return fc(fr(tr(i)));
    }
}
// Othello's canonical orientation (othello/lib.rs:26, 436-443, 618-646): up to SYMMETRY_PLY_LIMIT = 20 discs, the first
// symmetry whose image of (black, white) is lexicographically minimal; past it the identity.  `canon` receives the
// state in that orientation (turn and pass flag are not geometric).
inline int othello_canonical(const mr_leaf &s, mr_leaf &canon) {
    canon = s;
    if (__builtin_popcountll(s.board[0] | s.board[1]) > 20) return 0;
    // the eight images of both boards, sharing the intermediate flips (7 word-parallel transforms per board)
    uint64_t img[2][8];
    for (int c = 0; c < 2; ++c) {
        const uint64_t x = s.board[c], fc = flip_cols8(x), fr = flip_rows8(x), rot = flip_rows8(fc);
        img[c][0] = x;
        img[c][1] = fc;
        img[c][2] = fr;
        img[c][3] = transpose8(x);
        img[c][4] = rot;
        img[c][5] = transpose8(fc);
        img[c][6] = transpose8(fr);
        img[c][7] = transpose8(rot);
    }
    int best = 0;
    for (int k = 1; k < 8; ++k)
        if (img[0][k] < img[0][best] || (img[0][k] == img[0][best] && img[1][k] < img[1][best])) best = k;
    canon.board[0] = img[0][best];
    canon.board[1] = img[1][best];
    return best;
}

} // namespace mr_host
