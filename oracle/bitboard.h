/*
 * oracle/bitboard.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's generic bitboard,
 * `bitboard/src/board.rs` in thomasmarsh/mcts.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call this.
 *
 * What is restated, with the reference lines each piece follows:
 *   - cell index = row*cols + col, word = idx/64, bit = idx%64, low word first
 *       (board.rs:76-104)
 *   - ascending-bit iteration (board.rs:144-157, 864-879)
 *   - Not masks to the board's real length (board.rs:205-215, 824-836)
 *   - multi-word raw shifts with carry (board.rs:225-258)
 *   - wall masks N/E/S/W (board.rs:298-327)
 *   - masked one-cell shifts; shift_south is UNMASKED (board.rs:360-417)
 *   - flood6 fix-point (board.rs:509-531)
 *
 * One struct covers both storages the five games use: Board<u64> (nw == 1) and
 * Board<[u64;2]> (nw == 2, Hex 11x11).  Dimensions travel with the value, as they
 * do in the reference's `Board { bits, rows, cols }`.
 */
#ifndef ORACLE_BITBOARD_H
#define ORACLE_BITBOARD_H

#include <stdint.h>

typedef struct {
    uint64_t w[2];
    uint8_t rows, cols, nw;
} bb_t;

/* ---- operation counters ("re-derive ops/ply by instrumenting the oracle with op counters") ----
 * Built only into the counting variant of the library (-DORC_COUNT_OPS, _build/liboracle_ops.so).  Every primitive
 * below reports what it executes per 64-bit WORD the board really has; games.c / playout.c report move-list stores,
 * draws and MAST score evaluations.  Plain globals: counting runs are single-threaded. */
enum {
    ORC_OP_LOGIC64 = 0, /* and / or / xor / not / compare on one 64-bit word */
    ORC_OP_SHIFT64,     /* one 64-bit shift (the carry shift of a two-word board counts separately) */
    ORC_OP_POPC64,      /* 64-bit population count */
    ORC_OP_CTZ64,       /* lowest-set-bit search of the ascending iteration (its clear-lowest is 2 LOGIC64) */
    ORC_OP_LIST_STORE,  /* one action written to a materialised action list */
    ORC_OP_DRAW,        /* one random word consumed */
    ORC_OP_SCORE_EVAL,  /* one unigram / bigram score evaluated and compared (MAST-style policies) */
    ORC_OP_PLY,         /* plies played */
    ORC_OP_FLOOD_ITER,  /* flood6 fix-point iterations (Hex) */
    ORC_OP_KINDS
};
#ifdef ORC_COUNT_OPS
extern uint64_t orc_ops[ORC_OP_KINDS];
#define BB_OP(kind, n) (orc_ops[kind] += (uint64_t)(n))
#else
#define BB_OP(kind, n) ((void)0)
#endif

enum { BB_NORTH = 0, BB_EAST = 1, BB_SOUTH = 2, BB_WEST = 3 };

static inline bb_t bb_empty(int rows, int cols) {
    bb_t b;
    b.w[0] = 0;
    b.w[1] = 0;
    b.rows = (uint8_t)rows;
    b.cols = (uint8_t)cols;
    b.nw = (uint8_t)((rows * cols + 63) / 64);
    return b;
}

static inline int bb_len(bb_t b) { return (int)b.rows * (int)b.cols; }

/* board.rs:89-92 */
static inline int bb_get_index(bb_t b, int idx) {
    BB_OP(ORC_OP_SHIFT64, 1);
    BB_OP(ORC_OP_LOGIC64, 1);
    return (int)((b.w[idx / 64] >> (idx % 64)) & 1u);
}
/* board.rs:95-98 */
static inline bb_t bb_set_index(bb_t b, int idx) {
    BB_OP(ORC_OP_SHIFT64, 1);
    BB_OP(ORC_OP_LOGIC64, 1);
    b.w[idx / 64] |= (uint64_t)1 << (idx % 64);
    return b;
}
static inline bb_t bb_from_index(int rows, int cols, int idx) { return bb_set_index(bb_empty(rows, cols), idx); }

static inline bb_t bb_and(bb_t a, bb_t b) {
    BB_OP(ORC_OP_LOGIC64, a.nw);
    a.w[0] &= b.w[0];
    a.w[1] &= b.w[1];
    return a;
}
static inline bb_t bb_or(bb_t a, bb_t b) {
    BB_OP(ORC_OP_LOGIC64, a.nw);
    a.w[0] |= b.w[0];
    a.w[1] |= b.w[1];
    return a;
}
static inline bb_t bb_xor(bb_t a, bb_t b) {
    BB_OP(ORC_OP_LOGIC64, a.nw);
    a.w[0] ^= b.w[0];
    a.w[1] ^= b.w[1];
    return a;
}
static inline int bb_eq(bb_t a, bb_t b) {
    BB_OP(ORC_OP_LOGIC64, a.nw);
    return a.w[0] == b.w[0] && a.w[1] == b.w[1];
}
static inline int bb_none_set(bb_t a) {
    BB_OP(ORC_OP_LOGIC64, a.nw);
    return (a.w[0] | a.w[1]) == 0;
}
static inline int bb_intersects(bb_t a, bb_t b) {
    BB_OP(ORC_OP_LOGIC64, 2 * a.nw);
    return ((a.w[0] & b.w[0]) | (a.w[1] & b.w[1])) != 0;
}
static inline int bb_count_ones(bb_t a) {
    BB_OP(ORC_OP_POPC64, a.nw);
    return __builtin_popcountll(a.w[0]) + __builtin_popcountll(a.w[1]);
}

/* board.rs:205-215: mask of word w covering only bits inside 0..len */
static inline uint64_t bb_word_mask(bb_t b, int w) {
    int total = bb_len(b);
    int start = w * 64;
    if (w >= b.nw || start >= total) return 0;
    if (total - start >= 64) return ~(uint64_t)0;
    return ((uint64_t)1 << (total - start)) - 1;
}

/* board.rs:824-836: `!board` never sets padding bits */
static inline bb_t bb_not(bb_t a) {
    BB_OP(ORC_OP_LOGIC64, a.nw); /* the length mask is a constant: ~a & mask is one operation */
    a.w[0] = ~a.w[0] & bb_word_mask(a, 0);
    a.w[1] = ~a.w[1] & bb_word_mask(a, 1);
    return a;
}

/* board.rs:225-240 raw_shl (rhs < 64), carrying across words */
static inline bb_t bb_shl(bb_t a, int rhs) {
    if (rhs == 0) return a;
    BB_OP(ORC_OP_SHIFT64, 2 * a.nw - 1); /* one shift per word, one carry shift per word boundary */
    BB_OP(ORC_OP_LOGIC64, a.nw - 1);     /* the carry OR */
    bb_t out = a;
    for (int w = a.nw - 1; w >= 0; --w) {
        uint64_t v = a.w[w] << rhs;
        if (w > 0) v |= a.w[w - 1] >> (64 - rhs);
        out.w[w] = v;
    }
    return out;
}

/* board.rs:242-258 raw_shr (rhs < 64) */
static inline bb_t bb_shr(bb_t a, int rhs) {
    if (rhs == 0) return a;
    BB_OP(ORC_OP_SHIFT64, 2 * a.nw - 1);
    BB_OP(ORC_OP_LOGIC64, a.nw - 1);
    bb_t out = a;
    for (int w = 0; w < a.nw; ++w) {
        uint64_t v = a.w[w] >> rhs;
        if (w + 1 < a.nw) v |= a.w[w + 1] << (64 - rhs);
        out.w[w] = v;
    }
    return out;
}

/* board.rs:298-317 wall_words_uncached */
static inline bb_t bb_wall_uncached(int rows, int cols, int dir) {
    bb_t out = bb_empty(rows, cols);
    int limit = (dir == BB_NORTH || dir == BB_SOUTH) ? cols : rows;
    for (int i = 0; i < limit; ++i) {
        int k;
        switch (dir) {
        case BB_NORTH: k = (rows - 1) * cols + i; break;
        case BB_EAST: k = (i + 1) * cols - 1; break;
        case BB_SOUTH: k = i; break;
        default: k = i * cols; break; /* West */
        }
        out.w[k / 64] |= (uint64_t)1 << (k % 64);
    }
    return out;
}

/* The reference's Const<N>-dimension boards recompute walls per call (board.rs:329-333), which rustc
 * constant-folds per monomorphisation.  C has no monomorphisation, so the same constants are held in a
 * table filled once at load time: identical values, and a CPU baseline that is not handicapped by a
 * loop the reference never executes at run time. */
#define BB_WALL_MAX_DIM 12
extern bb_t bb_wall_table[BB_WALL_MAX_DIM + 1][BB_WALL_MAX_DIM + 1][4];
static inline bb_t bb_wall(int rows, int cols, int dir) {
    if (rows <= BB_WALL_MAX_DIM && cols <= BB_WALL_MAX_DIM) return bb_wall_table[rows][cols][dir];
    return bb_wall_uncached(rows, cols, dir);
}
static inline bb_t bb_wall_of(bb_t b, int dir) { return bb_wall(b.rows, b.cols, dir); }

/* board.rs:360-417: one-cell displacements */
static inline bb_t bb_shift_north(bb_t a) { return bb_shl(bb_and(a, bb_not(bb_wall_of(a, BB_NORTH))), a.cols); }
static inline bb_t bb_shift_east(bb_t a) { return bb_shl(bb_and(a, bb_not(bb_wall_of(a, BB_EAST))), 1); }
static inline bb_t bb_shift_south(bb_t a) { return bb_shr(a, a.cols); } /* unmasked, as in the reference */
static inline bb_t bb_shift_west(bb_t a) { return bb_shr(bb_and(a, bb_not(bb_wall_of(a, BB_WEST))), 1); }
static inline bb_t bb_shift_northeast(bb_t a) {
    return bb_shl(bb_and(bb_and(a, bb_not(bb_wall_of(a, BB_NORTH))), bb_not(bb_wall_of(a, BB_EAST))), a.cols + 1);
}
static inline bb_t bb_shift_southwest(bb_t a) {
    return bb_shr(bb_and(bb_and(a, bb_not(bb_wall_of(a, BB_SOUTH))), bb_not(bb_wall_of(a, BB_WEST))), a.cols + 1);
}
static inline bb_t bb_shift_northwest(bb_t a) {
    return bb_shl(bb_and(bb_and(a, bb_not(bb_wall_of(a, BB_NORTH))), bb_not(bb_wall_of(a, BB_WEST))), a.cols - 1);
}
static inline bb_t bb_shift_southeast(bb_t a) {
    return bb_shr(bb_and(bb_and(a, bb_not(bb_wall_of(a, BB_SOUTH))), bb_not(bb_wall_of(a, BB_EAST))), a.cols - 1);
}

/* board.rs:509-531 flood6: all six shifts OR'd in ONE expression, then masked by self */
static inline bb_t bb_flood6(bb_t self, bb_t seed, int *iterations) {
    bb_t flood = bb_and(seed, self);
    int it = 0;
    if (bb_none_set(flood)) {
        if (iterations) *iterations = 0;
        return flood;
    }
    for (;;) {
        BB_OP(ORC_OP_FLOOD_ITER, 1);
        bb_t temp = flood;
        bb_t f = flood;
        f = bb_or(f, bb_shift_north(flood));
        f = bb_or(f, bb_shift_south(flood));
        f = bb_or(f, bb_shift_east(flood));
        f = bb_or(f, bb_shift_west(flood));
        f = bb_or(f, bb_shift_northeast(flood));
        f = bb_or(f, bb_shift_southwest(flood));
        flood = bb_and(f, self);
        ++it;
        if (bb_eq(flood, temp)) break;
    }
    if (iterations) *iterations = it;
    return flood;
}

/* board.rs:864-879 `for idx in board`: pops the lowest set bit, low word first.
 * Returns -1 when exhausted. */
static inline int bb_pop_lowest(bb_t *b) {
    for (int w = 0; w < b->nw; ++w) {
        uint64_t word = b->w[w];
        if (word != 0) {
            BB_OP(ORC_OP_CTZ64, 1);
            BB_OP(ORC_OP_LOGIC64, 2);
            int bit = __builtin_ctzll(word);
            b->w[w] = word & (word - 1);
            return w * 64 + bit;
        }
    }
    return -1;
}

#endif
