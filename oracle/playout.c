/*
 * oracle/playout.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's playout loop and the three in-scope
 * simulate policies, driven by the counter-based Philox4x32-10 draw discipline
 * this repo defines (DESIGN.md "Draw discipline"), plus the per-playout
 * statistics backprop consumes.
 *
 * Reference lines followed (paths relative to the reference root):
 *   playout loop ............ mcts/src/strategies/mcts/simulate.rs:84-143 (default), :173-216 (Uniform)
 *   random_action default ... mcts/src/game.rs:177-185
 *   terminal_status default . mcts/src/game.rs:202-211
 *   utilities ............... mcts/src/game.rs:128-141, :278-286; backprop.rs:686-690 (non-terminal => draw)
 *   EpsilonGreedy ........... simulate.rs:538-570 (default epsilon 0.1, :525)
 *   Mast / unigram_score .... simulate.rs:794-848
 *   random_best + PRIMES .... mcts/src/util.rs:129-163
 *   DecisiveMove (Win) ...... simulate.rs:636-687 (loser.or(draw) fallback :672-687)
 *   GLOBAL/MAST write ....... backprop.rs:857-870
 *   Nst ..................... simulate.rs:852-930 (bigram score with a visit-count backoff to the unigram score);
 *                             bigram write backprop.rs:885-899
 *   Lgr2 (LGRF-2) ........... simulate.rs:1036-1142 (2-ply reply, else the inner Lgr); table write with forgetting
 *                             backprop.rs:926-966
 *   Lgr (LGR-1) ............. simulate.rs:936-1033 (reply if recorded and legal, else inner); table write
 *                             backprop.rs:908-923 (last write wins, only players whose utility is the maximum)
 *   GRAVE/AMAF occurrences .. backprop.rs:565-640
 *
 * The reference draws from rand 0.8's SmallRng (crates.io, absent from /root/reference,
 * unpinned by any reference test).  Here every random choice is a pure function of
 * (seed, leaf_id, playout_id, ply, stream):
 *   uniform / decisive : x0 = Philox4x32-10(key = (seed_lo, seed_hi), ctr = (leaf_id, playout_id, ply >> 2, 0))[ply & 3]
 *   epsilon-MAST       : block = Philox4x32-10(key, ctr = (leaf_id, playout_id, ply >> 1, 1));
 *                        x0 = block[2*(ply & 1)] (move choice), x1 = block[2*(ply & 1) + 1] (epsilon test)
 *   uniform pick over n      : index = mulhi32(x0, n)
 *   epsilon test             : explore iff x1 < eps_threshold (eps_threshold = floor(eps * 2^32))
 *   MAST greedy (random_best): r = mulhi32(x0, 16 n); i = r / 16; stride = PRIMES[r % 16]
 */
#include "oracle.h"
#include "bitboard.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- Philox4x32-10 (Salmon et al., SC'11; the Random123 reference constants) ---- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

uint32_t orc_draw(uint64_t seed, uint32_t leaf_id, uint32_t playout_id, uint32_t ply, uint32_t stream) {
    uint32_t ctr[4] = {leaf_id, playout_id, ply >> 2, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    BB_OP(ORC_OP_DRAW, 1);
    return out[ply & 3];
}

void orc_draw_pair(uint64_t seed, uint32_t leaf_id, uint32_t playout_id, uint32_t ply, uint32_t *x0, uint32_t *x1) {
    uint32_t ctr[4] = {leaf_id, playout_id, ply >> 1, 1u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    BB_OP(ORC_OP_DRAW, 2);
    *x0 = out[2 * (ply & 1)];
    *x1 = out[2 * (ply & 1) + 1];
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* mcts/src/util.rs:129-132 */
static const uint32_t PRIMES[16] = {14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991,
                                    53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693};

static inline int action_id(int game, uint16_t a) {
    if (game == ORC_BREAKTHROUGH || game == ORC_KNIGHTTHROUGH) return (int)(a >> 8) * 64 + (int)(a & 0xff);
    return (int)a;
}

/* simulate.rs:806-817 unigram_score: unvisited => 1.0, else score / visits (f64) */
static inline double unigram_score(const orc_action_stat *tab, int id) {
    if (tab[id].visits > 0) return (double)tab[id].score / (double)tab[id].visits;
    return 1.0;
}

/* util.rs:135-163 random_best over a score array; r is the single combined draw */
static int random_best_index(const double *score, int n, uint32_t r) {
    int i = (int)(r / 16u);
    uint32_t stride = PRIMES[r % 16u];
    double best_score = -INFINITY;
    int best = -1;
    for (int k = 0; k < n; ++k) {
        if (score[i] > best_score) {
            best_score = score[i];
            best = i;
        }
        i = (int)(((uint32_t)i + stride) % (uint32_t)n);
    }
    return best;
}

/* simulate.rs:636-735 DecisiveMove::choose, modes Win / WinLoss / WinLossDraw / AntiDecisive */
static int decisive_choose(int game, int mode, const orc_state *state, const uint16_t *available, int n, int player) {
    int draw = -1, loser = -1;
    if (mode == ORC_DECISIVE_ANTI) { /* :689-735 */
        /* pass 1: an immediate win of the mover, wherever it is in the list */
        for (int i = 0; i < n; ++i) {
            orc_state child = *state;
            orc_apply(game, &child, available[i]);
            int st = orc_terminal_status(game, &child);
            if ((st == ORC_WINNER_P0 && player == 0) || (st == ORC_WINNER_P1 && player == 1)) return i;
        }
        /* pass 2: the first move with a non-terminal child that leaves the opponent no immediately winning reply
         * (any Winner(_) among the replies counts) */
        uint16_t replies[ORC_MAX_ACTIONS];
        for (int i = 0; i < n; ++i) {
            orc_state child = *state;
            orc_apply(game, &child, available[i]);
            if (orc_terminal_status(game, &child) != ORC_NOT_TERMINAL) continue;
            int nr = orc_generate_actions(game, &child, replies);
            int opponent_can_win = 0;
            for (int j = 0; j < nr && !opponent_can_win; ++j) {
                orc_state g = child;
                orc_apply(game, &g, replies[j]);
                int st = orc_terminal_status(game, &g);
                opponent_can_win = st == ORC_WINNER_P0 || st == ORC_WINNER_P1;
            }
            if (!opponent_can_win) return i;
        }
        return -1;
    }
    for (int i = 0; i < n; ++i) {
        orc_state child = *state;
        orc_apply(game, &child, available[i]);
        int st = orc_terminal_status(game, &child);
        if (mode == ORC_DECISIVE_WINLOSSDRAW) { /* :645-656 any terminal child */
            if (st != ORC_NOT_TERMINAL) return i;
        } else if (mode == ORC_DECISIVE_WINLOSS) { /* :658-669 any winner at once, else the last draw */
            if (st == ORC_WINNER_P0 || st == ORC_WINNER_P1) return i;
            if (st == ORC_DRAW) draw = i;
        } else { /* Win, :671-687 */
            if (st == ORC_WINNER_P0 || st == ORC_WINNER_P1) {
                int winner = st == ORC_WINNER_P0 ? 0 : 1;
                if (winner == player) return i;
                loser = i;
            } else if (st == ORC_DRAW) {
                draw = i;
            }
        }
    }
    if (mode == ORC_DECISIVE_WINLOSSDRAW) return -1;
    if (mode == ORC_DECISIVE_WINLOSS) return draw;
    return loser >= 0 ? loser : draw; /* loser.or(draw) */
}

void orc_playout(const orc_batch_desc *d, const orc_state *leaf, uint32_t leaf_id, uint32_t playout_id,
                 const orc_action_stat *mast, uint16_t *trace, orc_playout_record *rec, orc_state *final_state) {
    const int game = d->game;
    const int na = orc_num_actions(game);
    orc_state state = *leaf;
    state.prev_plus1 = 0; /* call context, not game state */
    state.prev2_plus1 = 0;
    int prev = leaf->prev_plus1 ? (int)leaf->prev_plus1 - 1 : -1; /* playout()'s prev_action */
    int own_prev[2] = {-1, -1};                                   /* simulate.rs:98: filled only inside this playout */
    uint32_t depth = 0;
    int status, end_type;
    uint16_t available[ORC_MAX_ACTIONS];
    double score[ORC_MAX_ACTIONS];
    const uint32_t max_depth = d->max_playout_depth ? d->max_playout_depth : 0xffffffffu;

    for (;;) {
        status = orc_terminal_status(game, &state);
        if (status != ORC_NOT_TERMINAL) {
            end_type = ORC_END_TERMINAL;
            break;
        }
        if (depth >= max_depth) {
            end_type = ORC_END_TURN_LIMIT;
            break;
        }
        int n = orc_generate_actions(game, &state, available);
        if (n == 0) {
            end_type = ORC_END_NO_ACTIONS; /* NaturalEnd with NotTerminal: scored as a draw */
            break;
        }
        int player = orc_player_to_move(game, &state);
        int pick = -1;
        /* DecisiveMove::select_move (simulate.rs:742-760): the decisive scan, else the inner strategy */
        if (d->policy >= ORC_DECISIVE_WIN && d->policy <= ORC_DECISIVE_ANTI) pick = decisive_choose(game, d->policy, &state, available, n, player);
        if (d->policy == ORC_DECISIVE_MAST_WIN) pick = decisive_choose(game, ORC_DECISIVE_WIN, &state, available, n, player);
        if (d->policy == ORC_DECISIVE_MAST_WINLOSS) pick = decisive_choose(game, ORC_DECISIVE_WINLOSS, &state, available, n, player);
        if (pick < 0) {
            uint32_t x0;
            int greedy = 0;
            if (d->policy == ORC_EPS_MAST || d->policy == ORC_DECISIVE_MAST_WIN || d->policy == ORC_DECISIVE_MAST_WINLOSS ||
                d->policy == ORC_EPS_NST) {
                uint32_t x1;
                orc_draw_pair(d->seed, leaf_id, playout_id, depth, &x0, &x1);
                greedy = !((uint64_t)x1 < d->eps_threshold);
            } else if (d->policy == ORC_EPS_LGR || d->policy == ORC_EPS_LGR2 || d->policy == ORC_EPS_LGR2_MAST) {
                uint32_t x1;
                orc_draw_pair(d->seed, leaf_id, playout_id, depth, &x0, &x1);
                int explore = (uint64_t)x1 < d->eps_threshold;
                if (!explore && d->policy != ORC_EPS_LGR && own_prev[player] >= 0 && prev >= 0 && d->replies2) {
                    /* Lgr2::select_move, simulate.rs:1113-1141: the reply to (own previous move, opponent's answer) */
                    const int S = orc_num_slots(game);
                    uint16_t reply = d->replies2[((size_t)player * S + (size_t)orc_action_slot(game, player, (uint16_t)own_prev[player])) * S +
                                                 (size_t)orc_action_slot(game, 1 - player, (uint16_t)prev)];
                    if (reply)
                        for (int i = 0; i < n; ++i)
                            if (available[i] == (uint16_t)(reply - 1)) pick = i;
                }
                if (pick < 0 && !explore && prev >= 0 && d->replies) { /* Lgr::select_move, simulate.rs:1013-1032 */
                    uint16_t reply = d->replies[(size_t)player * na + action_id(game, (uint16_t)prev)];
                    if (reply)
                        for (int i = 0; i < n; ++i)
                            if (available[i] == (uint16_t)(reply - 1)) pick = i;
                }
                /* Lgr<G, Mast>: the innermost strategy is the plain greedy Mast (no second epsilon test) */
                if (d->policy == ORC_EPS_LGR2_MAST && pick < 0 && !explore) greedy = 1;
            } else {
                x0 = orc_draw(d->seed, leaf_id, playout_id, depth, 0);
            }
            if (greedy) {
                const orc_action_stat *tab = mast + (size_t)player * na;
                BB_OP(ORC_OP_SCORE_EVAL, n);
                for (int i = 0; i < n; ++i) score[i] = unigram_score(tab, action_id(game, available[i]));
                if (d->policy == ORC_EPS_NST && prev >= 0 && d->bigram) { /* Nst::select_move, simulate.rs:898-930 */
                    const int S = orc_num_slots(game);
                    const orc_action_stat *row = d->bigram + ((size_t)player * S + (size_t)orc_action_slot(game, 1 - player, (uint16_t)prev)) * S;
                    for (int i = 0; i < n; ++i) {
                        const orc_action_stat *b = &row[orc_action_slot(game, player, available[i])];
                        if (b->visits >= d->nst_threshold && b->visits > 0) score[i] = (double)b->score / (double)b->visits;
                    }
                }
                pick = random_best_index(score, n, mulhi32(x0, 16u * (uint32_t)n));
            } else if (pick < 0) { /* (an LGR reply may already be chosen) */
                pick = (int)mulhi32(x0, (uint32_t)n);
            }
        }
        uint16_t action = available[pick];
        prev = action;
        own_prev[player] = action;
        if (trace && depth < d->max_trace) trace[depth] = action;
        orc_apply(game, &state, action);
        BB_OP(ORC_OP_PLY, 1);
        ++depth;
    }
    rec->depth = depth;
    rec->end_type = (uint8_t)end_type;
    rec->outcome = (uint8_t)(status == ORC_NOT_TERMINAL ? ORC_DRAW : status);
    rec->pad[0] = rec->pad[1] = 0;
    if (final_state) *final_state = state;
}

/* ---- batch ---- */

typedef struct {
    const orc_batch_desc *d;
    const orc_state *leaves;
    uint32_t n_leaves;
    const orc_action_stat *mast;
    orc_leaf_result *results;
    orc_action_stat *amaf_out;
    orc_action_stat *mast_delta; /* thread-private when multi-threaded */
    uint16_t *trace_out;
    orc_playout_record *rec_out;
    orc_state *final_out;
    uint64_t begin, end; /* global playout index range [begin, end) */
    uint64_t *reply_order;  /* thread-private */
    uint16_t *reply_action; /* thread-private */
    orc_action_stat *bigram_delta; /* thread-private */
} batch_job;

static void add_trace_stats(int game, const orc_state *leaf, const uint16_t *trace, uint32_t depth, int outcome,
                            orc_action_stat *table /* [2][na] */) {
    /* every (action, player) occurrence: visits += 1, score += utilities[player]
     * (backprop.rs:619-640 for GRAVE, :857-870 for GLOBAL/MAST).  Players alternate strictly
     * from the leaf's mover in all five games (Othello passes are explicit actions). */
    const int na = orc_num_actions(game);
    for (uint32_t t = 0; t < depth; ++t) {
        int p = (leaf->turn + (int)t) & 1;
        int u = outcome == ORC_DRAW ? 0 : ((outcome == ORC_WINNER_P0 ? 0 : 1) == p ? 1 : -1);
        orc_action_stat *e = &table[(size_t)p * na + action_id(game, trace[t])];
        e->visits += 1;
        e->score += u;
    }
}

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    const orc_batch_desc *d = j->d;
    const uint32_t k = d->playouts_per_leaf;
    const int na = orc_num_actions(d->game);
    const int need_trace = j->amaf_out || j->mast_delta || j->reply_order || d->last_win_out || j->bigram_delta;
    uint32_t cap = d->max_trace;
    uint16_t *scratch = NULL;
    orc_batch_desc local = *d;
    if (need_trace) {
        uint32_t need = d->max_playout_depth ? d->max_playout_depth : 4096;
        if (need > cap) cap = need;
        scratch = (uint16_t *)malloc(sizeof(uint16_t) * cap);
        local.max_trace = cap;
    }
    for (uint64_t g = j->begin; g < j->end; ++g) {
        uint32_t li = (uint32_t)(g / k), pi = (uint32_t)(g % k);
        orc_playout_record rec;
        uint16_t *user_trace = j->trace_out ? j->trace_out + g * d->max_trace : NULL;
        uint16_t *tr = need_trace ? scratch : user_trace;
        orc_playout(&local, &j->leaves[li], d->leaf_id_base + li, pi, j->mast, tr, &rec,
                    j->final_out ? &j->final_out[g] : NULL);
        if (need_trace && user_trace) {
            uint32_t ncopy = rec.depth < d->max_trace ? rec.depth : d->max_trace;
            memcpy(user_trace, scratch, sizeof(uint16_t) * ncopy);
        }
        if (j->rec_out) j->rec_out[g] = rec;
        orc_leaf_result *r = &j->results[li];
        /* a leaf's playouts may be split over threads: atomics keep the sums exact */
        if (rec.outcome == ORC_WINNER_P0)
            __atomic_fetch_add(&r->wins[0], 1u, __ATOMIC_RELAXED);
        else if (rec.outcome == ORC_WINNER_P1)
            __atomic_fetch_add(&r->wins[1], 1u, __ATOMIC_RELAXED);
        else
            __atomic_fetch_add(&r->draws, 1u, __ATOMIC_RELAXED);
        if (rec.end_type == ORC_END_TURN_LIMIT) __atomic_fetch_add(&r->cutoffs, 1u, __ATOMIC_RELAXED);
        __atomic_fetch_add(&r->sum_depth, (uint64_t)rec.depth, __ATOMIC_RELAXED);
        if (j->amaf_out) {
            /* per-leaf table; leaves never straddle threads unless k is split, so use a lock-free
             * private pass: occurrences are added with atomics */
            orc_action_stat *tab = j->amaf_out + (size_t)li * 2 * na;
            for (uint32_t t = 0; t < rec.depth; ++t) {
                int p = (j->leaves[li].turn + (int)t) & 1;
                int u = rec.outcome == ORC_DRAW ? 0 : ((rec.outcome == ORC_WINNER_P0 ? 0 : 1) == p ? 1 : -1);
                orc_action_stat *e = &tab[(size_t)p * na + action_id(d->game, scratch[t])];
                __atomic_fetch_add(&e->visits, 1u, __ATOMIC_RELAXED);
                __atomic_fetch_add(&e->score, u, __ATOMIC_RELAXED);
            }
        }
        if (j->mast_delta) add_trace_stats(d->game, &j->leaves[li], scratch, rec.depth, rec.outcome, j->mast_delta);
        if (j->bigram_delta) { /* backprop.rs:885-899 over the playout part of the chronological list */
            const int S = orc_num_slots(d->game);
            for (uint32_t t = 0; t < rec.depth; ++t) {
                int p = (j->leaves[li].turn + (int)t) & 1;
                int pv = t > 0 ? (int)scratch[t - 1] : (j->leaves[li].prev_plus1 ? (int)j->leaves[li].prev_plus1 - 1 : -1);
                if (pv < 0) continue;
                int u = rec.outcome == ORC_DRAW ? 0 : ((rec.outcome == ORC_WINNER_P0 ? 0 : 1) == p ? 1 : -1);
                orc_action_stat *e = &j->bigram_delta[((size_t)p * S + (size_t)orc_action_slot(d->game, 1 - p, (uint16_t)pv)) * S +
                                                      (size_t)orc_action_slot(d->game, p, scratch[t])];
                e->visits += 1;
                e->score += u;
            }
        }
        if (j->reply_order || d->last_win_out) {
            /* backprop.rs:908-923: a player "won" when its utility is the trial's maximum: the winner, or both
             * players of a zero-utility ending */
            for (int p = 0; p < 2; ++p) {
                int lost = (rec.outcome == ORC_WINNER_P0 && p == 1) || (rec.outcome == ORC_WINNER_P1 && p == 0);
                if (!lost && d->last_win_out) {
                    uint32_t *lw = &d->last_win_out[(size_t)li * 2 + p], cur = __atomic_load_n(lw, __ATOMIC_RELAXED);
                    while (cur < pi + 1 && !__atomic_compare_exchange_n(lw, &cur, pi + 1, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                    }
                }
            }
            if (j->reply_order)
                for (uint32_t t = 0; t < rec.depth; ++t) {
                    int p = (j->leaves[li].turn + (int)t) & 1;
                    int lost = (rec.outcome == ORC_WINNER_P0 && p == 1) || (rec.outcome == ORC_WINNER_P1 && p == 0);
                    int pv = t > 0 ? (int)scratch[t - 1] : (j->leaves[li].prev_plus1 ? (int)j->leaves[li].prev_plus1 - 1 : -1);
                    if (lost || pv < 0) continue;
                    size_t key = (size_t)p * na + action_id(d->game, (uint16_t)pv);
                    uint64_t order = ((g + 1) << 13) | (4096u + t);
                    if (order > j->reply_order[key]) {
                        j->reply_order[key] = order;
                        j->reply_action[key] = (uint16_t)(scratch[t] + 1);
                    }
                }
        }
    }
    free(scratch);
    return NULL;
}

/* backprop.rs:908-966 applied LITERALLY, one trial after the other in global playout order, to the windows of the
 * chronological action list [leaf.prev2, leaf.prev, playout...] that end in a playout action. */
static void reply_tables_literal(const orc_batch_desc *d, const orc_state *leaves, uint32_t n_leaves, const orc_action_stat *mast) {
    const int game = d->game, na = orc_num_actions(game), S = orc_num_slots(game);
    const uint32_t k = d->playouts_per_leaf;
    uint16_t *t1 = d->replies_final_out, *t2 = d->replies2_final_out;
    if (t1) {
        if (d->replies) memcpy(t1, d->replies, sizeof(uint16_t) * 2 * (size_t)na);
        else memset(t1, 0, sizeof(uint16_t) * 2 * (size_t)na);
    }
    if (t2) {
        if (d->replies2) memcpy(t2, d->replies2, sizeof(uint16_t) * 2 * (size_t)S * S);
        else memset(t2, 0, sizeof(uint16_t) * 2 * (size_t)S * S);
    }
    orc_batch_desc local = *d;
    uint32_t cap = d->max_playout_depth ? d->max_playout_depth : 4096;
    local.max_trace = cap;
    uint16_t *chron = (uint16_t *)malloc(sizeof(uint16_t) * ((size_t)cap + 2));
    for (uint64_t g = 0; g < (uint64_t)n_leaves * k; ++g) {
        const orc_state *leaf = &leaves[g / k];
        orc_playout_record rec;
        orc_playout(&local, leaf, d->leaf_id_base + (uint32_t)(g / k), (uint32_t)(g % k), mast, chron + 2, &rec, NULL);
        /* chron[0] = leaf.prev2 (made by the leaf's mover), chron[1] = leaf.prev (by the other player), then the playout */
        const int have_prev = leaf->prev_plus1 != 0, have_prev2 = have_prev && leaf->prev2_plus1 != 0;
        if (have_prev) chron[1] = (uint16_t)(leaf->prev_plus1 - 1);
        if (have_prev2) chron[0] = (uint16_t)(leaf->prev2_plus1 - 1);
        const int first = have_prev2 ? 0 : (have_prev ? 1 : 2);
        const uint32_t depth = rec.depth < cap ? rec.depth : cap;
        for (uint32_t t = 0; t < depth; ++t) {
            const int i = (int)t + 2; /* position in chron */
            const int p = (leaf->turn + (int)t) & 1;
            const int lost = (rec.outcome == ORC_WINNER_P0 && p == 1) || (rec.outcome == ORC_WINNER_P1 && p == 0);
            if (t1 && i - 1 >= first && !lost) /* :916-923 */
                t1[(size_t)p * na + action_id(game, chron[i - 1])] = (uint16_t)(chron[i] + 1);
            if (t2 && i - 2 >= first) { /* :947-964 */
                uint16_t *e = &t2[((size_t)p * S + (size_t)orc_action_slot(game, p, chron[i - 2])) * S +
                                  (size_t)orc_action_slot(game, 1 - p, chron[i - 1])];
                if (!lost) *e = (uint16_t)(chron[i] + 1);
                else if (*e == (uint16_t)(chron[i] + 1)) *e = 0;
            }
        }
    }
    free(chron);
}

int orc_playout_batch(const orc_batch_desc *d, const orc_state *leaves, uint32_t n_leaves, const orc_action_stat *mast,
                      orc_leaf_result *results, orc_action_stat *amaf_out, orc_action_stat *mast_delta_out,
                      uint16_t *trace_out, orc_playout_record *rec_out, orc_state *final_out, int n_threads) {
    if (d->game < 0 || d->game >= ORC_NUM_GAMES) return -1;
    if ((d->policy == ORC_EPS_MAST || d->policy == ORC_DECISIVE_MAST_WIN || d->policy == ORC_DECISIVE_MAST_WINLOSS ||
         d->policy == ORC_EPS_NST || d->policy == ORC_EPS_LGR2_MAST) && !mast) return -2;
    /* trace-derived statistics need a bounded trace: only Knightthrough can run unboundedly long */
    if ((amaf_out || mast_delta_out || d->reply_order_out || d->bigram_delta_out) && d->max_playout_depth == 0 && d->game == ORC_KNIGHTTHROUGH) return -3;
    if (d->reply_order_out && !d->reply_action_out) return -4;
    if ((d->replies_final_out || d->replies2_final_out) && d->max_playout_depth == 0 && d->game == ORC_KNIGHTTHROUGH) return -3;
    if (d->replies_final_out || d->replies2_final_out) reply_tables_literal(d, leaves, n_leaves, mast);
    const int na = orc_num_actions(d->game);
    const uint64_t total = (uint64_t)n_leaves * d->playouts_per_leaf;
    memset(results, 0, sizeof(orc_leaf_result) * n_leaves);
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > total) n_threads = total ? (int)total : 1;
    batch_job *jobs = (batch_job *)calloc((size_t)n_threads, sizeof(batch_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        batch_job *j = &jobs[t];
        j->d = d;
        j->leaves = leaves;
        j->n_leaves = n_leaves;
        j->mast = mast;
        j->results = results;
        j->amaf_out = amaf_out;
        j->trace_out = trace_out;
        j->rec_out = rec_out;
        j->final_out = final_out;
        j->begin = total * (uint64_t)t / (uint64_t)n_threads;
        j->end = total * (uint64_t)(t + 1) / (uint64_t)n_threads;
        j->mast_delta = NULL;
        if (mast_delta_out) j->mast_delta = (orc_action_stat *)calloc((size_t)2 * na, sizeof(orc_action_stat));
        if (d->bigram_delta_out) {
            const size_t S = (size_t)orc_num_slots(d->game);
            j->bigram_delta = (orc_action_stat *)calloc(2 * S * S, sizeof(orc_action_stat));
        }
        if (d->reply_order_out) {
            j->reply_order = (uint64_t *)calloc((size_t)2 * na, sizeof(uint64_t));
            j->reply_action = (uint16_t *)calloc((size_t)2 * na, sizeof(uint16_t));
        }
    }
    if (d->last_win_out) memset(d->last_win_out, 0, sizeof(uint32_t) * 2 * n_leaves);
    if (n_threads == 1) {
        batch_worker(&jobs[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&tids[t], NULL, batch_worker, &jobs[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(tids[t], NULL);
    }
    if (mast_delta_out) {
        for (int t = 0; t < n_threads; ++t) {
            for (int i = 0; i < 2 * na; ++i) {
                mast_delta_out[i].visits += jobs[t].mast_delta[i].visits;
                mast_delta_out[i].score += jobs[t].mast_delta[i].score;
            }
            free(jobs[t].mast_delta);
        }
    }
    if (d->bigram_delta_out) {
        const size_t S = (size_t)orc_num_slots(d->game), cnt = 2 * S * S;
        for (int t = 0; t < n_threads; ++t) {
            for (size_t i = 0; i < cnt; ++i) {
                d->bigram_delta_out[i].visits += jobs[t].bigram_delta[i].visits;
                d->bigram_delta_out[i].score += jobs[t].bigram_delta[i].score;
            }
            free(jobs[t].bigram_delta);
        }
    }
    if (d->reply_order_out) {
        memset(d->reply_order_out, 0, sizeof(uint64_t) * 2 * na);
        memset(d->reply_action_out, 0, sizeof(uint16_t) * 2 * na);
        for (int t = 0; t < n_threads; ++t) {
            for (int i = 0; i < 2 * na; ++i)
                if (jobs[t].reply_order[i] > d->reply_order_out[i]) {
                    d->reply_order_out[i] = jobs[t].reply_order[i];
                    d->reply_action_out[i] = jobs[t].reply_action[i];
                }
            free(jobs[t].reply_order);
            free(jobs[t].reply_action);
        }
    }
    free(jobs);
    free(tids);
    return 0;
}

void orc_random_positions(int game, uint64_t seed, uint32_t plies, uint32_t n, orc_state *out) {
    /* position j = final state of the j-th SURVIVING playout (leaf_id 0, playout ids 0,1,2,...) of a
     * uniform playout from the initial position cut at `plies`; a playout that ends earlier is skipped.
     * The product's GpuRollout.random_positions uses the same definition, so both arms of the bench
     * see the same positions. */
    orc_batch_desc d;
    memset(&d, 0, sizeof d);
    d.game = game;
    d.policy = ORC_UNIFORM;
    d.max_playout_depth = plies;
    d.seed = seed;
    d.playouts_per_leaf = 1;
    orc_state init;
    orc_initial_state(game, &init);
    uint32_t got = 0;
    for (uint32_t pid = 0; got < n; ++pid) {
        if (plies == 0) {
            out[got++] = init;
            continue;
        }
        orc_playout_record rec;
        orc_state fin;
        orc_playout(&d, &init, 0, pid, NULL, NULL, &rec, &fin);
        if (rec.end_type == ORC_END_TURN_LIMIT && rec.depth == plies) out[got++] = fin;
    }
}

/* ---- operation counters of the counting build (bitboard.h) ---- */
#ifdef ORC_COUNT_OPS
uint64_t orc_ops[ORC_OP_KINDS];
#endif
int orc_ops_enabled(void) {
#ifdef ORC_COUNT_OPS
    return 1;
#else
    return 0;
#endif
}
void orc_ops_reset(void) {
#ifdef ORC_COUNT_OPS
    memset(orc_ops, 0, sizeof orc_ops);
#endif
}
void orc_ops_read(uint64_t *out /* [ORC_OP_KINDS] */) {
#ifdef ORC_COUNT_OPS
    memcpy(out, orc_ops, sizeof orc_ops);
#else
    memset(out, 0, sizeof(uint64_t) * ORC_OP_KINDS);
#endif
}
