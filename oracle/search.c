/*
 * oracle/search.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's MCTS loop around the playout path, used (a) as BASELINE.json
 * config 1 (Connect4 UCT, uniform rollouts, single thread, "rollouts/sec/core" as defined at
 * mcts/src/strategies/mcts/search/core.rs:395-398) and (b) to check the product's GPU-driven tree driver
 * root statistics bit for bit.
 *
 * Reference lines followed:
 *   iteration loop ........ search/search_impl.rs:76-114 (select, k rollouts of the leaf, k backprops)
 *   select_step ........... search/shared.rs:569-736; expand :220-289; virtual loss :665-671, :787-805
 *   UCB1 / UCB1-Tuned ..... select/ucb.rs:9-146; unvisited value node.rs:139-159; expected score node.rs:126-133
 *   random-offset arg-max . select/mod.rs:343-386 (strict '>', proven-loss skipping inert without the solver)
 *   backprop .............. backprop.rs:686-731; NodeStats/ChildArray update node.rs:295-302, 673-681
 *   GLOBAL (MAST) write ... backprop.rs:833-870 (playout trace + tree path)
 *   final action .......... RobustChild select/basic.rs:13-45 through core.rs:138-198
 *   LGRF-2 table .......... TreeStats.player_replies2, backprop.rs:926-966, applied LITERALLY: trial after trial in global
 *                           playout order over the whole chronological list (tree path, then playout), inserts for
 *                           players that did not lose, removals of still-matching entries for players that lost
 *   LGR table ............. TreeStats.player_replies, backprop.rs:908-923 (tree-path pairs and playout pairs of every
 *                           trial a player did not lose, last write wins)
 *   GRAVE ................. select/rave.rs:60-75 (beta), :157-184 (get_ref), :196-230 (score); tables written by
 *                           update_grave, backprop.rs:619-640, with amaf_actions grown per :833-854
 *
 * Batched mode (batch_leaves > 1) is the deterministic, sequential statement of what the tree-parallel
 * workers do concurrently (search/parallel.rs:188-251): B descents that each leave their virtual losses in
 * place, then all simulations, then all backprops -- with up to pipeline_depth batches gathered before the oldest
 * one is applied.  As in the product's driver the expansion test counts the trials still in flight through a node
 * (num_visits + in-flight < expand_threshold => simulate from here), so a node whose first simulation has not come
 * back is expanded by the next descent that reaches it instead of being queued again; with one leaf per batch
 * nothing is ever in flight and the loop is the reference's (search_impl.rs:76-114).
 * Tree tie-break draws: Philox4x32-10, ctr = (descent id, depth, 0, 2), word 0.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const uint32_t PRIMES[16] = {14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991,
                                    53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693};

typedef struct {
    uint32_t visits, vloss;
    int64_t score[2], sumsq[2];
    uint32_t amaf_visits; /* ChildArray amaf rows, node.rs:392-396,711-719 (all players' counts move together) */
    int64_t amaf_score[2];
} stats_t;

typedef struct {
    orc_state state;
    int status; /* 0 leaf (unexpanded), 1 expanded, 2 terminal */
    int player;
    int n_children;
    uint32_t first;
} node_t;

typedef struct {
    int game;
    node_t *nodes;
    size_t n_nodes, cap_nodes;
    uint16_t *action;
    int32_t *child;
    stats_t *edge;
    size_t n_edges, cap_edges;
    stats_t root_stats;
    uint64_t root_inflight; /* trials gathered through the root whose results have not been applied yet (the root has no
                               edge to carry virtual losses, and root_stats itself stays as the reference keeps it) */
    /* use_transpositions: open-addressing table state -> node id (table.rs keyed by zobrist hash in the
     * reference; keyed by the whole state here, which cannot collide) */
    int use_table;
    uint32_t *tab;
    size_t tab_cap, tab_used;
} tree_t;

static uint64_t state_hash(const orc_state *s) {
    uint64_t w[5], h = 0x9e3779b97f4a7c15ull;
    memcpy(w, s, sizeof w);
    for (int i = 0; i < 5; ++i) {
        h = (h ^ w[i]) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    return h;
}

typedef struct {
    uint32_t node, idx;
} entry_t;

/* TreeStats.grave (search/shared.rs:52): state -> [2][na] (visits, score).  The reference keys by node.hash;
 * the key here is the whole state. */
typedef struct {
    uint32_t visits;
    int64_t score;
} gstat_t;
typedef struct {
    orc_state *states;
    gstat_t **tabs;
    size_t n, cap;
    uint32_t *slots;
    size_t slot_cap;
    int na;
} grave_t;

static gstat_t *grave_find(grave_t *g, const orc_state *s, int create) {
    if (g->slot_cap == 0 || (g->n + 1) * 2 > g->slot_cap) {
        if (!create && g->slot_cap == 0) return NULL;
        if (create) {
            size_t ncap = g->slot_cap ? g->slot_cap * 2 : 4096;
            uint32_t *ns = (uint32_t *)malloc(ncap * sizeof(uint32_t));
            memset(ns, 0xff, ncap * sizeof(uint32_t));
            for (size_t i = 0; i < g->n; ++i) {
                size_t j = state_hash(&g->states[i]) & (ncap - 1);
                while (ns[j] != 0xffffffffu) j = (j + 1) & (ncap - 1);
                ns[j] = (uint32_t)i;
            }
            free(g->slots);
            g->slots = ns;
            g->slot_cap = ncap;
        }
    }
    size_t j = state_hash(s) & (g->slot_cap - 1);
    while (g->slots[j] != 0xffffffffu) {
        if (memcmp(&g->states[g->slots[j]], s, sizeof *s) == 0) return g->tabs[g->slots[j]];
        j = (j + 1) & (g->slot_cap - 1);
    }
    if (!create) return NULL;
    if (g->n == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 1024;
        g->states = (orc_state *)realloc(g->states, g->cap * sizeof(orc_state));
        g->tabs = (gstat_t **)realloc(g->tabs, g->cap * sizeof(gstat_t *));
    }
    g->states[g->n] = *s;
    g->tabs[g->n] = (gstat_t *)calloc((size_t)2 * g->na, sizeof(gstat_t));
    g->slots[j] = (uint32_t)g->n;
    return g->tabs[g->n++];
}

static double expected_score(const stats_t *s, int player) { /* node.rs:126-133 */
    if (s->visits == 0) return 0.0;
    double lv = (double)s->vloss;
    return ((double)s->score[player] - lv) / ((double)s->visits + lv);
}

static uint32_t tree_draw(uint64_t seed, uint32_t descent, uint32_t depth) {
    uint32_t ctr[4] = {descent, depth, 0, 2}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[0];
}

static uint32_t add_node(tree_t *t, const orc_state *s) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
        t->nodes = (node_t *)realloc(t->nodes, t->cap_nodes * sizeof(node_t));
    }
    node_t *n = &t->nodes[t->n_nodes];
    memset(n, 0, sizeof *n);
    n->state = *s;
    n->player = orc_player_to_move(t->game, s);
    return (uint32_t)t->n_nodes++;
}

/* returns the node holding `s`, creating (and registering) it when absent */
static uint32_t table_resolve(tree_t *t, const orc_state *s) {
    if ((t->tab_used + 1) * 2 > t->tab_cap) {
        size_t ncap = t->tab_cap ? t->tab_cap * 2 : 4096;
        uint32_t *nt = (uint32_t *)malloc(ncap * sizeof(uint32_t));
        memset(nt, 0xff, ncap * sizeof(uint32_t));
        for (size_t i = 0; i < t->tab_cap; ++i)
            if (t->tab[i] != 0xffffffffu) {
                size_t j = state_hash(&t->nodes[t->tab[i]].state) & (ncap - 1);
                while (nt[j] != 0xffffffffu) j = (j + 1) & (ncap - 1);
                nt[j] = t->tab[i];
            }
        free(t->tab);
        t->tab = nt;
        t->tab_cap = ncap;
    }
    size_t j = state_hash(s) & (t->tab_cap - 1);
    while (t->tab[j] != 0xffffffffu) {
        if (memcmp(&t->nodes[t->tab[j]].state, s, sizeof *s) == 0) return t->tab[j];
        j = (j + 1) & (t->tab_cap - 1);
    }
    uint32_t id = add_node(t, s);
    t->tab[j] = id;
    t->tab_used++;
    return id;
}

static void expand(tree_t *t, uint32_t id) { /* shared.rs:220-289 */
    node_t *n = &t->nodes[id];
    if (orc_terminal_status(t->game, &n->state) != ORC_NOT_TERMINAL) {
        n->status = 2;
        return;
    }
    uint16_t acts[ORC_MAX_ACTIONS];
    int cnt = orc_generate_actions(t->game, &n->state, acts);
    if (cnt == 0) { /* the reference debug_asserts this away; a stuck position is played out as a draw */
        n->status = 2;
        return;
    }
    if (t->n_edges + (size_t)cnt > t->cap_edges) {
        while (t->n_edges + (size_t)cnt > t->cap_edges) t->cap_edges = t->cap_edges ? t->cap_edges * 2 : 4096;
        t->action = (uint16_t *)realloc(t->action, t->cap_edges * sizeof(uint16_t));
        t->child = (int32_t *)realloc(t->child, t->cap_edges * sizeof(int32_t));
        t->edge = (stats_t *)realloc(t->edge, t->cap_edges * sizeof(stats_t));
    }
    n->first = (uint32_t)t->n_edges;
    n->n_children = cnt;
    for (int i = 0; i < cnt; ++i) {
        t->action[t->n_edges] = acts[i];
        t->child[t->n_edges] = -1;
        memset(&t->edge[t->n_edges], 0, sizeof(stats_t));
        t->n_edges++;
    }
    n->status = 1;
}

static const stats_t *current_stats(const tree_t *t, const entry_t *path, int len) {
    if (len == 1) return &t->root_stats;
    return &t->edge[t->nodes[path[len - 2].node].first + path[len - 1].idx];
}

typedef struct {
    int select; /* 0 UCB1, 1 UCB1-Tuned, 2 GRAVE, 3 AMAF */
    int q_init; /* 0 Parent, 1 Infinity */
    double c;
    uint32_t expand_threshold;
    uint64_t seed;
    uint32_t grave_threshold;
    int rave_schedule;
    uint32_t rave_param;
    grave_t *grave;
    double amaf_alpha;
    double rave_bias;
} sel_cfg;

static int action_id(int game, uint16_t a);

/* rave.rs:157-184 get_ref: nearest node on the path, the child first, whose edge has total_visits >= threshold */
static uint32_t grave_ref(const tree_t *t, const sel_cfg *cfg, const entry_t *path, int len, uint32_t child_node, uint32_t child_idx) {
    const node_t *cur = &t->nodes[path[len - 1].node];
    const stats_t *ce = &t->edge[cur->first + child_idx];
    if (ce->visits + ce->vloss >= cfg->grave_threshold) return child_node;
    for (int i = len - 1; i >= 1; --i) {
        const stats_t *e = &t->edge[t->nodes[path[i - 1].node].first + path[i].idx];
        if (e->visits + e->vloss >= cfg->grave_threshold) return path[i].node;
    }
    return 0;
}

/* value of a child that does not exist yet: node.rs:139-159 value_estimate_unvisited plus the strategy's exploration
 * term (ucb.rs:55-61, :134-146; none for Rave / Amaf, rave.rs:226-230, amaf.rs:76-79) */
static double unvisited_score(const sel_cfg *cfg, const stats_t *cur, int player, double parent_log) {
    double q_unvisited = cfg->q_init == 1 ? 10000.0 : (cur->visits == 0 ? 10000.0 : expected_score(cur, player));
    return cfg->select == 1   ? q_unvisited + (parent_log * fmin(1.0, 1.0) + cfg->c * sqrt(parent_log))
           : (cfg->select == 2 || cfg->select == 3) ? q_unvisited
                              : q_unvisited + cfg->c * sqrt(parent_log);
}

/* score of a created child; amaf_n / amaf_q = the GRAVE reference node's entry for (mover, action) */
static double created_score(const sel_cfg *cfg, const stats_t *e, int player, double parent_log, uint32_t amaf_n, double amaf_q) {
    double exploit = expected_score(e, player);
    double total = (double)(e->visits + e->vloss);
    if (cfg->select == 1) { /* ucb.rs:90-132 */
        double var = fmax(0.0, (double)e->sumsq[player] / total - exploit * exploit);
        double frac = parent_log / total;
        return exploit + (frac * fmin(1.0, var) + cfg->c * sqrt(frac));
    }
    if (cfg->select == 2) { /* rave.rs:196-224 */
        uint32_t n_tot = e->visits + e->vloss;
        double explore = cfg->c * sqrt(parent_log / (double)n_tot);
        double b;
        if (cfg->rave_schedule == 1) {
            double rave = (double)cfg->rave_param;
            b = fmax(0.0, rave - (double)n_tot) / rave;
        } else if (cfg->rave_schedule == 2) { /* MinMSE { bias }, rave.rs:68 */
            double nn = (double)n_tot, an = (double)amaf_n;
            b = an / (nn + an + 4. * nn * an * cfg->rave_bias * cfg->rave_bias);
        } else {
            double kk = (double)cfg->rave_param;
            b = sqrt(kk / (3.0 * (double)n_tot + kk));
        }
        double amaf = amaf_n == 0 ? 0.0 : amaf_q / (double)amaf_n;
        return (1.0 - b) * exploit + b * amaf + explore;
    }
    if (cfg->select == 3) { /* amaf.rs:58-74 */
        double an = (double)(e->amaf_visits > 1u ? e->amaf_visits : 1u);
        double amaf = (double)e->amaf_score[player] / an;
        double ucb1 = exploit + cfg->c * sqrt(parent_log / total);
        return cfg->amaf_alpha * amaf + (1.0 - cfg->amaf_alpha) * ucb1;
    }
    return exploit + cfg->c * sqrt(parent_log / total); /* ucb.rs:40-52 */
}

static uint32_t best_child(const tree_t *t, const sel_cfg *cfg, const entry_t *path, int len, uint32_t descent) {
    const node_t *node = &t->nodes[path[len - 1].node];
    int player = node->player;
    const stats_t *cur = current_stats(t, path, len);
    double parent_log = log(fmax(1.0, (double)cur->visits)); /* ucb.rs:34-37 */
    double unvisited_value = unvisited_score(cfg, cur, player, parent_log);
    uint32_t n = (uint32_t)node->n_children;
    uint32_t r = (uint32_t)(((uint64_t)tree_draw(cfg->seed, descent, (uint32_t)len - 1) * (16u * n)) >> 32);
    uint32_t i = r / 16u, stride = PRIMES[r % 16u];
    int have = 0;
    double best = 0.0;
    uint32_t best_i = i;
    for (uint32_t k = 0; k < n; ++k) {
        const stats_t *e = &t->edge[node->first + i];
        double score;
        if (t->child[node->first + i] < 0) {
            score = unvisited_value;
        } else {
            uint32_t amaf_n = 0;
            double amaf_q = 0.0;
            if (cfg->select == 2) {
                uint32_t ref = grave_ref(t, cfg, path, len, (uint32_t)t->child[node->first + i], i);
                gstat_t *tab = grave_find(cfg->grave, &t->nodes[ref].state, 0);
                if (tab) {
                    const gstat_t *g = &tab[(size_t)player * cfg->grave->na + action_id(t->game, t->action[node->first + i])];
                    amaf_n = g->visits;
                    amaf_q = (double)g->score;
                }
            }
            score = created_score(cfg, e, player, parent_log, amaf_n, amaf_q);
        }
        if (!have || score > best) {
            have = 1;
            best = score;
            best_i = i;
        }
        i = (i + stride) % n;
    }
    return best_i;
}

/* ---- canonical-symmetry transposition keys (Othello): games/othello/src/lib.rs:26, 425-480, 618-665 with the D4 index
 * maps of games/game-core/src/symmetry.rs:88-147, restated square by square (the product's host rules use word-parallel
 * flips, so the two are independent implementations) ---- */
#define SYMMETRY_PLY_LIMIT 20 /* othello/lib.rs:26: discs on the board up to which states are canonicalised */

static int sym_flip_cols(int i) { return (i / 8) * 8 + (7 - i % 8); }
static int sym_flip_rows(int i) { return (7 - i / 8) * 8 + i % 8; }
static int sym_transpose(int i) { return (i % 8) * 8 + i / 8; }

/* symmetry.rs:113-127 index_symmetries(i)[k]: identity, flip_cols, flip_rows, transpose, flip_rows o flip_cols,
 * transpose o flip_cols, transpose o flip_rows, transpose o flip_rows o flip_cols */
static int sym_index(int i, int k) {
    switch (k) {
    case 0: return i;
    case 1: return sym_flip_cols(i);
    case 2: return sym_flip_rows(i);
    case 3: return sym_transpose(i);
    case 4: return sym_flip_rows(sym_flip_cols(i));
    case 5: return sym_transpose(sym_flip_cols(i));
    case 6: return sym_transpose(sym_flip_rows(i));
    default: return sym_transpose(sym_flip_rows(sym_flip_cols(i)));
    }
}

/* symmetry.rs:134-147 invert_symmetry */
static int sym_invert(int i, int k) {
    switch (k) {
    case 0: return i;
    case 1: return sym_flip_cols(i);
    case 2: return sym_flip_rows(i);
    case 3: return sym_transpose(i);
    case 4: return sym_flip_cols(sym_flip_rows(i));
    case 5: return sym_flip_cols(sym_transpose(i));
    case 6: return sym_flip_rows(sym_transpose(i));
    default: return sym_flip_cols(sym_flip_rows(sym_transpose(i)));
    }
}

/* D4Symmetry::apply_to_bits: every set square moves to its image */
static uint64_t sym_bits(uint64_t b, int k) {
    uint64_t out = 0;
    for (int i = 0; i < 64; ++i)
        if ((b >> i) & 1ull) out |= 1ull << sym_index(i, k);
    return out;
}

/* othello/lib.rs:436-443 canonical_symmetry: the FIRST symmetry whose image of (black, white) is lexicographically minimal */
static int othello_canonical_symmetry(uint64_t black, uint64_t white) {
    int best = 0;
    uint64_t bb = black, bw = white;
    for (int k = 1; k < 8; ++k) {
        uint64_t b = sym_bits(black, k), w = sym_bits(white, k);
        if (b < bb || (b == bb && w < bw)) {
            best = k;
            bb = b;
            bw = w;
        }
    }
    return best;
}

/* othello/lib.rs:618-646 canonical_representation: past the disc limit the state itself with the identity; else its
 * canonical image and the symmetry that produced it */
static int othello_canonical(const orc_state *s, orc_state *canon) {
    *canon = *s;
    if (__builtin_popcountll(s->board[0] | s->board[1]) > SYMMETRY_PLY_LIMIT) return 0;
    int k = othello_canonical_symmetry(s->board[0], s->board[1]);
    canon->board[0] = sym_bits(s->board[0], k);
    canon->board[1] = sym_bits(s->board[1], k);
    return k;
}

/* test hook: Game::canonical_representation (identity for every game but Othello); returns the symmetry index */
int orc_canonical_representation(int game, const orc_state *s, orc_state *canon) {
    if (game == ORC_OTHELLO) return othello_canonical(s, canon);
    *canon = *s;
    return 0;
}
/* test hooks: symmetry.rs:113-127 index_symmetries(i)[k] and :134-147 invert_symmetry(i, k) */
int orc_symmetry_index(int i, int k) { return sym_index(i, k); }
int orc_symmetry_invert(int i, int k) { return sym_invert(i, k); }

/* shared.rs:569-736; returns the path length.  With `symmetry` (use_transpositions on a game with canonical keys) the
 * descent carries the REAL state: a node's action list is in its canonical orientation (shared.rs:233-252) and every
 * chosen action is translated back through the symmetry of the real state at that node (symmetry.rs:38-50 incoming_sym,
 * node.rs real_action, othello/lib.rs:658-665 invert_action); the root's list is literal.  `leaf_state` receives the real
 * state the playouts start from. */
static int select_step(tree_t *t, const sel_cfg *cfg, uint32_t descent, entry_t *path, int symmetry, orc_state *leaf_state) {
    orc_state real = t->nodes[0].state;
    int real_sym = 0;
    int len = 0;
    uint32_t cur = 0, incoming = 0;
    for (;;) {
        path[len].node = cur;
        path[len].idx = incoming;
        ++len;
        const stats_t *cs = current_stats(t, path, len);
        /* trials in flight through this node from other descents: k virtual losses per gathered descent on each of its
         * edges (root_inflight for the root); this descent has put 1 on the edge it just came down */
        uint64_t in_flight = len == 1 ? t->root_inflight : cs->vloss - 1u;
        if (leaf_state) *leaf_state = symmetry ? real : t->nodes[cur].state; /* (for every return below) */
        if (t->nodes[cur].status == 2) return len;
        if ((uint64_t)cs->visits + in_flight < cfg->expand_threshold) return len;
        if (t->nodes[cur].status == 0) {
            expand(t, cur);
            if (t->nodes[cur].status == 2) return len;
        }
        uint32_t best = best_child(t, cfg, path, len, descent);
        incoming = best;
        uint32_t slot = t->nodes[cur].first + best;
        t->edge[slot].vloss += 1;
        orc_state key; /* what the child node is keyed by */
        if (symmetry) {
            uint16_t a = t->action[slot];
            /* PASS maps to itself (othello/lib.rs:650-652); real_sym = canonical_representation(real).1 at this node */
            if (cur != 0 && a < 64) a = (uint16_t)sym_invert((int)a, real_sym);
            orc_apply(t->game, &real, a);
            real_sym = othello_canonical(&real, &key);
        }
        if (t->child[slot] >= 0) {
            cur = (uint32_t)t->child[slot];
            /* shared.rs:382-405 verified_child_id: the cached slot belongs to whichever orientation created it; past the
             * disc limit another orientation's child is a different node, found (or made) through the table */
            if (symmetry && memcmp(&t->nodes[cur].state, &key, sizeof key) != 0) cur = table_resolve(t, &key);
        } else {
            orc_state s = t->nodes[cur].state;
            if (symmetry) s = key;
            else orc_apply(t->game, &s, t->action[slot]);
            uint32_t id = t->use_table ? table_resolve(t, &s) : add_node(t, &s);
            t->child[slot] = (int32_t)id;
            cur = id;
            if (cfg->expand_threshold > 0) {
                path[len].node = cur;
                path[len].idx = incoming;
                ++len;
                if (leaf_state) *leaf_state = symmetry ? real : t->nodes[cur].state;
                return len;
            }
        }
    }
}

static int action_id(int game, uint16_t a) {
    if (game == ORC_BREAKTHROUGH || game == ORC_KNIGHTTHROUGH) return (int)(a >> 8) * 64 + (int)(a & 0xff);
    return (int)a;
}

/* update_amaf (backprop.rs:565-617) for the k trials of one leaf, for every non-root path node X, leaf first:
 * occurrences of (action, mover of X's parent) in [playout trace + path edges below X] that name a created child of the
 * parent add one AMAF visit and the trial's utilities to that child's row.  `la` = the leaf's occurrence table [2][na]. */
static void update_amaf_path(tree_t *t, const entry_t *path, int len, const orc_action_stat *la, int na, const int64_t u[2],
                             uint32_t k, uint32_t *below_a, int *below_p) {
    int n_below = 0;
    for (int i = len - 1; i >= 1; --i) {
        const node_t *parent = &t->nodes[path[i - 1].node];
        int mover = parent->player;
        for (int j = 0; j < parent->n_children; ++j) {
            if (t->child[parent->first + j] < 0) continue;
            uint32_t a = (uint32_t)action_id(t->game, t->action[parent->first + j]);
            uint32_t cnt = la[(size_t)mover * na + a].visits;
            int64_t sc = la[(size_t)mover * na + a].score;
            for (int q = 0; q < n_below; ++q)
                if (below_p[q] == mover && below_a[q] == a) {
                    cnt += k;
                    sc += u[mover];
                }
            stats_t *e = &t->edge[parent->first + j];
            e->amaf_visits += cnt;
            e->amaf_score[mover] += sc;
            e->amaf_score[1 - mover] -= sc;
        }
        below_a[n_below] = (uint32_t)action_id(t->game, t->action[parent->first + path[i].idx]);
        below_p[n_below] = mover;
        ++n_below;
    }
}

int orc_search(const orc_search_cfg *c, const orc_state *root, uint16_t *best_action, orc_root_child *children,
               uint32_t *n_children, uint64_t *playout_plies) {
    tree_t t;
    memset(&t, 0, sizeof t);
    t.game = c->game;
    t.use_table = c->use_transpositions != 0;
    /* use_transpositions 2: the table is keyed as the reference keys it for Othello, by the canonical-symmetry image up to
     * SYMMETRY_PLY_LIMIT discs (othello/lib.rs:474-480) */
    const int symmetry = (c->use_transpositions & 2) != 0 && c->game == ORC_OTHELLO;
    if (t.use_table) table_resolve(&t, root);
    else add_node(&t, root);
    grave_t grave;
    memset(&grave, 0, sizeof grave);
    grave.na = orc_num_actions(c->game);
    sel_cfg cfg = {c->select, c->q_init, c->exploration_c, c->expand_threshold, c->seed, c->grave_threshold, c->rave_schedule,
                   c->rave_param, &grave, c->amaf_alpha, c->rave_bias};
    int depth = c->pipeline_depth < 1 ? 1 : (c->pipeline_depth > ORC_MAX_PIPELINE ? ORC_MAX_PIPELINE : c->pipeline_depth);
    const int want_amaf = c->select == 2 || c->select == 3;
    const uint32_t B = c->batch_leaves, k = c->playouts_per_leaf;
    const int na = orc_num_actions(c->game);
    const int nst = c->policy == ORC_EPS_NST;
    orc_action_stat *mast = (c->policy == ORC_EPS_MAST || c->policy == ORC_DECISIVE_MAST_WIN || c->policy == ORC_DECISIVE_MAST_WINLOSS || nst || c->policy == ORC_EPS_LGR2_MAST) ? (orc_action_stat *)calloc((size_t)2 * na, sizeof(orc_action_stat)) : NULL;
    /* TreeStats.player_bigram_actions (shared.rs:60-66) as [2][S][S] over compact action slots, and one delta per buffer set */
    const size_t S = (size_t)orc_num_slots(c->game);
    orc_action_stat *bigram = nst ? (orc_action_stat *)calloc(2 * S * S, sizeof(orc_action_stat)) : NULL;
    orc_action_stat *bg_delta[ORC_MAX_PIPELINE] = {NULL};
    for (int q = 0; q < depth && nst; ++q) bg_delta[q] = (orc_action_stat *)calloc(2 * S * S, sizeof(orc_action_stat));
    const int lgr2 = c->policy == ORC_EPS_LGR2 || c->policy == ORC_EPS_LGR2_MAST;
    const int lgr = c->policy == ORC_EPS_LGR || lgr2; /* LGRF-2 keeps the LGR-1 table too (its inner strategy) */
    uint16_t *replies2 = lgr2 ? (uint16_t *)calloc(2 * S * S, sizeof(uint16_t)) : NULL;
    const uint32_t trace_cap = c->max_playout_depth ? c->max_playout_depth : 256;
    uint16_t *trace2[ORC_MAX_PIPELINE] = {NULL};
    orc_playout_record *rec2[ORC_MAX_PIPELINE] = {NULL};
    for (int q = 0; q < depth && lgr2; ++q) {
        trace2[q] = (uint16_t *)calloc((size_t)c->batch_leaves * c->playouts_per_leaf * trace_cap, sizeof(uint16_t));
        rec2[q] = (orc_playout_record *)calloc((size_t)c->batch_leaves * c->playouts_per_leaf, sizeof(orc_playout_record));
    }
    uint16_t *replies = lgr ? (uint16_t *)calloc((size_t)2 * na, sizeof(uint16_t)) : NULL;
    uint64_t *upd_order[ORC_MAX_PIPELINE] = {NULL};
    uint16_t *upd_action[ORC_MAX_PIPELINE] = {NULL};
    uint32_t *last_win[ORC_MAX_PIPELINE] = {NULL};
    for (int q = 0; q < depth && lgr; ++q) {
        upd_order[q] = (uint64_t *)calloc((size_t)2 * na, sizeof(uint64_t));
        upd_action[q] = (uint16_t *)calloc((size_t)2 * na, sizeof(uint16_t));
        last_win[q] = (uint32_t *)calloc((size_t)2 * c->batch_leaves, sizeof(uint32_t));
    }
    const int max_path = 1024;
    orc_state *leaves = (orc_state *)malloc(sizeof(orc_state) * B);
    uint32_t *below_a = (uint32_t *)malloc(sizeof(uint32_t) * 1024);
    int *below_p = (int *)malloc(sizeof(int) * 1024);
    uint64_t plies = 0;

    orc_batch_desc bd;
    memset(&bd, 0, sizeof bd);
    bd.game = c->game;
    bd.policy = c->policy;
    bd.eps_threshold = c->eps_threshold;
    bd.max_playout_depth = c->max_playout_depth;
    bd.seed = c->seed;
    bd.playouts_per_leaf = k;

    /* One buffer set per batch in flight: batch n is gathered and simulated into set n % depth before the results of
     * batches n - depth + 1 .. n - 1 are applied -- the fixed order gather(n), simulate(n), backprop(n - depth + 1). */
    entry_t *set_paths[ORC_MAX_PIPELINE] = {NULL};
    int *set_plen[ORC_MAX_PIPELINE] = {NULL};
    orc_leaf_result *set_res[ORC_MAX_PIPELINE] = {NULL};
    orc_action_stat *set_amaf[ORC_MAX_PIPELINE] = {NULL};
    orc_action_stat *set_delta[ORC_MAX_PIPELINE] = {NULL};
    uint32_t set_nb[ORC_MAX_PIPELINE] = {0};
    for (int q = 0; q < depth; ++q) {
        set_paths[q] = (entry_t *)malloc(sizeof(entry_t) * (size_t)B * max_path);
        set_plen[q] = (int *)malloc(sizeof(int) * B);
        set_res[q] = (orc_leaf_result *)malloc(sizeof(orc_leaf_result) * B);
        set_amaf[q] = want_amaf ? (orc_action_stat *)malloc(sizeof(orc_action_stat) * (size_t)B * 2 * na) : NULL;
        set_delta[q] = mast ? (orc_action_stat *)calloc((size_t)2 * na, sizeof(orc_action_stat)) : NULL;
    }

    uint32_t done = 0, submitted = 0, applied = 0;
    for (;;) {
        int slot = -1;
        if (done < c->max_iterations) {
            /* ---- gather + simulate one batch into `slot` ---- */
            slot = (int)(submitted % (uint32_t)depth);
            ++submitted;
            uint32_t nb = B < c->max_iterations - done ? B : c->max_iterations - done;
            set_nb[slot] = nb;
            for (uint32_t b = 0; b < nb; ++b) {
                entry_t *path = set_paths[slot] + (size_t)b * max_path;
                set_plen[slot][b] = select_step(&t, &cfg, done + b, path, symmetry, &leaves[b]);
                for (int i = 1; i < set_plen[slot][b]; ++i) /* add_path_virtual_loss(k - 1), shared.rs:787-805 */
                    t.edge[t.nodes[path[i - 1].node].first + path[i].idx].vloss += k - 1;
                t.root_inflight += k; /* the k trials now in flight through the root */
                if ((lgr || nst) && set_plen[slot][b] > 1) { /* playout()'s prev_action: the last edge of the descent */
                    int L = set_plen[slot][b];
                    leaves[b].prev_plus1 = (uint16_t)(t.action[t.nodes[path[L - 2].node].first + path[L - 1].idx] + 1);
                    if (lgr2 && L > 2) /* and the edge before it, for the kernel-side contract (unused by this literal mirror) */
                        leaves[b].prev2_plus1 = (uint16_t)(t.action[t.nodes[path[L - 3].node].first + path[L - 2].idx] + 1);
                }
            }
            bd.replies = replies; /* frozen for the batch: the playouts run before anything below is merged */
            bd.reply_order_out = lgr ? upd_order[slot] : NULL;
            bd.reply_action_out = lgr ? upd_action[slot] : NULL;
            bd.last_win_out = lgr ? last_win[slot] : NULL;
            bd.replies2 = replies2; /* frozen for the batch */
            bd.max_trace = lgr2 ? trace_cap : 0;
            bd.bigram = bigram; /* frozen for the batch, like the MAST table */
            bd.nst_threshold = c->nst_threshold;
            bd.bigram_delta_out = nst ? bg_delta[slot] : NULL;
            if (nst) memset(bg_delta[slot], 0, sizeof(orc_action_stat) * 2 * S * S);
            bd.leaf_id_base = done;
            if (set_delta[slot]) memset(set_delta[slot], 0, sizeof(orc_action_stat) * 2 * (size_t)na);
            if (set_amaf[slot]) memset(set_amaf[slot], 0, sizeof(orc_action_stat) * (size_t)nb * 2 * na);
            orc_playout_batch(&bd, leaves, nb, mast, set_res[slot], set_amaf[slot], set_delta[slot], lgr2 ? trace2[slot] : NULL,
                              lgr2 ? rec2[slot] : NULL, NULL, c->threads > 0 ? c->threads : 1);
            done += nb;
        }
        /* ---- which batch gets applied now: the oldest one, once `depth` batches are in flight (or nothing is left to gather) ---- */
        if (slot >= 0 && submitted - applied < (uint32_t)depth) continue;
        if (applied == submitted) break;
        const int apply = (int)(applied % (uint32_t)depth);
        ++applied;
        if (lgr) { /* backprop.rs:908-923 over the whole batch: tree-path pairs join the playout stamps, then one write per key */
            for (uint32_t b = 0; b < set_nb[apply]; ++b) {
                const entry_t *path = set_paths[apply] + (size_t)b * max_path;
                for (int i = 2; i < set_plen[apply][b]; ++i) {
                    const node_t *grand = &t.nodes[path[i - 2].node], *parent = &t.nodes[path[i - 1].node];
                    int p = parent->player;
                    uint32_t lw = last_win[apply][(size_t)b * 2 + p];
                    if (lw == 0) continue;
                    uint16_t prev = t.action[grand->first + path[i - 1].idx], act = t.action[parent->first + path[i].idx];
                    uint64_t order = ((((uint64_t)b * k + (lw - 1)) + 1) << 13) | (uint64_t)(i < 4095 ? i : 4095);
                    size_t key = (size_t)p * na + action_id(c->game, prev);
                    if (order > upd_order[apply][key]) {
                        upd_order[apply][key] = order;
                        upd_action[apply][key] = (uint16_t)(act + 1);
                    }
                }
            }
            for (int i = 0; i < 2 * na; ++i)
                if (upd_order[apply][i]) replies[i] = upd_action[apply][i];
        }
        if (lgr2) { /* backprop.rs:926-966, literally: every trial of the batch in order, tree path then playout */
            uint16_t *chron = (uint16_t *)malloc(sizeof(uint16_t) * ((size_t)max_path + trace_cap));
            int *who = (int *)malloc(sizeof(int) * ((size_t)max_path + trace_cap));
            for (uint32_t b = 0; b < set_nb[apply]; ++b) {
                const entry_t *path = set_paths[apply] + (size_t)b * max_path;
                const int L = set_plen[apply][b];
                int n_tree = 0;
                for (int i = 1; i < L; ++i) {
                    const node_t *parent = &t.nodes[path[i - 1].node];
                    chron[n_tree] = t.action[parent->first + path[i].idx];
                    who[n_tree++] = parent->player;
                }
                const int leaf_turn = t.nodes[path[L - 1].node].player;
                for (uint32_t j = 0; j < k; ++j) {
                    const orc_playout_record *rc = &rec2[apply][(size_t)b * k + j];
                    const uint16_t *tr = trace2[apply] + ((size_t)b * k + j) * trace_cap;
                    const uint32_t d = rc->depth < trace_cap ? rc->depth : trace_cap;
                    int n = n_tree;
                    for (uint32_t q = 0; q < d; ++q) {
                        chron[n] = tr[q];
                        who[n++] = (leaf_turn + (int)q) & 1;
                    }
                    for (int i = 2; i < n; ++i) {
                        const int p = who[i];
                        if (who[i - 2] != p) continue;
                        const int lost = (rc->outcome == ORC_WINNER_P0 && p == 1) || (rc->outcome == ORC_WINNER_P1 && p == 0);
                        uint16_t *e = &replies2[((size_t)p * S + (size_t)orc_action_slot(c->game, p, chron[i - 2])) * S +
                                                (size_t)orc_action_slot(c->game, 1 - p, chron[i - 1])];
                        if (!lost) *e = (uint16_t)(chron[i] + 1);
                        else if (*e == (uint16_t)(chron[i] + 1)) *e = 0;
                    }
                }
            }
            free(chron);
            free(who);
        }
        for (uint32_t b = 0; b < set_nb[apply]; ++b) {
            const entry_t *path = set_paths[apply] + (size_t)b * max_path;
            const int *plen_a = set_plen[apply];
            const orc_leaf_result *r = &set_res[apply][b];
            int64_t nonzero = (int64_t)k - r->draws;
            int64_t u[2] = {(int64_t)r->wins[0] - (nonzero - r->wins[0]), (int64_t)r->wins[1] - (nonzero - r->wins[1])};
            if (set_amaf[apply] && c->select == 3)
                update_amaf_path(&t, path, plen_a[b], set_amaf[apply] + (size_t)b * 2 * na, na, u, k, below_a, below_p);
            if (set_amaf[apply] && c->select == 2) { /* update_grave for every non-root path node, leaf first (backprop.rs:619-640, 833-854) */
                const orc_action_stat *la = set_amaf[apply] + (size_t)b * 2 * na;
                int n_below = 0;
                for (int i = plen_a[b] - 1; i >= 1; --i) {
                    gstat_t *tab = grave_find(&grave, &t.nodes[path[i].node].state, 1);
                    for (int e = 0; e < 2 * na; ++e) {
                        tab[e].visits += la[e].visits;
                        tab[e].score += la[e].score;
                    }
                    for (int q = 0; q < n_below; ++q) {
                        gstat_t *g = &tab[(size_t)below_p[q] * na + below_a[q]];
                        g->visits += k;
                        g->score += u[below_p[q]];
                    }
                    const node_t *parent = &t.nodes[path[i - 1].node];
                    below_a[n_below] = (uint32_t)action_id(c->game, t.action[parent->first + path[i].idx]);
                    below_p[n_below] = parent->player;
                    ++n_below;
                }
            }
            for (int i = plen_a[b] - 1; i >= 0; --i) {
                stats_t *s = i == 0 ? &t.root_stats : &t.edge[t.nodes[path[i - 1].node].first + path[i].idx];
                s->visits += k;
                for (int p = 0; p < 2; ++p) {
                    s->score[p] += u[p];
                    s->sumsq[p] += nonzero;
                }
                if (i == 0) t.root_inflight -= k;
                if (i > 0) {
                    s->vloss -= k;
                    if (mast) { /* tree-path part of the GLOBAL write */
                        const node_t *parent = &t.nodes[path[i - 1].node];
                        orc_action_stat *m = &mast[(size_t)parent->player * na + action_id(c->game, t.action[parent->first + path[i].idx])];
                        m->visits += k;
                        m->score += (int32_t)u[parent->player];
                    }
                }
            }
            if (nst) { /* backprop.rs:885-899, tree-path part: consecutive edges of the descent, once per trial of the leaf */
                for (int i = 2; i < plen_a[b]; ++i) {
                    const node_t *grand = &t.nodes[path[i - 2].node], *parent = &t.nodes[path[i - 1].node];
                    int p = parent->player;
                    uint16_t prev = t.action[grand->first + path[i - 1].idx], act = t.action[parent->first + path[i].idx];
                    orc_action_stat *e = &bigram[((size_t)p * S + (size_t)orc_action_slot(c->game, 1 - p, prev)) * S +
                                                 (size_t)orc_action_slot(c->game, p, act)];
                    e->visits += k;
                    e->score += (int32_t)u[p];
                }
            }
            plies += r->sum_depth;
        }
        if (nst) /* the playout part: (last edge, first playout action) and every later consecutive pair */
            for (size_t i = 0; i < 2 * S * S; ++i) {
                bigram[i].visits += bg_delta[apply][i].visits;
                bigram[i].score += bg_delta[apply][i].score;
            }
        if (mast)
            for (int i = 0; i < 2 * na; ++i) {
                mast[i].visits += set_delta[apply][i].visits;
                mast[i].score += set_delta[apply][i].score;
            }
    }

    const node_t *rootn = &t.nodes[0];
    int rc = 0;
    if (rootn->status != 1 || rootn->n_children == 0) {
        *best_action = 0xffff;
        if (n_children) *n_children = 0;
    } else {
        uint32_t n = (uint32_t)rootn->n_children;
        int player = rootn->player;
        double q_unvisited = c->q_init == 1 ? 10000.0 : (t.root_stats.visits == 0 ? 10000.0 : expected_score(&t.root_stats, player));
        uint32_t r = (uint32_t)(((uint64_t)tree_draw(c->seed, 0xffffffffu, 0) * (16u * n)) >> 32);
        uint32_t i = r / 16u, stride = PRIMES[r % 16u], best_i = i;
        int have = 0;
        int64_t best_v = 0;
        double best_q = 0.0;
        for (uint32_t q = 0; q < n; ++q) { /* RobustChild, basic.rs:13-45 */
            const stats_t *e = &t.edge[rootn->first + i];
            int created = t.child[rootn->first + i] >= 0;
            int64_t v = created ? (int64_t)e->visits : 0;
            double qq = created ? expected_score(e, player) : q_unvisited;
            if (!have || v > best_v || (v == best_v && qq > best_q)) {
                have = 1;
                best_v = v;
                best_q = qq;
                best_i = i;
            }
            i = (i + stride) % n;
        }
        *best_action = t.action[rootn->first + best_i];
        if (children && n_children) {
            *n_children = n;
            for (uint32_t q = 0; q < n; ++q) {
                const stats_t *e = &t.edge[rootn->first + q];
                children[q].action = t.action[rootn->first + q];
                children[q].pad = 0;
                children[q].visits = e->visits;
                children[q].score[0] = e->score[0];
                children[q].score[1] = e->score[1];
            }
        }
    }
    if (playout_plies) *playout_plies = plies;
    free(t.nodes);
    free(t.action);
    free(t.child);
    free(t.edge);
    free(t.tab);
    for (size_t i = 0; i < grave.n; ++i) free(grave.tabs[i]);
    free(grave.tabs);
    free(grave.states);
    free(grave.slots);
    free(replies);
    free(replies2);
    free(bigram);
    for (int q = 0; q < ORC_MAX_PIPELINE; ++q) {
        free(trace2[q]);
        free(rec2[q]);
        free(bg_delta[q]);
        free(upd_order[q]);
        free(upd_action[q]);
        free(last_win[q]);
        free(set_paths[q]);
        free(set_plen[q]);
        free(set_res[q]);
        free(set_amaf[q]);
        free(set_delta[q]);
    }
    free(below_a);
    free(below_p);
    free(mast);
    free(leaves);
    return rc;
}

/* ---- known-answer hooks: the formulas and the AMAF update above on hand-built inputs (the checker's side of the
 * product's mr_select_score / mr_kat_update_amaf) ---- */
static void stats_from_edge(stats_t *s, const orc_edge_stats *e) {
    memset(s, 0, sizeof *s);
    s->visits = e->visits;
    s->vloss = e->vloss;
    for (int p = 0; p < 2; ++p) {
        s->score[p] = e->score[p];
        s->sumsq[p] = e->sumsq[p];
        s->amaf_score[p] = e->amaf_score[p];
    }
    s->amaf_visits = e->amaf_visits;
}

double orc_select_score(const orc_search_cfg *c, int player, const orc_edge_stats *current, const orc_edge_stats *child,
                        uint32_t grave_n, int64_t grave_q) {
    sel_cfg cfg = {c->select, c->q_init, c->exploration_c, c->expand_threshold, c->seed, c->grave_threshold, c->rave_schedule,
                   c->rave_param, NULL, c->amaf_alpha, c->rave_bias};
    stats_t cur, e;
    stats_from_edge(&cur, current);
    double parent_log = log(fmax(1.0, (double)cur.visits));
    if (!child) return unvisited_score(&cfg, &cur, player, parent_log);
    stats_from_edge(&e, child);
    return created_score(&cfg, &e, player, parent_log, grave_n, (double)grave_q);
}

int orc_kat_update_amaf(int game, const orc_state *root, const uint16_t *created, uint32_t n_created, uint16_t processed,
                        const orc_action_stat *amaf, const int64_t utilities[2], uint32_t k, orc_edge_stats *children_out,
                        uint32_t *n_children) {
    tree_t t;
    memset(&t, 0, sizeof t);
    t.game = game;
    add_node(&t, root);
    expand(&t, 0);
    if (t.nodes[0].status != 1) return 1;
    int pj = -1;
    for (uint32_t c = 0; c <= n_created; ++c) {
        uint16_t a = c < n_created ? created[c] : processed;
        int found = -1;
        for (int j = 0; j < t.nodes[0].n_children; ++j)
            if (t.action[t.nodes[0].first + j] == a) found = j;
        if (found < 0) return 2;
        if (t.child[t.nodes[0].first + found] < 0) {
            orc_state s = t.nodes[0].state;
            orc_apply(game, &s, a);
            t.child[t.nodes[0].first + found] = (int32_t)add_node(&t, &s);
        }
        if (c == n_created) pj = found;
    }
    entry_t path[2] = {{0, 0}, {(uint32_t)t.child[t.nodes[0].first + pj], (uint32_t)pj}};
    uint32_t below_a[4];
    int below_p[4];
    update_amaf_path(&t, path, 2, amaf, orc_num_actions(game), utilities, k, below_a, below_p);
    *n_children = (uint32_t)t.nodes[0].n_children;
    for (int j = 0; j < t.nodes[0].n_children; ++j) {
        orc_edge_stats *o = &children_out[j];
        const stats_t *e = &t.edge[t.nodes[0].first + j];
        memset(o, 0, sizeof *o);
        o->action = t.action[t.nodes[0].first + j];
        o->created = t.child[t.nodes[0].first + j] >= 0;
        o->amaf_visits = e->amaf_visits;
        o->amaf_score[0] = e->amaf_score[0];
        o->amaf_score[1] = e->amaf_score[1];
    }
    free(t.nodes);
    free(t.action);
    free(t.child);
    free(t.edge);
    return 0;
}
