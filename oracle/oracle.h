/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the simulation (playout) phase of thomasmarsh/mcts:
 * the five bitboard games, the playout loop, every simulation policy of strategy.rs but
 * MetaMcts, the statistics a playout feeds into backprop, and the MCTS loop around them.  It exists to CHECK the CUDA path
 * (tests/, __graft_entry__.smoke()) and to be TIMED as the CPU baseline
 * (bench.py cpu_baseline / --impl reference).  The product (mcts_b200/, include/)
 * never links, imports or calls anything in this directory.
 *
 * Pinning status: the reference is a Rust workspace and no Rust toolchain exists in
 * this image, so the reference itself cannot be run here (no oracle/_ref).  The
 * restatement is pinned against every golden vector / known-answer test the
 * reference's own tests hold for this path (tests/test_oracle_golden.py lists them
 * with file:line), and against restatements of the reference tests' own naive
 * array-board oracles (oracle/naive_ref.c).  NOT pinned by any reference test, and
 * therefore pinned only by reading code: the ORDER of generate_actions output, the
 * random_best visiting order, DecisiveMove's Win-mode fallback, "no actions => draw".
 * The RNG stream is deliberately NOT the reference's (rand 0.8 SmallRng, a crates.io
 * dependency absent from /root/reference): both this oracle and the CUDA kernels
 * use the counter-based Philox4x32-10 draw discipline documented in DESIGN.md.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CONNECT4 = 0, ORC_BREAKTHROUGH = 1, ORC_KNIGHTTHROUGH = 2, ORC_OTHELLO = 3, ORC_HEX11 = 4, ORC_NUM_GAMES = 5 };
enum { ORC_UNIFORM = 0, ORC_EPS_MAST = 1, ORC_DECISIVE_WIN = 2, ORC_DECISIVE_WINLOSS = 3, ORC_DECISIVE_WINLOSSDRAW = 4, ORC_DECISIVE_ANTI = 5, ORC_EPS_LGR = 6,
       ORC_DECISIVE_MAST_WIN = 7, ORC_DECISIVE_MAST_WINLOSS = 8, /* DecisiveMove<EpsilonGreedy<Mast>>, modes Win / WinLoss */
       ORC_EPS_NST = 9, /* EpsilonGreedy<Nst> */
       ORC_EPS_LGR2 = 10, /* EpsilonGreedy<Lgr2<Lgr<Uniform>>> */
       ORC_EPS_LGR2_MAST = 11 /* EpsilonGreedy<Lgr2<Lgr<Mast>>> (strategy.rs:158-171 Ucb1Lgr2Mast) */ };
enum { ORC_NOT_TERMINAL = 0, ORC_DRAW = 1, ORC_WINNER_P0 = 2, ORC_WINNER_P1 = 3 };
enum { ORC_END_TERMINAL = 0, ORC_END_TURN_LIMIT = 1, ORC_END_NO_ACTIONS = 2 };

#define ORC_MAX_ACTIONS 256
#define ORC_OTHELLO_PASS 64

/* Same byte layout as include/mcts_rollout.h's mr_leaf (declared separately on
 * purpose: the oracle shares no header with the product).
 *   Connect4/Breakthrough/Knightthrough/Othello: board[0]=black, board[1]=white in the
 *     reference's row-major layout (idx = row*cols + col, row 0 = south/bottom).
 *   Hex 11x11: board[0..1] = P0 words (low word first), board[2..3] = P1 words.
 *   turn: player to move (0 = Black / P0).
 *   flag: Connect4/Breakthrough/Knightthrough `winner`; Othello `last_pass`.
 *   prev_plus1: 1 + compact code of the action played just before this state (the last tree edge), 0 = none;
 *     the `prev_action` argument of SimulateStrategy::playout (simulate.rs:67-90), read by ORC_EPS_LGR only and
 *     not part of the game state (final states report 0). */
typedef struct {
    uint64_t board[4];
    uint8_t turn;
    uint8_t flag;
    uint16_t prev_plus1;  /* call context: 1 + the action that led to this state (playout()'s prev_action), 0 = none */
    uint16_t prev2_plus1; /* call context: 1 + the action before that (made by the player now to move), 0 = none */
    uint8_t pad[2];
} orc_state;

typedef struct {
    uint32_t wins[2];
    uint32_t draws;   /* every zero-utility ending: terminal draw, depth cutoff, no-actions */
    uint32_t cutoffs; /* the subset of `draws` that ended by max_playout_depth */
    uint64_t sum_depth;
} orc_leaf_result;

typedef struct {
    uint32_t visits;
    int32_t score;
} orc_action_stat;

typedef struct {
    uint32_t depth;
    uint8_t outcome;  /* ORC_DRAW / ORC_WINNER_P0 / ORC_WINNER_P1 as utilities see it; NOT_TERMINAL endings map to ORC_DRAW */
    uint8_t end_type; /* ORC_END_* */
    uint8_t pad[2];
} orc_playout_record;

/* ---- rules (games.c) ---- */
void orc_initial_state(int game, orc_state *out);
int orc_num_actions(int game); /* size of the dense action-id space (MAST/AMAF tables) */
int orc_num_slots(int game);   /* size of the compact per-player action index (bigram tables) */
int orc_action_slot(int game, int player, uint16_t code); /* compact index of `player`'s action, -1 if it is not one */
int orc_generate_actions(int game, const orc_state *s, uint16_t *out);
void orc_apply(int game, orc_state *s, uint16_t action);
int orc_is_terminal(int game, const orc_state *s);
int orc_winner(int game, const orc_state *s); /* -1 none, else player index */
int orc_terminal_status(int game, const orc_state *s);
int orc_player_to_move(int game, const orc_state *s);
uint64_t orc_hex_flood_iterations(void); /* instrumentation: flood6 fix-point iterations since reset */
void orc_hex_flood_iterations_reset(void);
/* generic-size Connect4 (the reference's 4x5 draw test, connect4/lib.rs:534-553) */
int orc_c4_generic_random_game(int rows, int cols, uint64_t seed, int *out_plies, int *out_has_winner);
/* bare Othello bitboard functions for the KATs (othello/lib.rs:132-234) */
uint64_t orc_othello_generate_moves(uint64_t player, uint64_t opponent);
uint64_t orc_othello_get_flips(uint64_t player, uint64_t opponent, uint64_t move_mask);

/* ---- Philox4x32-10 and the draw discipline (playout.c) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t orc_draw(uint64_t seed, uint32_t leaf_id, uint32_t playout_id, uint32_t ply, uint32_t stream);
void orc_draw_pair(uint64_t seed, uint32_t leaf_id, uint32_t playout_id, uint32_t ply, uint32_t *x0, uint32_t *x1);

/* ---- playouts (playout.c) ---- */
typedef struct {
    int game;
    int policy;
    uint64_t eps_threshold; /* floor(eps * 2^32), in [0, 2^32] */
    uint32_t max_playout_depth; /* 0 = unlimited */
    uint64_t seed;
    uint32_t leaf_id_base;
    uint32_t playouts_per_leaf;
    uint32_t max_trace; /* u16 slots per playout in trace_out */
    /* ORC_EPS_LGR = EpsilonGreedy<Lgr<Uniform>> (simulate.rs:538-570, 936-1033; strategy.rs:130-143 Ucb1Lgr).
     * replies: [2][num_actions] = 1 + compact code of player p's recorded reply to the preceding action with that
     * dense id (0 = none); frozen for the batch.  Outputs (nullable), the batch's last-write-wins updates
     * (backprop.rs:908-923) in a mergeable form: for every (p, id of the preceding action) the LAST occurrence, in
     * (global playout index, ply) order, of a reply by p in a playout p did not lose:
     *   reply_order_out  [2][num_actions] = ((g + 1) << 13) | (4096 + ply), g = leaf * k + playout; 0 = none
     *   reply_action_out [2][num_actions] = 1 + compact code of that reply
     *   last_win_out     [n_leaves][2]    = 1 + the leaf's last playout index that p did not lose (the trials in
     *                                       which the tree-path pairs above the leaf are written); 0 = none */
    const uint16_t *replies;
    uint64_t *reply_order_out;
    uint16_t *reply_action_out;
    uint32_t *last_win_out;
    /* ORC_EPS_NST = EpsilonGreedy<Nst> (simulate.rs:852-930; strategy.rs:113-126 Ucb1Nst): an action is scored by the
     * bigram average of (preceding action, action) once that pair has >= nst_threshold visits, else by its unigram
     * (MAST) score; random_best over the scores.  bigram: [2][S][S] indexed [mover][slot of the preceding action, in
     * the other player's numbering][slot of the action], S = orc_num_slots; frozen for the batch; the unigram table
     * is the `mast` argument.  bigram_delta_out (nullable, ADDED to): the playouts' consecutive pairs, including
     * (leaf's preceding action, first playout action), visits += 1 and score += utilities[mover] (backprop.rs:885-899). */
    const orc_action_stat *bigram;
    uint32_t nst_threshold;
    orc_action_stat *bigram_delta_out;
    /* ORC_EPS_LGR2 = EpsilonGreedy<Lgr2<Lgr<Uniform>>> (simulate.rs:1036-1142; strategy.rs:143-156 Ucb1Lgr2): unless the
     * ply explores, the mover plays replies2[mover][its own previous action IN THIS PLAYOUT][the opponent's reply to it]
     * when recorded and legal, else the LGR-1 reply (`replies`), else the uniform pick.  replies2: [2][S][S] over action
     * slots, values 1 + compact code, 0 = none; frozen for the batch.
     * replies_final_out [2][num_actions] / replies2_final_out [2][S][S] (nullable): the two tables after the LITERAL
     * trial-by-trial updates of backprop.rs:908-966 for every playout of the batch in global playout order, restricted
     * to the windows that end in a playout action (the leaf's prev / prev2 supply the first contexts; windows lying
     * wholly in the tree path are the caller's): LGR-1 last write wins; LGR-2 inserts when the player did not lose
     * and REMOVES the entry when the player lost and the entry still names that action. */
    const uint16_t *replies2;
    uint16_t *replies_final_out;
    uint16_t *replies2_final_out;
} orc_batch_desc;

/* One playout.  trace (nullable) receives up to max_trace action ids.  final_state nullable. */
void orc_playout(const orc_batch_desc *d, const orc_state *leaf, uint32_t leaf_id, uint32_t playout_id,
                 const orc_action_stat *mast /* [2][num_actions], nullable */, uint16_t *trace, orc_playout_record *rec,
                 orc_state *final_state);

/* A whole batch, optionally multi-threaded (pthreads).  All outputs nullable except results.
 *   amaf_out / mast_delta_out are ADDED to (caller zeroes).
 *   amaf_out:       [n_leaves][2][num_actions]
 *   mast_delta_out: [2][num_actions]
 *   trace_out:      [n_leaves*k][max_trace] u16,  rec_out / final_out: [n_leaves*k] */
int orc_playout_batch(const orc_batch_desc *d, const orc_state *leaves, uint32_t n_leaves, const orc_action_stat *mast,
                      orc_leaf_result *results, orc_action_stat *amaf_out, orc_action_stat *mast_delta_out,
                      uint16_t *trace_out, orc_playout_record *rec_out, orc_state *final_out, int n_threads);

/* Random-play openings: position j = final state of the j-th surviving uniform playout (leaf_id 0,
 * playout ids 0,1,2,...) from the initial position cut at `plies` plies; playouts that end earlier are
 * skipped.  Same definition as the product's GpuRollout.random_positions. */
void orc_random_positions(int game, uint64_t seed, uint32_t plies, uint32_t n, orc_state *out);

/* ---- restated reference TEST oracles (naive_ref.c) ---- */
uint64_t orc_naive_othello_generate_moves(uint64_t player, uint64_t opponent);
uint64_t orc_naive_othello_get_flips(uint64_t player, uint64_t opponent, uint64_t move_mask);
int orc_naive_breakthrough_moves(const int8_t board[64], int turn, uint16_t *out);
int orc_naive_knightthrough_moves(const int8_t board[64], int turn, uint16_t *out);
int orc_naive_hex_winner(const uint8_t occ0[121], const uint8_t occ1[121], int last_mover);
/* the reference's own stress sweeps, restated (mcts-tests/tests/stress.rs); return 0 on agreement,
 * otherwise a non-zero code identifying the first mismatch */
int orc_stress_othello(uint64_t num_games, uint64_t *games_done, uint64_t *plies_done);
int orc_stress_through(int game, uint64_t num_games, uint64_t *goal_wins, uint64_t *stuck_games);

/* ---- operation counters (bitboard.h): live only in _build/liboracle_ops.so (-DORC_COUNT_OPS); single-threaded runs ---- */
int orc_ops_enabled(void);
void orc_ops_reset(void);
void orc_ops_read(uint64_t *out /* [ORC_OP_KINDS] */);

/* ---- the MCTS loop around the playout path (search.c): BASELINE.json config 1 on the CPU, and the checker of
 * the product's GPU-driven tree driver ---- */
typedef struct {
    int game;
    int policy;
    int select;                 /* 0 UCB1, 1 UCB1-Tuned (select/ucb.rs), 2 GRAVE (select/rave.rs, RaveUcb::Ucb1),
                                   3 AMAF (select/amaf.rs) */
    int q_init;                 /* 0 Parent, 1 Infinity (node.rs:95-159) */
    uint32_t max_iterations;    /* leaf selections */
    uint32_t batch_leaves;      /* 1 = the reference's sequential loop */
    uint32_t playouts_per_leaf; /* num_rollouts_per_leaf */
    uint32_t expand_threshold;
    uint32_t max_playout_depth;
    int threads;                /* for the playouts of one batch */
    int pipeline_depth;         /* batches in flight (<= 1: none): batch n is gathered before the results of batches
                                   n - depth + 1 .. n - 1 are applied (deterministic order) */
    int use_transpositions;     /* SearchConfig.use_transpositions: identical states share a node (shared.rs:412-436);
                                   | 2: keyed as the reference keys Othello, by the canonical-symmetry image up to 20 discs
                                   (othello/lib.rs:474-480), with the action translation of symmetry.rs:38-50 */
    double exploration_c;
    uint64_t eps_threshold;
    uint64_t seed;
    uint32_t grave_threshold;   /* Rave.threshold */
    int rave_schedule;          /* 0 HandSelected{k}, 1 Threshold{rave}, 2 MinMSE{bias} (rave.rs:42-76) */
    uint32_t rave_param;
    double amaf_alpha;          /* select 3: alpha * amaf + (1 - alpha) * ucb1 (select/amaf.rs:58-74), reference default 1 */
    uint32_t nst_threshold;     /* ORC_EPS_NST: Nst::backoff_threshold (simulate.rs:875-881, reference default 5) */
    double rave_bias;           /* rave_schedule 2 = MinMSE{bias} (rave.rs:68) */
} orc_search_cfg;
#define ORC_MAX_PIPELINE 4

typedef struct {
    uint16_t action;
    uint16_t pad;
    uint32_t visits;
    int64_t score[2];
} orc_root_child;

int orc_search(const orc_search_cfg *c, const orc_state *root, uint16_t *best_action, orc_root_child *children,
               uint32_t *n_children, uint64_t *playout_plies);

/* known-answer hooks (search.c): one NodeStats / ChildArray row, the selection formulas and update_amaf on hand-built
 * inputs -- the restated forms of mcts-tests/tests/ttt_strategies.rs:658-736 and of select/ucb.rs, rave.rs, amaf.rs */
typedef struct {
    uint32_t visits, vloss;
    int64_t score[2], sumsq[2];
    uint32_t amaf_visits;
    uint16_t action, created;
    int64_t amaf_score[2];
} orc_edge_stats;
/* Game::canonical_representation (othello/lib.rs:618-646; the identity for the other games) and the D4 index maps of
 * game-core symmetry.rs:113-147, exposed for the tests */
int orc_canonical_representation(int game, const orc_state *s, orc_state *canon);
int orc_symmetry_index(int i, int k);
int orc_symmetry_invert(int i, int k);
double orc_select_score(const orc_search_cfg *c, int player, const orc_edge_stats *current, const orc_edge_stats *child,
                        uint32_t grave_n, int64_t grave_q);
int orc_kat_update_amaf(int game, const orc_state *root, const uint16_t *created, uint32_t n_created, uint16_t processed,
                        const orc_action_stat *amaf, const int64_t utilities[2], uint32_t k, orc_edge_stats *children_out,
                        uint32_t *n_children);

#ifdef __cplusplus
}
#endif
#endif
