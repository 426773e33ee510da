/*
 * mcts_rollout.h -- C ABI of the B200-native rollout engine (libmcts_b200.so).
 *
 * This is the drop-in boundary for the SIMULATION PHASE of thomasmarsh/mcts.  The reference has no
 * FFI of its own: its plug-in surface for this path is the Rust trait
 *     SimulateStrategy<G>::playout(state, max_playout_depth, &TreeStats, prev_action, rng) -> Trial<G>
 *         (mcts/src/strategies/mcts/simulate.rs:48-155)
 * called from simulate_step (mcts/src/strategies/mcts/search/shared.rs:763-778) by
 * TreeSearch::simulate / simulate_many (search/core.rs:201-258) and the tree-parallel workers
 * (search/parallel.rs:237-249).  The entry points below are what a Rust
 * `GpuRollout<G>: SimulateStrategy<G>` plus a batched leaf-gather driver bind (the `extern "C"`
 * block is shown in INTEGRATION.md).  Plain C structs, caller-owned buffers, integer return codes,
 * nothing unwinds across the boundary, no torch types.
 *
 * Semantics of one batch: for every leaf i in [0, n_leaves) run `playouts_per_leaf` playouts of the
 * named game and policy from `leaves[i]`, exactly as the reference's playout loop would
 * (simulate.rs:84-143 / 173-216), with every random choice drawn from the counter-based Philox4x32-10
 * stream keyed by (seed, leaf_id_base + i, playout index, ply) -- see DESIGN.md "Draw discipline".
 * What comes back is the AGGREGATE the reference's backprop would have accumulated from those
 * Trials (backprop.rs:644-964, node.rs:673-681): per leaf {wins[2], draws, cutoffs, sum_depth}; optionally
 * the per-leaf AMAF/GRAVE occurrence table (backprop.rs:565-640) and the batch-wide MAST (GLOBAL) delta
 * (backprop.rs:857-870); optionally, for parity checking, every playout's move trace, outcome record and
 * final state.
 *
 * Threading: one mr_ctx per driving host thread; calls on one ctx are not thread-safe.
 * There is NO CPU fallback: every entry point fails with MR_ERR_NO_DEVICE / MR_ERR_CUDA when no
 * usable sm_100 device is present.
 */
#ifndef MCTS_ROLLOUT_H
#define MCTS_ROLLOUT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MR_ABI_VERSION 2
#define MR_NUM_SLOTS 4 /* batches in flight per ctx: each slot has its own stream; submit batch N+1.. while batch N runs */

typedef struct mr_ctx mr_ctx;

enum mr_status {
    MR_OK = 0,
    MR_ERR_INVALID = 1,     /* bad argument / unsupported combination; mr_last_error() says which */
    MR_ERR_CUDA = 2,        /* a CUDA runtime call failed */
    MR_ERR_NO_DEVICE = 3,   /* no CUDA device, or not compute capability 10.x */
    MR_ERR_BUSY = 4,        /* slot already has a batch in flight */
    MR_ERR_UNSUPPORTED = 5, /* game/policy/flag combination not built */
    MR_ERR_COMM = 6,        /* multi-process communicator failure */
    MR_ERR_NO_MEMORY = 7    /* the host tree of mr_search could not allocate (nothing unwinds across the boundary) */
};

/* Games on the path.  Leaf layout per game is documented at mr_leaf. */
enum mr_game {
    MR_CONNECT4 = 0,      /* games/connect4/src/lib.rs      Connect4<6,7> */
    MR_BREAKTHROUGH = 1,  /* games/breakthrough/src/lib.rs  Breakthrough<8,8> */
    MR_KNIGHTTHROUGH = 2, /* games/knightthrough/src/lib.rs Knightthrough<8,8> */
    MR_OTHELLO = 3,       /* games/othello/src/lib.rs */
    MR_HEX11 = 4,         /* games/hex-gen/src/lib.rs       Hex<11,2> */
    MR_NUM_GAMES = 5
};

/* Simulation policies (the SimulateStrategy implementors in scope). */
enum mr_policy {
    MR_UNIFORM = 0,     /* simulate.rs:157-216  Uniform */
    MR_EPS_MAST = 1,    /* simulate.rs:485-570 + 819-848  EpsilonGreedy<G, Mast> */
    MR_DECISIVE_WIN = 2, /* simulate.rs:573-760  DecisiveMove<G, Uniform>, mode Win */
    /* modes WinLoss and WinLossDraw (simulate.rs:645-669).  In all five games a move can only end the game in the
     * mover's favour, or as the single remaining action, so the three modes pick the same action and share one
     * kernel (DESIGN.md "Decisive move"). */
    MR_DECISIVE_WINLOSS = 3,
    MR_DECISIVE_WINLOSSDRAW = 4,
    /* mode AntiDecisive (simulate.rs:689-735): an immediate win first; else the FIRST action in list order whose
     * child is not terminal and leaves the opponent no immediately winning reply; else the inner (uniform) pick.
     */
    MR_DECISIVE_ANTI = 5,
    /* simulate.rs:936-1033 + 538-570  EpsilonGreedy<G, Lgr<G, Uniform>> (strategy.rs:130-143 Ucb1Lgr): with
     * probability epsilon a uniform pick; else the reply recorded for (mover, preceding action) if it is legal now;
     * else a uniform pick.  Two draws per ply, as MR_EPS_MAST.  The reply table is frozen for the batch
     * (mr_set_replies); the batch's last-write-wins updates come back through mr_reply_updates.  Needs a
     * bounded playout (max_playout_depth) for Knightthrough. */
    MR_EPS_LGR = 6,
    /* DecisiveMove<G, EpsilonGreedy<G, Mast>> (mcts-tune SimulateSpec::DecisiveMoveMast, config_ir.rs): the decisive
     * scan first, the epsilon-greedy MAST pick when it finds nothing.  Mode Win is the `ucb1_tuned_dm_mast` family,
     * mode WinLoss the `rave` family (family_catalog.rs:470-514); as for the plain decisive policies the modes pick
     * the same action in these games and share a kernel.  Two draws per ply, as MR_EPS_MAST. */
    MR_DECISIVE_MAST_WIN = 7,
    MR_DECISIVE_MAST_WINLOSS = 8,
    /* simulate.rs:852-930 + 538-570  EpsilonGreedy<G, Nst> (strategy.rs:113-126 Ucb1Nst): MAST whose score of an action
     * is the BIGRAM average of (preceding action, action) once that pair has at least `backoff_threshold` visits, the
     * unigram (MAST) score otherwise.  The unigram table is mr_submit's mast_in, the bigram table comes from
     * mr_set_bigrams; the batch's bigram statistics come back through mr_bigram_delta (the unigram ones as the usual
     * MR_WANT_MAST_DELTA).  Two draws per ply, as MR_EPS_MAST. */
    MR_EPS_NST = 9,
    /* simulate.rs:1036-1142 + 538-570  EpsilonGreedy<G, Lgr2<G, Lgr<G, Uniform>>> (strategy.rs:143-156 Ucb1Lgr2, LGRF-2):
     * unless the ply explores, the reply recorded for (mover, its own previous action IN THIS PLAYOUT, the opponent's
     * answer) if legal; else the LGR-1 reply of MR_EPS_LGR; else a uniform pick.  Both tables are frozen for the batch
     * (mr_set_replies, mr_set_replies2); the LGR-1 updates come back through mr_reply_updates, the 2-ply table's
     * inserts and removals ("forgetting") through mr_reply2_inserts / mr_reply2_resolve.  Two draws per ply. */
    MR_EPS_LGR2 = 10,
    /* EpsilonGreedy<G, Lgr2<G, Lgr<G, Mast>>> (strategy.rs:158-171 Ucb1Lgr2Mast): as MR_EPS_LGR2 with the greedy MAST pick
     * (random_best over the unigram scores of mast_in) instead of the uniform pick as the last fallback. */
    MR_EPS_LGR2_MAST = 11
};

enum mr_flags {
    MR_WANT_AMAF = 1,        /* per-leaf AMAF/GRAVE occurrence table (every game) */
    MR_WANT_MAST_DELTA = 2,  /* batch-wide MAST (GLOBAL) table delta */
    MR_WANT_TRACE = 4,       /* parity/debug: per-playout move trace + record */
    MR_WANT_FINAL_STATE = 8, /* parity/debug: per-playout final state */
    /* amaf_out / mast_delta_out (and d_amaf / d_mast_delta of mr_launch_device) are indexed by COMPACT per-player action
     * slot, [2][mr_num_slots(game)], instead of by dense action id, [2][mr_num_actions(game)]: 192 / 512 instead of 4096
     * entries per player for Breakthrough / Knightthrough (mr_action_slot / mr_slot_action convert); identical for the
     * other games.  What a host tree or an all-reduce wants; the dense form is what the parity tests compare. */
    MR_TABLES_IN_SLOTS = 16
};

/* TerminalStatus as utilities see it (mcts/src/game.rs:117-141). */
enum mr_outcome { MR_NOT_TERMINAL = 0, MR_DRAW = 1, MR_WINNER_P0 = 2, MR_WINNER_P1 = 3 };
/* EndType (simulate.rs:15-20) plus the "no actions but not terminal" ending (simulate.rs:112-116). */
enum mr_end_type { MR_END_TERMINAL = 0, MR_END_TURN_LIMIT = 1, MR_END_NO_ACTIONS = 2 };

/*
 * One leaf state: a 40-byte POD record, the playout-relevant fields of the reference's G::S.
 *   Connect4 / Breakthrough / Knightthrough / Othello:
 *       board[0] = black, board[1] = white, in the reference's row-major layout
 *       (bit idx = row*cols + col, row 0 = south/bottom; bitboard/src/board.rs:76-104)
 *   Hex 11x11: board[0..1] = occupied[P0] words (low word first), board[2..3] = occupied[P1] words
 *   turn: player to move (0 = Black / P0, 1 = White / P1)
 *   flag: Connect4 / Breakthrough / Knightthrough `winner`; Othello `last_pass`; Hex unused
 *   prev_action_plus1: 1 + compact code of the action played just before this state (the last edge of the tree
 *       descent), 0 = none -- the `prev_action` argument of SimulateStrategy::playout (simulate.rs:67-90).  Call
 *       context, not game state: read by MR_EPS_LGR / MR_EPS_NST / MR_EPS_LGR2 only, reported as 0 in final states.
 *   prev2_action_plus1: the same for the action before that one (the second-to-last edge of the descent, made by the
 *       player now to move), 0 = none; ignored when prev_action_plus1 is 0.  Read by MR_EPS_LGR2 only, and only for
 *       the table UPDATES of the first two plies (a playout's own-previous-move context starts empty, simulate.rs:98).
 * Zobrist hashes carried by the reference structs are tree-side state and are not part of the record.
 */
typedef struct {
    uint64_t board[4];
    uint8_t turn;
    uint8_t flag;
    uint16_t prev_action_plus1;
    uint16_t prev2_action_plus1;
    uint8_t pad[2];
} mr_leaf;

/* Aggregate of one leaf's playouts; what k calls of ChildArray::update / NodeStats::update need:
 *   num_visits += k; score[p] += wins[p] - (k - draws - wins[p]); sum_sq[p] += k - draws   (node.rs:295-302, 673-681)
 *   accum_depth += sum_depth (+ k * (stack.len() - 1) added by the host; shared.rs:814-818) */
typedef struct {
    uint32_t wins[2];
    uint32_t draws;   /* every zero-utility ending: terminal draw, depth cutoff, no-actions */
    uint32_t cutoffs; /* the subset of draws that ended by max_playout_depth (EndType::TurnLimit) */
    uint64_t sum_depth;
} mr_leaf_result;

/* node.rs:51-55 ActionStats; score is the integer sum of +1/0/-1 utilities */
typedef struct {
    uint32_t visits;
    int32_t score;
} mr_action_stat;

typedef struct {
    uint32_t depth;
    uint8_t outcome;  /* mr_outcome; non-terminal endings are reported as MR_DRAW */
    uint8_t end_type; /* mr_end_type */
    uint8_t pad[2];
} mr_playout_record;

typedef struct {
    uint32_t game;              /* mr_game */
    uint32_t policy;            /* mr_policy */
    uint32_t n_leaves;
    uint32_t playouts_per_leaf; /* k */
    uint32_t max_playout_depth; /* SearchConfig.max_playout_depth (config.rs:440); 0 = unlimited */
    uint32_t flags;             /* mr_flags */
    uint64_t seed;              /* Philox key */
    uint32_t leaf_id_base;      /* Philox counter word 0 of leaf i is leaf_id_base + i */
    uint32_t max_trace;         /* u16 slots per playout in trace_out (MR_WANT_TRACE) */
    double epsilon;             /* EpsilonGreedy.epsilon (simulate.rs:525 default 0.1); MR_EPS_MAST only */
} mr_batch_desc;

/* Dense action-id space of a game (the index into MAST / AMAF tables):
 *   Connect4 7 (column); Othello 65 (square, 64 = PASS); Hex 121 (cell);
 *   Breakthrough / Knightthrough 4096 (src*64 + dst).
 * Trace entries use the compact u16 action code: (src << 8) | dst for Breakthrough / Knightthrough,
 * the action id itself for the other games. */
int mr_num_actions(int game);
int mr_abi_version(void);

/* Creates a context driving `n_devices` GPUs of this process (one stream pair, pinned staging and
 * device buffers per GPU; batches are sharded across them by playout range with no data-path
 * collective).  device_ids == NULL means devices 0..n_devices-1. */
int mr_init(const int *device_ids, int n_devices, mr_ctx **out);
void mr_destroy(mr_ctx *ctx);
const char *mr_last_error(const mr_ctx *ctx); /* valid until the next call on ctx; ctx may be NULL */
int mr_num_devices(const mr_ctx *ctx);

/* Asynchronous batch: copies `leaves` (and the MAST snapshot, mast_in = [2][mr_num_actions], required
 * for MR_EPS_MAST) into pinned staging before returning and enqueues the batch on the slot's own stream, so up to
 * MR_NUM_SLOTS batches are in flight and mr_wait(slot) waits for that slot's batch only.  Batches without the parity /
 * LGR / NST outputs take the low-latency path: ONE kernel launch whose last block writes the results into host-mapped
 * memory and raises a flag the host spins on (no separate memset / copy / synchronize calls).  Replaces one
 * `simulate_many` fan-out (search/core.rs:218-258) per leaf of the batch. */
int mr_submit(mr_ctx *ctx, int slot, const mr_batch_desc *desc, const mr_leaf *leaves,
              const mr_action_stat *mast_in);

/* Blocks until the slot's batch is done and copies the results out.  Every pointer except `out` is
 * nullable and must be non-NULL only if the matching flag was set at submit:
 *   out            [n_leaves]
 *   amaf_out       [n_leaves][2][mr_num_actions]   (MR_WANT_AMAF; playout-trace occurrences only, the
 *                                                   tree-path suffix is added by the host, backprop.rs:833-854)
 *   mast_delta_out [2][mr_num_actions]             (MR_WANT_MAST_DELTA)
 *                  with MR_TABLES_IN_SLOTS both tables are [..][2][mr_num_slots] instead
 *   trace_out      [n_leaves*k][max_trace] u16, rec_out [n_leaves*k]   (MR_WANT_TRACE)
 *   final_out      [n_leaves*k]                    (MR_WANT_FINAL_STATE) */
int mr_wait(mr_ctx *ctx, int slot, mr_leaf_result *out, mr_action_stat *amaf_out, mr_action_stat *mast_delta_out,
            uint16_t *trace_out, mr_playout_record *rec_out, mr_leaf *final_out);

/* mr_submit + mr_wait on slot 0. */
int mr_playout_batch(mr_ctx *ctx, const mr_batch_desc *desc, const mr_leaf *leaves, const mr_action_stat *mast_in,
                     mr_leaf_result *out, mr_action_stat *amaf_out, mr_action_stat *mast_delta_out,
                     uint16_t *trace_out, mr_playout_record *rec_out, mr_leaf *final_out);

/* Device-resident variant for hosts that keep leaves in HBM (and for kernel-only timing): the leaf and
 * result arrays are caller-owned DEVICE pointers on device `device_index` of the ctx, the launch goes to
 * `stream` (a cudaStream_t passed as void*, NULL = the ctx's own stream) and nothing is copied.
 * d_results must be zeroed by the caller (results are accumulated with atomics).  d_amaf / d_mast_delta
 * are nullable device pointers matching desc->flags.  The MAST snapshot is the one last uploaded with
 * mr_set_mast. */
int mr_launch_device(mr_ctx *ctx, int device_index, const mr_batch_desc *desc, const void *d_leaves, void *d_results,
                     void *d_amaf, void *d_mast_delta, void *stream);
int mr_set_mast(mr_ctx *ctx, int game, const mr_action_stat *mast_in);

/* ---- Last Good Reply (MR_EPS_LGR) ----
 * TreeStats.player_replies (search/shared.rs; read simulate.rs:1013-1032, written backprop.rs:908-923) crosses the
 * ABI as replies[2][mr_num_actions]: 1 + compact code of player p's recorded reply to the preceding action with
 * that dense id, 0 = none.  mr_set_replies copies the table; every later MR_EPS_LGR submit uses that copy, frozen
 * for the batch.  NULL clears it.
 *
 * mr_reply_updates, called after mr_wait and before the slot's next submit, returns what the batch's trials would
 * have written, in a form the host can merge with its tree-path pairs: per (p, preceding action id) the LAST
 * occurrence -- in (global playout index g = leaf * k + playout, ply) order -- of a reply by p in a playout that p
 * did not lose (utility == the trial's maximum: the winner, or both players of a zero-utility ending):
 *   updates_out  [2][mr_num_actions]  order = ((g + 1) << 13) | (4096 + ply), 0 = no occurrence
 *   last_win_out [n_leaves][2]        1 + the leaf's last playout index that p did not lose, 0 = none (the trials in
 *                                     which the tree-path pairs above the leaf get written) */
typedef struct {
    uint64_t order;
    uint16_t action_plus1; /* 1 + compact code of the reply */
    uint16_t pad[3];
} mr_reply_update;
int mr_set_replies(mr_ctx *ctx, int game, const uint16_t *replies);
int mr_reply_updates(mr_ctx *ctx, int slot, mr_reply_update *updates_out, uint32_t *last_win_out);

/* ---- Last Good Reply with Forgetting, 2-ply (MR_EPS_LGR2) ----
 * TreeStats.player_replies2 (search/shared.rs:77-84; read simulate.rs:1113-1141, written backprop.rs:926-966) crosses the
 * ABI as replies2[2][S][S], S = mr_num_slots(game): [mover][slot of the mover's own previous action][slot of the
 * opponent's answer] = 1 + compact code of the recorded reply, 0 = none.  mr_set_replies2 copies it (NULL clears).
 *
 * The reference updates the table trial by trial: a player that did not lose INSERTS (context -> action) for each of
 * its plies, a player that lost REMOVES the entry of the context if it still names the action it just played.  The
 * outcome after a batch depends only on, per context, the LAST insert and whether a removal of that same action
 * follows it (all of a player's events in one trial are of one kind, so "follows" is decided by the global playout
 * index g alone).  The batch therefore reports in two steps, called after mr_wait and before the slot's next submit:
 *   mr_reply2_inserts   inserts_out [2][S][S]: per context the last insert of the PLAYOUT part (windows of three
 *                       consecutive actions of [leaf prev2, leaf prev, playout...] that end in a playout action), same
 *                       order stamp as mr_reply_update; last_lose_out [n_leaves][2] (nullable) = 1 + the leaf's last
 *                       playout index that p LOST (the trials in which tree-path windows remove), 0 = none
 *   mr_reply2_resolve   merged [2][S][S]: the CANDIDATE entry per context -- the caller's last insert after adding its
 *                       tree-path windows (order > 0), or the entry its table holds now (order = 0, action_plus1 =
 *                       that entry), or nothing (action_plus1 = 0).  Replays the batch on the GPU and sets removed_out
 *                       [2][S][S] to 1 where a losing playout with a LATER g played the candidate action in that
 *                       context, else 0.
 * New entry of a context = none if removed (or removed by one of the caller's own later tree-path windows), else the
 * merged insert's action, else the old entry. */
int mr_set_replies2(mr_ctx *ctx, int game, const uint16_t *replies2);
int mr_reply2_inserts(mr_ctx *ctx, int slot, mr_reply_update *inserts_out, uint32_t *last_lose_out);
int mr_reply2_resolve(mr_ctx *ctx, int slot, const mr_reply_update *merged, uint8_t *removed_out);

/* ---- N-gram Selection Technique (MR_EPS_NST) ----
 * Bigram tables are keyed by COMPACT per-player action indices ("slots"), because the dense id space of Breakthrough /
 * Knightthrough would be 4096 x 4096 per player:
 *   Connect4 column (7); Othello square or PASS (65); Hex cell (121);
 *   Breakthrough src*3 + kind, kind = 0,1,2 = the player's west-diagonal, straight, east-diagonal move
 *   (dst - src = -9,-8,-7 for Black, +7,+8,+9 for White) (192);
 *   Knightthrough src*8 + kind, kind = index of the jump in ascending-dst order -17,-15,-10,-6,+6,+10,+15,+17 (512).
 * TreeStats.player_bigram_actions (search/shared.rs:60-66) crosses the ABI as bigram[2][S][S], S = mr_num_slots(game),
 * indexed [mover][slot of the preceding action, in the OTHER player's numbering][slot of the action].
 * mr_set_bigrams copies the table (NULL = empty) and the backoff threshold (Nst::backoff_threshold, default 5); every
 * later MR_EPS_NST submit uses that copy.  mr_bigram_delta, called after mr_wait and before the slot's next submit,
 * ADDS NOTHING and OVERWRITES bigram_delta_out[2][S][S] with the playouts' consecutive (preceding action, action)
 * pairs -- including (leaf's prev_action_plus1, first playout action) -- visits += 1, score += utilities[mover]
 * (backprop.rs:885-899); the tree-path pairs are the host's. */
int mr_num_slots(int game);
int mr_action_slot(int game, int player, uint16_t code); /* -1 when `code` is not an action of `player` */
int mr_slot_action(int game, int player, int slot);      /* compact action code, -1 when no such action exists */
int mr_set_bigrams(mr_ctx *ctx, int game, const mr_action_stat *bigram, uint32_t backoff_threshold);
int mr_bigram_delta(mr_ctx *ctx, int slot, mr_action_stat *bigram_delta_out);

/* Kernel time of the last completed batch on `slot` (max over the ctx's devices), measured with CUDA
 * events on the launching stream; and the number of rollout-kernel launches the ctx has issued. */
int mr_last_kernel_ms(mr_ctx *ctx, int slot, float *ms);
uint64_t mr_kernel_launches(const mr_ctx *ctx);
/* Kernel timing costs two event records per batch and, in mr_wait, a wait for the end event; latency-critical hosts
 * (the search driver) switch it off: mr_last_kernel_ms then reports 0.  Enabled by default. */
int mr_set_kernel_timing(mr_ctx *ctx, int enabled);

/* Sums `n` int32 values element-wise across the devices/ranks of the search (the per-root-child
 * visit/score arrays of root-parallel search, search/parallel.rs:94-105).  Within one process it is a
 * host-side sum over the ctx's per-device partial arrays: bufs[d] points at device d's `n` values and
 * every bufs[d] receives the total.  Across processes use mr_comm_*. */
int mr_allreduce_root(mr_ctx *ctx, int32_t *const *bufs, int n_bufs, size_t n);

/* Multi-process (one rank per GPU) root-statistics merge over NCCL.  The unique id is created on rank 0
 * with mr_comm_unique_id, broadcast by the host's own means, and passed to every rank's mr_comm_init.
 * mr_comm_allreduce_i32 sums a DEVICE buffer in place on the ctx's stream (ncclAllReduce, ncclSum). */
#define MR_COMM_ID_BYTES 128
int mr_comm_unique_id(uint8_t id[MR_COMM_ID_BYTES]);
int mr_comm_init(mr_ctx *ctx, const uint8_t id[MR_COMM_ID_BYTES], int rank, int world_size);
int mr_comm_allreduce_i32(mr_ctx *ctx, void *d_buf, size_t n, void *stream);

/*
 * ---- Batched leaf-gather search driver ----
 * A host tree (arena, UCB1 / UCB1-Tuned selection, expansion, aggregated backprop) whose simulation phase
 * is the GPU engine: per batch it runs `batch_leaves` descents holding virtual loss (search/shared.rs:569-736,
 * 665-671, 787-805), submits the leaves as one batch of `playouts_per_leaf` playouts each, and applies the
 * aggregated results along every path (backprop.rs:704-731 with node.rs:673-681 summed over the k trials).
 * With batch_leaves = 1 it is the reference's sequential loop (search/search_impl.rs:76-114); larger batches
 * are the leaf-parallel generalisation of the tree-parallel worker loop (search/parallel.rs:188-251).
 * In the reference this driver is Rust; here it is C++ inside the library because no Rust toolchain exists
 * in the build image -- a Rust host would keep its own tree and call mr_submit / mr_wait directly.
 */
enum mr_select {
    MR_SELECT_UCB1 = 0,       /* select/ucb.rs:9-62 */
    MR_SELECT_UCB1_TUNED = 1, /* select/ucb.rs:64-146 */
    MR_SELECT_GRAVE = 2,      /* select/rave.rs:118-238 Rave{threshold, schedule, RaveUcb::Ucb1}: reads the AMAF/GRAVE
                                 tables the kernels accumulate (threshold 0 = RAVE, huge = HRAVE) */
    MR_SELECT_AMAF = 3        /* select/amaf.rs:48-86 alpha * amaf + (1 - alpha) * ucb1 over per-edge AMAF statistics
                                 (node.rs:711-719), written by update_amaf (backprop.rs:565-617) from the same
                                 per-leaf occurrence tables */
};
/* select/rave.rs:42-76: HandSelected sqrt(k / (3n + k)); Threshold max(0, rave - n) / rave;
 * MinMSE amaf_n / (n + amaf_n + 4 n amaf_n bias^2) (rave_bias) */
enum mr_rave_schedule { MR_RAVE_HAND_SELECTED = 0, MR_RAVE_THRESHOLD = 1, MR_RAVE_MIN_MSE = 2 };
enum mr_qinit { MR_QINIT_PARENT = 0, MR_QINIT_INFINITY = 1 };     /* node.rs:95-159 */
/* MR_SEARCH_TRANSPOSITIONS: SearchConfig.use_transpositions -- a new child whose state is already in the
 * table reuses that node (search/shared.rs:412-436, table.rs); statistics stay on the edges.  The table is
 * keyed by the full 40-byte state record (exact, no Zobrist collisions); without MR_SEARCH_SYMMETRY no symmetry
 * canonicalisation is applied. */
/* MR_SEARCH_PIPELINED: two batches in flight (= pipeline_depth 2) -- batch N+1 is gathered (with batch N's virtual
 * losses still in place) and submitted before batch N's results are applied. */
/* MR_SEARCH_SYMMETRY (implies MR_SEARCH_TRANSPOSITIONS; Othello, ignored for the other games): the table is keyed as the
 * reference keys it -- by the position's canonical D4 orientation while at most SYMMETRY_PLY_LIMIT = 20 discs are on the
 * board (games/othello/src/lib.rs:26, 436-443, 474-480), by the position itself later.  Every node but the root holds its
 * action list in the canonical orientation (search/shared.rs:233-252); a descent carries the real state, translates each
 * chosen action back through that state's symmetry (mcts/src/symmetry.rs:38-50, othello/lib.rs:658-665) and verifies a
 * cached child against the real successor (shared.rs:382-405).  Playouts start from the real state.  Only with
 * selection strategies and rollout policies that keep no action-keyed tables (UCB1 / UCB1-Tuned; uniform / decisive):
 * MR_ERR_UNSUPPORTED otherwise. */
enum mr_search_flags { MR_SEARCH_TRANSPOSITIONS = 1, MR_SEARCH_PIPELINED = 2, MR_SEARCH_SYMMETRY = 4 };

typedef struct {
    uint32_t game;              /* mr_game */
    uint32_t policy;            /* mr_policy */
    uint32_t select;            /* mr_select */
    uint32_t q_init;            /* mr_qinit */
    uint32_t max_iterations;    /* SearchConfig.max_iterations: number of leaf selections */
    uint32_t batch_leaves;      /* leaves gathered per GPU batch (1 = the reference's sequential loop) */
    uint32_t playouts_per_leaf; /* SearchConfig.num_rollouts_per_leaf */
    uint32_t expand_threshold;  /* SearchConfig.expand_threshold (default 1) */
    uint32_t max_playout_depth; /* 0 = unlimited */
    uint32_t flags;             /* mr_search_flags */
    double exploration_c;       /* sqrt(2) by default (select/ucb.rs:22-27) */
    double epsilon;             /* MR_EPS_MAST */
    uint64_t seed;
    uint32_t grave_threshold;   /* Rave.threshold (default 700) */
    uint32_t rave_schedule;     /* mr_rave_schedule */
    uint32_t rave_param;        /* HandSelected k (default 1000) / Threshold rave */
    uint32_t nst_backoff_threshold; /* MR_EPS_NST: Nst::backoff_threshold (simulate.rs:875-881; the reference's default is 5) */
    double amaf_alpha;          /* MR_SELECT_AMAF: 1 = pure AMAF (the reference default), 0 = UCB1 */
    double rave_bias;           /* MR_RAVE_MIN_MSE: RaveSchedule::MinMSE { bias } */
    /* Batches in flight, 1..MR_NUM_SLOTS (0 = 1, or 2 with MR_SEARCH_PIPELINED).  Batch n is gathered and submitted
     * before the results of batches n - depth + 1 .. n - 1 are applied: the host's select / backprop work overlaps the
     * GPU, at the price of descents that see in-flight batches only as virtual losses (what the reference's
     * tree-parallel workers see of each other, search/parallel.rs:188-251).  Deterministic for a given depth. */
    uint32_t pipeline_depth;
    uint32_t reserved;
} mr_search_desc;

typedef struct {
    uint16_t action; /* compact action code, as in traces */
    uint16_t pad;
    uint32_t visits;
    int64_t score[2]; /* sum of utilities per player */
} mr_root_child;

typedef struct {
    uint64_t playouts;
    uint64_t playout_plies; /* sum of playout depths (accum_depth without the tree part) */
    uint64_t nodes;
    uint32_t batches;
    uint32_t max_tree_depth;
    double seconds_select;   /* host: descents (select_step) of all batches */
    double seconds_gpu;      /* host time spent WAITING for batches (small when the pipeline hides the GPU) */
    double seconds_backprop; /* host: aggregated backprop + table merges */
    double seconds_submit;   /* host: mr_submit calls (staging copy, MAST ranking, kernel launch) */
} mr_search_stats;

/* The host-side `Game` surface the driver expands nodes with (generate_actions order == child order,
 * game.rs:144-211).  Exposed so that hosts and tests can step leaf records without a second rules copy. */
int mr_generate_actions(int game, const mr_leaf *state, uint16_t *out /* [MR_MAX_ROOT_CHILDREN] */);
int mr_apply(int game, mr_leaf *state, uint16_t action);
int mr_terminal_status(int game, const mr_leaf *state); /* mr_outcome */
/* Game::canonical_representation (game.rs; games/othello/src/lib.rs:618-646): writes the state in its canonical
 * orientation and returns the index (0..7, game-core symmetry.rs:113-127 order) of the symmetry that maps the state onto
 * it -- what MR_SEARCH_SYMMETRY keys the transposition table by.  Othello up to 20 discs; the identity (0, a plain
 * copy) for everything else.  Negative on bad arguments. */
int mr_canonical_representation(int game, const mr_leaf *state, mr_leaf *canonical_out);

#define MR_MAX_ROOT_CHILDREN 256
/* Runs one search from `root`; writes the chosen action (RobustChild: most visits, then expected score,
 * select/basic.rs:13-45), the root children's statistics (children[MR_MAX_ROOT_CHILDREN]) and timing. */
int mr_search(mr_ctx *ctx, const mr_search_desc *desc, const mr_leaf *root, uint16_t *best_action,
              mr_root_child *children, uint32_t *n_children, mr_search_stats *stats);

/* Root-parallel search, choose_action_root_parallel (search/parallel.rs:56-115): `n_searchers` independent searches of
 * `desc->max_iterations` iterations each, one host thread and one ctx (its own streams; the ctxs may share a GPU) per
 * searcher, searcher i seeded from (desc->seed, i).  The root children's (visits, score) totals are summed per action and
 * the action with the most merged visits wins (random_best).  children / n_children receive the MERGED totals in root
 * action-list order, stats[n_searchers] each searcher's own counters. */
int mr_search_root_parallel(mr_ctx *const *ctxs, int n_searchers, const mr_search_desc *desc, const mr_leaf *root,
                            uint16_t *best_action, mr_root_child *children, uint32_t *n_children, mr_search_stats *stats);

/* ---- Known-answer hooks ----
 * The driver's selection formulas and its AMAF update, callable on hand-built inputs, so that the reference's own unit
 * tests of this logic (mcts-tests/tests/ttt_strategies.rs:658-736 for update_amaf; the formulas of select/ucb.rs:40-146,
 * select/rave.rs:60-75,196-224, select/amaf.rs:58-79) can be restated against the product code and not only against a
 * twin implementation.  One row of a ChildArray / NodeStats (node.rs:295-302, 392-396, 673-719): */
typedef struct {
    uint32_t visits, vloss;
    int64_t score[2];      /* sum of utilities per player */
    int64_t sumsq[2];      /* sum of squared utilities per player */
    uint32_t amaf_visits;
    uint16_t action;       /* (out) compact action code of the child */
    uint16_t created;      /* (out) the child node exists */
    int64_t amaf_score[2];
} mr_edge_stats;
/* FlatMonteCarloStrategy::choose_action (mcts/src/strategies/flat_mc.rs:58-151) as ONE batch: every root child is a leaf
 * with `samples_per_move` uniform playouts of at most `max_rollout_depth` moves (the reference's rollout loop never looks at
 * the state its last move reaches: a playout with max_playout_depth = max_rollout_depth - 1 scored 0 at the cutoff); a
 * child's value is the number of playouts the root's mover won, or with `use_ucb1` w / n + ucb1_c * ln(n) / n (:135); the
 * pick is random_best (util.rs:135-163) with the combined draw `r` in [0, 16 n).  wins_out[MR_MAX_ROOT_CHILDREN] receives
 * the wins per root action (generate_actions order), actions_out (nullable) the actions.  MR_ERR_INVALID for a root
 * without actions (the reference panics). */
int mr_flat_monte_carlo(mr_ctx *ctx, int game, const mr_leaf *root, uint32_t samples_per_move, uint32_t max_rollout_depth,
                        int use_ucb1, double ucb1_c, uint32_t r, uint64_t seed, uint16_t *best_action, uint32_t *wins_out,
                        uint16_t *actions_out, uint32_t *n_actions);

/* The value `player` (the mover at the current node) assigns to one child: `current` = the statistics of the node being
 * left (root_stats, or the edge into it), `child` = the child's edge row, NULL = a child that does not exist yet
 * (unvisited_value).  grave_n / grave_q = the reference node's AMAF count and score sum for (player, action), used by
 * MR_SELECT_GRAVE only.  Reads select, q_init, exploration_c, rave_*, amaf_alpha of `desc`. */
double mr_select_score(const mr_search_desc *desc, int player, const mr_edge_stats *current, const mr_edge_stats *child,
                       uint32_t grave_n, int64_t grave_q);
/* update_amaf (backprop.rs:565-617) on a two-level tree: `root` expanded, the children named in created[] and
 * `processed` created, the descent path = root -> processed.  `amaf` is a per-leaf occurrence table as mr_wait returns
 * it ([2][mr_num_actions]: per (player, action) the number of occurrences in the k trials and the sum of that player's
 * utilities over them), utilities[p] the k trials' utility sums.  children_out[MR_MAX_ROOT_CHILDREN] receives every
 * root child's AMAF row. */
int mr_kat_update_amaf(int game, const mr_leaf *root, const uint16_t *created, uint32_t n_created, uint16_t processed,
                       const mr_action_stat *amaf, const int64_t utilities[2], uint32_t k, mr_edge_stats *children_out,
                       uint32_t *n_children);

/* Integer-issue micro-benchmark used as the roofline denominator (DESIGN.md "Roofline"): runs a
 * dependent-chain kernel of `kind` (0 = LOP3, 1 = IADD3, 2 = SHF, 3 = IMAD, 4 = POPC, 5 = LOP3+IMAD mix)
 * on every SM and returns thread-level operations per second (further kinds: 6 = IMAD.HI, 7 = IMAD.WIDE + LOP3,
 * 8 = SHF by immediate, 9 = ISETP + SEL pair, each counted as one operation per chain step). */
int mr_measure_int_peak(mr_ctx *ctx, int device_index, int kind, double *ops_per_second, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* MCTS_ROLLOUT_H */
