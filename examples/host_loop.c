/*
 * examples/host_loop.c -- a plain-C host of the rollout engine, the shape a non-Python caller takes.
 *
 * What it does is what thomasmarsh/mcts's own hosts do around the simulation phase:
 *   1. one batch of playouts from a set of leaves           (TreeSearch::simulate_many, search/core.rs:218-258)
 *   2. a whole move decision with the batched tree driver   (Search::choose_action, search/core.rs:138-198)
 * Build:  gcc -O2 -std=c99 -Iinclude examples/host_loop.c -Lmcts_b200/_build -lmcts_b200 -Wl,-rpath,$PWD/mcts_b200/_build -o host_loop
 * Run  :  ./host_loop            (needs a B200; the library has no CPU fallback and says so)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcts_rollout.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != MR_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mr_last_error(ctx));       \
            return rc_ == MR_ERR_NO_DEVICE ? 3 : 1;                                  \
        }                                                                            \
    } while (0)

int main(void) {
    mr_ctx *ctx = NULL;
    CHECK(mr_init(NULL, 1, &ctx));

    /* ---- 1. a batch: the empty Connect4 board and its seven children, 4096 uniform playouts each ---- */
    mr_leaf leaves[8];
    memset(leaves, 0, sizeof leaves); /* the empty board, Black to move */
    uint16_t actions[MR_MAX_ROOT_CHILDREN];
    int n_actions = mr_generate_actions(MR_CONNECT4, &leaves[0], actions);
    for (int i = 0; i < n_actions; ++i) {
        leaves[1 + i] = leaves[0];
        CHECK(mr_apply(MR_CONNECT4, &leaves[1 + i], actions[i]));
    }
    mr_batch_desc d;
    memset(&d, 0, sizeof d);
    d.game = MR_CONNECT4;
    d.policy = MR_UNIFORM;
    d.n_leaves = 1 + (uint32_t)n_actions;
    d.playouts_per_leaf = 4096;
    d.seed = 0x5eedULL;
    mr_leaf_result res[8];
    CHECK(mr_playout_batch(ctx, &d, leaves, NULL, res, NULL, NULL, NULL, NULL, NULL));
    for (uint32_t i = 0; i < d.n_leaves; ++i) {
        /* k calls of ChildArray::update, summed (node.rs:673-681) */
        long k = d.playouts_per_leaf, w0 = res[i].wins[0], w1 = res[i].wins[1], dr = res[i].draws;
        if (w0 + w1 + dr != k) {
            fprintf(stderr, "leaf %u: %ld + %ld + %ld != %ld\n", i, w0, w1, dr, k);
            return 1;
        }
        printf("leaf %u: black %ld  white %ld  draws %ld  mean depth %.2f\n", i, w0, w1, dr, (double)res[i].sum_depth / (double)k);
    }

    /* ---- 2. a move: UCT (UCB1, c = sqrt 2), 64 leaves x 16 playouts per batch, four batches in flight (one per slot) ---- */
    mr_search_desc s;
    memset(&s, 0, sizeof s);
    s.game = MR_CONNECT4;
    s.policy = MR_UNIFORM;
    s.select = MR_SELECT_UCB1;
    s.q_init = MR_QINIT_PARENT;
    s.max_iterations = 8192;
    s.batch_leaves = 64;
    s.playouts_per_leaf = 16;
    s.expand_threshold = 1;
    s.pipeline_depth = MR_NUM_SLOTS;
    s.exploration_c = 1.4142135623730951;
    s.seed = 7;
    uint16_t best = 0;
    mr_root_child children[MR_MAX_ROOT_CHILDREN];
    uint32_t n_children = 0;
    mr_search_stats st;
    CHECK(mr_search(ctx, &s, &leaves[0], &best, children, &n_children, &st));
    uint64_t visits = 0;
    for (uint32_t i = 0; i < n_children; ++i) visits += children[i].visits;
    printf("search: %llu playouts in %u batches, %llu nodes, best column %u (%u children, %llu child visits)\n",
           (unsigned long long)st.playouts, st.batches, (unsigned long long)st.nodes, best, n_children, (unsigned long long)visits);
    /* ---- 3. the same decision by flat Monte Carlo (flat_mc.rs): every root child one leaf of one batch ---- */
    uint16_t flat_best = 0, flat_actions[MR_MAX_ROOT_CHILDREN];
    uint32_t flat_wins[MR_MAX_ROOT_CHILDREN], n_flat = 0;
    CHECK(mr_flat_monte_carlo(ctx, MR_CONNECT4, &leaves[0], 4096, 100, 0, 0.0, 5, 7, &flat_best, flat_wins, flat_actions, &n_flat));
    printf("flat MC: best column %u of %u (wins of 4096:", flat_best, n_flat);
    for (uint32_t i = 0; i < n_flat; ++i) printf(" %u:%u", flat_actions[i], flat_wins[i]);
    printf(")\n");
    printf("kernel launches: %llu\n", (unsigned long long)mr_kernel_launches(ctx));
    mr_destroy(ctx);
    if (flat_best != 3) return 4;
    return best == 3 ? 0 : 2; /* the centre column is the right first move */
}
