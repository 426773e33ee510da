"""The GPU-driven tree driver (mr_search) against the restated reference loop with CPU playouts
(oracle/search.c): same tree algorithm, same draws, bit-exact playouts => identical root statistics."""
import numpy as np
import pytest

import oracle_lib as o

pytestmark = pytest.mark.gpu
mb = pytest.importorskip("mcts_b200")

CASES = [
    # game, policy, iterations, batch, k, select, q_init, max_depth
    (o.CONNECT4, o.UNIFORM, 3000, 1, 1, 0, 0, 0),      # the reference's sequential loop
    (o.CONNECT4, o.UNIFORM, 4096, 64, 8, 0, 0, 0),     # batched leaf-parallel with virtual loss
    (o.CONNECT4, o.DECISIVE_WIN, 2048, 32, 4, 1, 1, 0),  # UCB1-Tuned + QInit::Infinity (Ucb1TunedDM)
    (o.BREAKTHROUGH, o.EPS_MAST, 2048, 32, 16, 0, 0, 200),  # MAST table fed by kernel delta + tree path
    (o.KNIGHTTHROUGH, o.EPS_MAST, 1024, 64, 8, 0, 0, 200),
    (o.CONNECT4, o.EPS_MAST, 2048, 32, 8, 0, 0, 0),    # Ucb1Mast (strategy.rs:98-111) on the games without move kinds
    (o.OTHELLO, o.EPS_MAST, 1024, 16, 8, 0, 0, 0),
    (o.HEX11, o.EPS_MAST, 1024, 32, 8, 0, 0, 0),       # the per-ply Hex kernel (hex_step.cuh) behind the driver
    (o.OTHELLO, o.DECISIVE_WIN, 1024, 32, 4, 1, 1, 0),
    (o.HEX11, o.UNIFORM, 1024, 64, 4, 0, 0, 0),
]


@pytest.mark.parametrize("game,policy,iters,batch,k,select,q_init,max_depth", CASES)
def test_search_root_statistics_bit_exact(game, policy, iters, batch, k, select, q_init, max_depth):
    strat = {o.UNIFORM: mb.Uniform(), o.EPS_MAST: mb.EpsilonGreedyMast(0.1), o.DECISIVE_WIN: mb.DecisiveMoveWin()}[policy]
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=max_depth, select=select, q_init=q_init, seed=77) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=policy, select=select,
                                     q_init=q_init, max_depth=max_depth, seed=77, threads=o.host_threads())
    assert got.playouts == iters * k
    assert len(got.children) == len(children)
    assert (got.children["action"] == children["action"]).all()
    assert (got.children["visits"] == children["visits"]).all()
    assert (got.children["score"] == children["score"]).all()
    assert got.best_action == best
    assert got.playout_plies == plies
    # only the very first descent stops at the unvisited root: the expansion test counts the trials in flight, so the
    # second descent of the first batch already expands the root (csrc/tree_search.cu header)
    assert int(got.children["visits"].sum()) == (iters - 1) * k


@pytest.mark.parametrize("game,policy,depth", [(o.CONNECT4, o.UNIFORM, 0), (o.BREAKTHROUGH, o.EPS_MAST, 200)])
def test_pipelined_search_bit_exact(game, policy, depth):
    """Double-buffered driver: batch N+1 gathered and submitted (slot 1) while batch N (slot 0) is in flight."""
    strat = {o.UNIFORM: mb.Uniform(), o.EPS_MAST: mb.EpsilonGreedyMast(0.1)}[policy]
    root = o.initial_state(game)
    iters, batch, k = 4096 + 17, 64, 8  # a ragged last batch
    with mb.GpuSearch(mb.Game(game), strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=depth, pipelined=True, seed=21) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=policy, max_depth=depth,
                                     seed=21, threads=o.host_threads(), pipelined=True)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies and got.playouts == iters * k


def test_search_with_transposition_table_bit_exact():
    """BASELINE.json config 4: Othello, decisive-move rollouts, transposition table on the host."""
    game, iters, batch, k = o.OTHELLO, 4096, 64, 8
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game.OTHELLO, mb.DecisiveMoveWin(), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      select=1, q_init=1, use_transpositions=True, seed=9) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
        plain_nodes = None
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.DECISIVE_WIN, select=1,
                                     q_init=1, seed=9, threads=o.host_threads(), use_transpositions=True)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    with mb.GpuSearch(mb.Game.OTHELLO, mb.DecisiveMoveWin(), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      select=1, q_init=1, use_transpositions=False, seed=9) as s:
        plain_nodes = s.choose_action(root.view(mb.LEAF_DTYPE)).nodes
    assert got.nodes < plain_nodes  # transpositions were actually shared


@pytest.mark.parametrize("game,iters,batch,k,thr,sched,param", [
    (o.HEX11, 2048, 32, 8, 50, 0, 1000),   # BASELINE.json config 3: Hex 11x11, GRAVE reading in-kernel AMAF tables
    (o.HEX11, 1024, 16, 4, 0, 1, 300),     # threshold 0 = plain RAVE, Threshold schedule
    (o.CONNECT4, 2048, 32, 8, 100, 0, 1000),
])
def test_grave_search_bit_exact(game, iters, batch, k, thr, sched, param):
    from mcts_b200.search import SELECT_GRAVE, QINIT_INFINITY

    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.Uniform(), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=thr, rave_schedule=sched, rave_param=param,
                      seed=13) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, select=2, q_init=1, seed=13,
                                     threads=o.host_threads(), grave_threshold=thr, rave_schedule=sched, rave_param=param)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    assert (got.children["visits"] > 0).sum() > 3  # the AMAF term actually spreads the search


@pytest.mark.parametrize("game,policy,iters,batch,k,alpha", [
    (o.HEX11, o.UNIFORM, 2048, 32, 8, 1.0),          # strategy.rs:205-217 `Amaf`: pure AMAF (alpha 1, the default)
    (o.CONNECT4, o.UNIFORM, 4096, 32, 4, 0.5),       # alpha blends AMAF with UCB1 (amaf.rs:70-74)
    (o.BREAKTHROUGH, o.EPS_MAST, 2048, 64, 8, 0.7),  # strategy.rs:219-231 `AmafMast`
    (o.OTHELLO, o.UNIFORM, 1024, 1, 1, 1.0),         # the reference's sequential loop (batch 1, k 1)
])
def test_amaf_search_bit_exact(game, policy, iters, batch, k, alpha):
    from mcts_b200.search import SELECT_AMAF, QINIT_INFINITY

    root = o.initial_state(game)
    strat = mb.EpsilonGreedyMast(0.1) if policy == o.EPS_MAST else mb.Uniform()
    with mb.GpuSearch(mb.Game(game), strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      select=SELECT_AMAF, q_init=QINIT_INFINITY, amaf_alpha=alpha, max_playout_depth=200, seed=21) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=policy, select=3, q_init=1,
                                     seed=21, threads=o.host_threads(), amaf_alpha=alpha, max_depth=200)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    assert (got.children["visits"] > 0).sum() > 3


@pytest.mark.parametrize("game,iters,batch,k,pipelined", [
    (o.CONNECT4, 4096, 32, 4, False),       # strategy.rs:130-143 `Ucb1Lgr`: UCB1 + EpsilonGreedy<Lgr<Uniform>>
    (o.BREAKTHROUGH, 2048, 64, 8, False),
    (o.KNIGHTTHROUGH, 1024, 16, 16, True),  # pipelined: batch N+1 plays with the table as of batch N-1
    (o.OTHELLO, 1024, 1, 1, False),         # the reference's sequential loop: every trial sees every earlier write
    (o.HEX11, 1024, 32, 4, False),
])
def test_lgr_search_bit_exact(game, iters, batch, k, pipelined):
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.EpsilonGreedyLgr(0.1), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=200, pipelined=pipelined, seed=33) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_LGR, eps=0.1, seed=33,
                                     threads=o.host_threads(), max_depth=200, pipelined=pipelined)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    # the table changes the playouts: same search with plain uniform rollouts over the same pair-draw stream differs
    _, uni_children, uni_plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_LGR, eps=1.0,
                                          seed=33, threads=o.host_threads(), max_depth=200, pipelined=pipelined)
    assert uni_plies != plies


@pytest.mark.parametrize("game,iters,batch,k,thr,pipelined", [
    (o.CONNECT4, 2048, 32, 8, 5, False),
    (o.BREAKTHROUGH, 1024, 16, 16, 2, True),
    (o.KNIGHTTHROUGH, 256, 8, 16, 5, False),
    (o.OTHELLO, 768, 1, 1, 1, False),  # the reference's sequential loop: every trial sees every earlier write
    (o.HEX11, 1024, 32, 4, 5, False),
])
def test_nst_search_bit_exact(game, iters, batch, k, thr, pipelined):
    """The `Ucb1Nst` strategy (strategy.rs:113-126) end to end: UCB1 over EpsilonGreedy<Nst> rollouts, the bigram table fed
    by the kernel's delta and the tree-path pairs (backprop.rs:885-899), the unigram table as for MAST."""
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.EpsilonGreedyNst(0.1, thr), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=200, pipelined=pipelined, seed=35) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_NST, eps=0.1, seed=35,
                                     threads=o.host_threads(), max_depth=200, pipelined=pipelined, nst_threshold=thr)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    # the bigram table changes the playouts: the same search with plain MAST rollouts differs
    _, _, mast_plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_MAST, eps=0.1, seed=35,
                                threads=o.host_threads(), max_depth=200, pipelined=pipelined)
    assert mast_plies != plies


@pytest.mark.parametrize("game,iters,batch,k,pipelined", [
    (o.CONNECT4, 2048, 32, 8, False),
    (o.BREAKTHROUGH, 1024, 16, 16, True),
    (o.KNIGHTTHROUGH, 256, 8, 16, False),
    (o.OTHELLO, 512, 1, 1, False),  # the reference's sequential loop: every trial sees every earlier insert and removal
    (o.HEX11, 512, 16, 8, False),
])
def test_lgrf2_search_bit_exact(game, iters, batch, k, pipelined):
    """The `Ucb1Lgr2` strategy (strategy.rs:143-156) end to end: the 2-ply table with forgetting, merged from the kernel's
    insert / replay passes and the tree-path windows, against the oracle's literal trial-by-trial table updates."""
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.EpsilonGreedyLgr2(0.1), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=200, pipelined=pipelined, seed=37) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_LGR2, eps=0.1, seed=37,
                                     threads=o.host_threads(), max_depth=200, pipelined=pipelined)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    # the 2-ply table changes the playouts: the same search with LGR-1 rollouts differs
    _, _, lgr_plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_LGR, eps=0.1, seed=37,
                               threads=o.host_threads(), max_depth=200, pipelined=pipelined)
    assert lgr_plies != plies


def test_lgrf2_mast_search_bit_exact():
    """`Ucb1Lgr2Mast` (strategy.rs:158-171): LGRF-2 -> LGR-1 -> greedy MAST, three tables maintained by the driver."""
    game, iters, batch, k = o.BREAKTHROUGH, 1024, 16, 16
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.EpsilonGreedyLgr2Mast(0.1), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=200, seed=39) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_LGR2_MAST, eps=0.1, seed=39,
                                     threads=o.host_threads(), max_depth=200)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    _, _, plain = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_LGR2, eps=0.1, seed=39,
                           threads=o.host_threads(), max_depth=200)
    assert plain != plies


def test_rave_family_search_bit_exact():
    """mcts-tune's `rave` family (family_catalog.rs:483-514): Rave select over the in-kernel AMAF tables with
    DecisiveMove<WinLoss, EpsilonGreedy<Mast>> rollouts -- AMAF tables, MAST delta and decisive scan in one kernel."""
    from mcts_b200.search import SELECT_GRAVE, QINIT_INFINITY

    game, iters, batch, k = o.BREAKTHROUGH, 2048, 32, 8
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.DecisiveMoveMast(0.1, "win_loss"), max_iterations=iters, batch_leaves=batch,
                      num_rollouts_per_leaf=k, select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=100,
                      max_playout_depth=200, seed=17) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.DECISIVE_MAST_WINLOSS,
                                     select=2, q_init=1, seed=17, threads=o.host_threads(), grave_threshold=100, max_depth=200)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies


def test_flat_monte_carlo_one_batch():
    from mcts_b200.search import flat_monte_carlo

    game = o.CONNECT4
    s = o.initial_state(game)
    for c in [3, 3, 2, 4]:
        s = o.apply(game, s, c)
    with mb.GpuRollout(mb.Game.CONNECT4, seed=4) as g:
        action, wins = flat_monte_carlo(g, s.view(mb.LEAF_DTYPE), samples_per_move=500, max_rollout_depth=100, r=37)
    acts = o.generate_actions(game, s)
    leaves = np.concatenate([o.apply(game, s, a) for a in acts])
    ref = o.playout_batch(game, leaves, 500, seed=4, max_depth=99)["results"]
    assert (wins == ref["wins"][:, 0]).all()
    assert action in acts and wins[acts.index(action)] == wins.max()


def test_flat_monte_carlo_ucb1_and_error_paths():
    """flat_mc.rs:135-149: with the UCB1 value w / n + c ln(n) / n every child has the same n, so the pick for a given draw
    is the pick by wins; a terminal root is refused (the reference panics), as is a zero sample count."""
    from mcts_b200.search import flat_monte_carlo

    game = o.BREAKTHROUGH
    s = o.random_positions(game, 4, 12, 1)
    with mb.GpuRollout(mb.Game.BREAKTHROUGH, seed=9) as g:
        for r in (0, 5, 123, 16 * 7 + 3):
            a0, w0 = flat_monte_carlo(g, s.view(mb.LEAF_DTYPE), samples_per_move=200, max_rollout_depth=60, r=r)
            a1, w1 = flat_monte_carlo(g, s.view(mb.LEAF_DTYPE), samples_per_move=200, max_rollout_depth=60, ucb1=0.7, r=r)
            assert a0 == a1 and (w0 == w1).all()
        acts = o.generate_actions(game, s)
        leaves = np.concatenate([o.apply(game, s, a) for a in acts])
        ref = o.playout_batch(game, leaves, 200, seed=9, max_depth=59)["results"]
        assert (w0 == ref["wins"][:, int(s["turn"][0])]).all()
        won = o.make_state([1 << 3, 1 << 60], turn=0, flag=1)  # a finished game (the winner flag set)
        with pytest.raises(ValueError):
            flat_monte_carlo(g, won.view(mb.LEAF_DTYPE), samples_per_move=10)
        with pytest.raises(ValueError):
            flat_monte_carlo(g, s.view(mb.LEAF_DTYPE), samples_per_move=0)


def test_search_prefers_the_centre_in_connect4():
    with mb.GpuSearch(mb.Game.CONNECT4, max_iterations=20000, batch_leaves=128, num_rollouts_per_leaf=8, seed=3) as s:
        r = s.choose_action(mb.initial_state(mb.Game.CONNECT4))
    assert r.best_action == 3
    assert r.children["visits"].argmax() == 3


# ---- the reference's leaf-parallel tests (mcts-tests/tests/ttt_strategies.rs:86-159), stated there on TicTacToe ----

def _legal_root_actions(game):
    return set(mb.generate_actions(mb.Game(game), mb.initial_state(mb.Game(game))))


def test_leaf_parallel_picks_a_legal_action():  # ttt_strategies.rs:86-106: 200 iterations, num_rollouts_per_leaf = 4
    with mb.GpuSearch(mb.Game.CONNECT4, mb.Uniform(), max_iterations=200, batch_leaves=1, num_rollouts_per_leaf=4, seed=1) as s:
        r = s.choose_action(mb.initial_state(mb.Game.CONNECT4))
    assert r.best_action in _legal_root_actions(o.CONNECT4) and r.playouts == 800


def test_num_rollouts_per_leaf_one_is_deterministic_given_a_seed():  # ttt_strategies.rs:108-130
    picks = []
    for _ in range(2):
        with mb.GpuSearch(mb.Game.CONNECT4, mb.Uniform(), max_iterations=200, batch_leaves=1, num_rollouts_per_leaf=1, seed=42) as s:
            r = s.choose_action(mb.initial_state(mb.Game.CONNECT4))
            picks.append((r.best_action, r.children.tobytes()))
    assert picks[0] == picks[1]


@pytest.mark.parametrize("batch", [1, 16])
def test_leaf_parallel_virtual_loss_balances_across_many_iterations(batch):  # ttt_strategies.rs:132-159
    """expand_threshold = 0 makes every descent run several levels deep, K = 3 adds two extra virtual losses per edge and
    removes them over the three aggregated backprops: the root statistics must still equal the oracle loop's, which
    applies the reference's per-trial accounting, and every playout must be accounted for."""
    iters, k = 500, 3
    root = o.initial_state(o.CONNECT4)
    with mb.GpuSearch(mb.Game.CONNECT4, mb.Uniform(), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      expand_threshold=0, seed=9) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(o.CONNECT4, root, iterations=iters, batch_leaves=batch, k=k, expand_threshold=0, seed=9)
    assert got.best_action in _legal_root_actions(o.CONNECT4)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    assert int(got.children["visits"].sum()) == iters * k  # with threshold 0 every descent passes through a root child


@pytest.mark.parametrize("game,policy,depth,batch,k,max_depth", [
    (o.CONNECT4, o.UNIFORM, 3, 64, 8, 0),
    (o.CONNECT4, o.UNIFORM, 4, 16, 64, 0),
    (o.BREAKTHROUGH, o.EPS_MAST, 4, 32, 16, 200),  # MAST table as of batch n - 3
    (o.HEX11, o.UNIFORM, 3, 32, 8, 0),
])
def test_deep_pipeline_search_bit_exact(game, policy, depth, batch, k, max_depth):
    """pipeline_depth batches in flight over the ABI's four slots (each with its own stream): batch n is gathered and
    submitted before the results of batches n - depth + 1 .. n - 1 are applied; deterministic, mirrored by the oracle."""
    strat = {o.UNIFORM: mb.Uniform(), o.EPS_MAST: mb.EpsilonGreedyMast(0.1)}[policy]
    root = o.initial_state(game)
    iters = 2048 + 13  # a ragged last batch
    with mb.GpuSearch(mb.Game(game), strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=max_depth, pipeline_depth=depth, seed=23) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=policy, max_depth=max_depth,
                                     seed=23, threads=o.host_threads(), pipeline_depth=depth)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies and got.playouts == iters * k
    assert int(got.children["visits"].sum()) == (iters - 1) * k


def test_grave_min_mse_schedule_bit_exact():
    """RaveSchedule::MinMSE { bias } (select/rave.rs:68): beta depends on the reference node's AMAF count."""
    from mcts_b200.search import QINIT_INFINITY, RAVE_MIN_MSE, SELECT_GRAVE

    game, iters, batch, k = o.HEX11, 1024, 32, 8
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.Uniform(), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=50, rave_schedule=RAVE_MIN_MSE, rave_bias=0.1,
                      seed=29) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, select=2, q_init=1, seed=29,
                                     threads=o.host_threads(), grave_threshold=50, rave_schedule=2, rave_bias=0.1)
    assert (got.children["visits"] == children["visits"]).all() and (got.children["score"] == children["score"]).all()
    assert got.best_action == best and got.playout_plies == plies
    # and the schedule matters: HandSelected from the same seed walks another tree
    _, other, _ = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, select=2, q_init=1, seed=29,
                           threads=o.host_threads(), grave_threshold=50, rave_schedule=0, rave_param=1000)
    assert (other["visits"] != children["visits"]).any()


def test_grave_with_decisive_mast_rollouts_picks_a_legal_action():
    """ttt_strategies.rs:367-392 (RaveMastDm, 300 iterations, expand_threshold 0, use_transpositions, 4 tree workers)
    on Connect4 through the GPU driver; tests/test_search_kats.py holds the oracle's run of the same case."""
    from mcts_b200.search import SELECT_GRAVE

    root = o.initial_state(o.CONNECT4)
    with mb.GpuSearch(mb.Game.CONNECT4, mb.DecisiveMoveMast(0.1, "win"), max_iterations=300, batch_leaves=4, num_rollouts_per_leaf=1,
                      select=SELECT_GRAVE, expand_threshold=0, use_transpositions=True, seed=5) as s:
        got = s.choose_action(root.view(mb.LEAF_DTYPE))
    best, children, plies = o.search(o.CONNECT4, root, iterations=300, batch_leaves=4, k=1, policy=o.DECISIVE_MAST_WIN, select=2,
                                     q_init=0, expand_threshold=0, use_transpositions=True, grave_threshold=700, seed=5)
    assert got.best_action in _legal_root_actions(o.CONNECT4)
    assert got.best_action == best and (got.children["visits"] == children["visits"]).all() and got.playout_plies == plies
    assert int(got.children["visits"].sum()) == 300


def _searcher_seed(seed, i):
    """mr_search_root_parallel's seed of searcher i: words 0..1 of Philox(seed; ctr = (i, 0, 0, 3))."""
    import ctypes as C

    ctr = (C.c_uint32 * 4)(i, 0, 0, 3)
    key = (C.c_uint32 * 2)(seed & 0xFFFFFFFF, seed >> 32)
    out = (C.c_uint32 * 4)()
    o.lib().orc_philox4x32_10(ctr, key, out)
    return (out[1] << 32) | out[0]


def test_root_parallel_threads_equal_the_sum_of_independent_searches():
    """choose_action_root_parallel (search/parallel.rs:56-115) with one host thread and one rollout context per
    searcher, all on the same GPU: the merged root-child totals are exactly the sum of the searchers' own totals, each
    searcher being the single-tree search with its derived seed, and the pick has the most merged visits."""
    game, iters, batch, k, n_thr, seed = o.BREAKTHROUGH, 512, 32, 16, 3, 77
    root = o.initial_state(game)
    with mb.GpuSearch(mb.Game(game), mb.EpsilonGreedyMast(0.1), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                      max_playout_depth=200, pipeline_depth=2, seed=seed, num_threads=n_thr) as s:
        merged = s.choose_action(root.view(mb.LEAF_DTYPE))
    visits = np.zeros(len(merged.children), dtype=np.int64)
    score = np.zeros((len(merged.children), 2), dtype=np.int64)
    for i in range(n_thr):
        _, ch, _ = o.search(game, root, iterations=iters, batch_leaves=batch, k=k, policy=o.EPS_MAST, eps=0.1, max_depth=200,
                            seed=_searcher_seed(seed, i), threads=o.host_threads(), pipeline_depth=2)
        visits += ch["visits"]
        score += ch["score"]
    assert (merged.children["visits"] == visits).all() and (merged.children["score"] == score).all()
    assert merged.playouts == n_thr * iters * k
    assert visits[list(merged.children["action"]).index(merged.best_action)] == visits.max()


def test_search_argument_errors_are_codes_not_crashes():
    """mr_search / mr_search_root_parallel validate their descriptors: bad arguments come back as MR_ERR_INVALID."""
    import ctypes as C

    from mcts_b200 import _lib
    from mcts_b200._lib import MAX_ROOT_CHILDREN, ROOT_CHILD_DTYPE, SearchStats, ptr

    L = _lib.load()
    root = mb.initial_state(mb.Game.CONNECT4)
    children = np.zeros(MAX_ROOT_CHILDREN, dtype=ROOT_CHILD_DTYPE)
    best, n, st = C.c_uint16(), C.c_uint32(), SearchStats()
    with mb.GpuSearch(mb.Game.CONNECT4, mb.Uniform(), max_iterations=64, batch_leaves=8, num_rollouts_per_leaf=4, seed=1) as s:
        for field, value in [("pipeline_depth", 5), ("select", 9), ("rave_schedule", 3), ("batch_leaves", 0), ("max_iterations", 0), ("game", 7)]:
            saved = getattr(s.desc, field)
            setattr(s.desc, field, value)
            rc = L.mr_search(s.rollout._h, C.byref(s.desc), ptr(root), C.byref(best), ptr(children), C.byref(n), C.byref(st))
            setattr(s.desc, field, saved)
            assert rc == 1, (field, rc)
        handles = (C.c_void_p * 2)(s.rollout._h, None)
        sts = (SearchStats * 2)()
        assert L.mr_search_root_parallel(handles, 2, C.byref(s.desc), ptr(root), C.byref(best), ptr(children), C.byref(n), sts) == 1
        assert L.mr_search_root_parallel(handles, 0, C.byref(s.desc), ptr(root), C.byref(best), ptr(children), C.byref(n), sts) == 1
        # and the context is still usable afterwards
        r = s.choose_action(root)
        assert r.playouts == 64 * 4 and r.best_action in _legal_root_actions(o.CONNECT4)
