"""Known-answer tests of the tree-side rows: the reference's OWN unit tests and formulas for the
selection strategies and the AMAF update, restated against BOTH the oracle (oracle/search.c) and the product's tree
driver (csrc/tree_search.cu, through its known-answer hooks mr_select_score / mr_kat_update_amaf -- host code, so these
run without a GPU).  The expected numbers are computed here, by hand, from the reference's formulas; neither
implementation is the other's yardstick.

Reference sources:
  mcts-tests/tests/ttt_strategies.rs:658-736  test_update_amaf_matches_by_movers_player_not_childs
  mcts-tests/tests/ttt_strategies.rs:367-392  test_tree_parallel_with_grave_picks_a_legal_action
  mcts/src/strategies/mcts/select/ucb.rs:40-61 (UCB1), :76-146 (UCB1-Tuned)
  mcts/src/strategies/mcts/select/rave.rs:60-75 (beta schedules), :196-230 (score)
  mcts/src/strategies/mcts/select/amaf.rs:58-79
  mcts/src/strategies/mcts/node.rs:126-159 (expected_score with virtual loss, value_estimate_unvisited)
"""
import math

import numpy as np
import pytest

import oracle_lib as o

C = math.sqrt(2.0)


def _product():
    from mcts_b200 import search as ps

    return ps


def _both(select, q_init=0, **kw):
    """(name, score function) for the oracle and for the product, same arguments."""
    ps = _product()
    ocfg = o.search_cfg(0, select=select, q_init=q_init, **{k: v for k, v in kw.items()})
    pkw = dict(kw)
    if "c" in pkw:
        pkw["exploration_constant"] = pkw.pop("c")
    pdesc = ps.make_search_desc(0, select=select, q_init=q_init, **pkw)

    def orc(player, cur, child, gn=0, gq=0):
        return o.select_score(ocfg, player, o.edge_stats(**cur), o.edge_stats(**child) if child is not None else None, gn, gq)

    def prod(player, cur, child, gn=0, gq=0):
        return ps.select_score(pdesc, player, ps.edge_stats(**cur), ps.edge_stats(**child) if child is not None else None, gn, gq)

    return [("oracle", orc), ("product", prod)]


def test_ucb1_hand_computed():
    """ucb.rs:40-52: exploit + c * sqrt(ln(N) / n_total); exploit = (score - vl) / (n + vl) (node.rs:126-133); the parent
    log is taken WITHOUT virtual loss (ucb.rs:34-37)."""
    cur = dict(visits=100, vloss=7, score=(10, -10), sumsq=(90, 90))
    child = dict(visits=20, vloss=0, score=(8, -8), sumsq=(18, 18))
    want0 = 8 / 20 + C * math.sqrt(math.log(100) / 20)
    want1 = -8 / 20 + C * math.sqrt(math.log(100) / 20)
    child_vl = dict(visits=20, vloss=4, score=(8, -8), sumsq=(18, 18))
    want_vl = (8 - 4) / (20 + 4) + C * math.sqrt(math.log(100) / 24)
    for name, f in _both(0):
        assert f(0, cur, child) == pytest.approx(want0, rel=0, abs=1e-15), name
        assert f(1, cur, child) == pytest.approx(want1, rel=0, abs=1e-15), name
        assert f(0, cur, child_vl) == pytest.approx(want_vl, rel=0, abs=1e-15), name


def test_unvisited_value_hand_computed():
    """node.rs:139-159 + ucb.rs:55-61: QInit::Parent = the parent's own expected score (10000 if the parent is unvisited),
    QInit::Infinity = 10000; UCB1 adds c * sqrt(ln N); Rave and Amaf add nothing (rave.rs:226-230, amaf.rs:76-79)."""
    cur = dict(visits=50, vloss=2, score=(-5, 5), sumsq=(45, 45))
    lnN = math.log(50)
    q_parent = (-5 - 2) / (50 + 2)
    for name, f in _both(0, q_init=0):
        assert f(0, cur, None) == pytest.approx(q_parent + C * math.sqrt(lnN), abs=1e-15), name
        assert f(0, dict(visits=0), None) == pytest.approx(10000.0 + C * math.sqrt(math.log(1.0)), abs=0), name
    for name, f in _both(0, q_init=1):
        assert f(0, cur, None) == pytest.approx(10000.0 + C * math.sqrt(lnN), abs=1e-11), name
    for name, f in _both(1, q_init=1):  # ucb1_tuned(c, q, VARIANCE_UPPER_BOUND, parent_log), ucb.rs:134-146
        assert f(0, cur, None) == pytest.approx(10000.0 + (lnN * 1.0 + C * math.sqrt(lnN)), abs=1e-11), name
    for sel in (2, 3):
        for name, f in _both(sel, q_init=0):
            assert f(1, cur, None) == pytest.approx((5 - 2) / 52, abs=1e-15), (name, sel)


def test_ucb1_tuned_hand_computed():
    """ucb.rs:90-132: exploit + (f * min(1, max(0, sumsq / n - exploit^2)) + c * sqrt(f)), f = ln(N) / n_total."""
    cur = dict(visits=400, score=(0, 0), sumsq=(300, 300))
    child = dict(visits=40, vloss=2, score=(12, -12), sumsq=(30, 30))
    n = 42.0
    exploit = (12 - 2) / n
    var = max(0.0, 30 / n - exploit * exploit)
    f_ = math.log(400) / n
    want = exploit + (f_ * min(1.0, var) + C * math.sqrt(f_))
    for name, f in _both(1):
        assert f(0, cur, child) == pytest.approx(want, abs=1e-15), name
    # a child whose sample variance clamps at 0 (all trials equal): sumsq / n - exploit^2 < 0 once virtual losses dilute sumsq
    child2 = dict(visits=10, vloss=10, score=(10, -10), sumsq=(10, 10))
    exploit2 = 0.0
    want2 = exploit2 + ((math.log(400) / 20) * min(1.0, max(0.0, 10 / 20 - 0.0)) + C * math.sqrt(math.log(400) / 20))
    for name, f in _both(1):
        assert f(0, cur, child2) == pytest.approx(want2, abs=1e-15), name


@pytest.mark.parametrize("schedule,param,bias,beta", [
    (0, 1000, 0.0, lambda n, an: math.sqrt(1000 / (3 * n + 1000))),                 # HandSelected { k } (rave.rs:63-66)
    (1, 300, 0.0, lambda n, an: max(0.0, 300 - n) / 300),                           # Threshold { rave } (rave.rs:70)
    (2, 0, 0.25, lambda n, an: an / (n + an + 4 * n * an * 0.25 * 0.25)),           # MinMSE { bias }   (rave.rs:68)
])
def test_grave_score_hand_computed(schedule, param, bias, beta):
    """rave.rs:196-224: (1 - beta) * mean + beta * amaf_q / amaf_n + c * sqrt(ln N / n)."""
    cur = dict(visits=250, score=(20, -20), sumsq=(200, 200))
    child = dict(visits=30, vloss=3, score=(-6, 6), sumsq=(28, 28))
    n = 33
    mean = (6 - 3) / n  # player 1's view
    gn, gq = 57, 19
    b = beta(n, gn)
    want = (1 - b) * mean + b * (gq / gn) + C * math.sqrt(math.log(250) / n)
    for name, f in _both(2, q_init=1, rave_schedule=schedule, rave_param=param, rave_bias=bias):
        assert f(1, cur, child, gn, gq) == pytest.approx(want, abs=1e-15), name
        # no AMAF data for this action at the reference node: amaf term 0 (rave.rs:186-193)
        b0 = beta(n, 0)
        assert f(1, cur, child, 0, 0) == pytest.approx((1 - b0) * mean + C * math.sqrt(math.log(250) / n), abs=1e-15), name


def test_amaf_select_hand_computed():
    """amaf.rs:58-74: alpha * amaf_score / max(1, amaf_n) + (1 - alpha) * ucb1."""
    cur = dict(visits=64, score=(4, -4), sumsq=(60, 60))
    child = dict(visits=16, vloss=1, score=(5, -5), sumsq=(15, 15), amaf_visits=40, amaf_score=(14, -14))
    ucb1 = (5 - 1) / 17 + C * math.sqrt(math.log(64) / 17)
    for alpha in (1.0, 0.5, 0.0):
        want = alpha * (14 / 40) + (1 - alpha) * ucb1
        for name, f in _both(3, q_init=1, amaf_alpha=alpha):
            assert f(0, cur, child) == pytest.approx(want, abs=1e-15), (name, alpha)
    empty = dict(visits=16, score=(5, -5), sumsq=(15, 15), amaf_visits=0, amaf_score=(0, 0))
    for name, f in _both(3, q_init=1, amaf_alpha=1.0):
        assert f(0, cur, empty) == 0.0, name  # 0 / max(1, 0)


def _connect4_after(cols):
    s = o.initial_state(o.CONNECT4)
    for c in cols:
        s = o.apply(o.CONNECT4, s, c)
    return s


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_update_amaf_matches_by_movers_player_not_childs(impl):
    """ttt_strategies.rs:658-736 on Connect4 (the reference states it on TicTacToe, which the rollout path does not
    carry; update_amaf is game-independent).  Root: White (player 1) to move; two created children, column 5 (the node
    being processed) and column 6 (its sibling).  A trace entry (column 6, player 1) -- the root's mover replaying the
    sibling's action later in the simulation -- must add one AMAF visit and the trial's utilities to the sibling's row;
    the same action by player 0 (the SIBLING's mover, who never had that choice at the root) must add nothing.
    Utilities are (-1, +1): the reference test's (0.25, 0.75) are not utilities any of the five games can produce, and
    the aggregated statistics this path carries are integer sums of +1 / 0 / -1."""
    root = _connect4_after([3])  # Black dropped in column 3: White to move
    assert int(root["turn"][0]) == 1
    na = 7

    def run(trace_player):
        amaf = np.zeros((2, na), dtype=o.STAT_DTYPE)
        amaf[trace_player, 6] = (1, [-1, 1][trace_player])  # one occurrence, scored with that player's utility
        if impl == "oracle":
            return o.kat_update_amaf(o.CONNECT4, root, [6], 5, amaf, [-1, 1], 1)
        import mcts_b200 as mb
        from mcts_b200 import search as ps

        return ps.kat_update_amaf(mb.Game.CONNECT4, root.view(mb.LEAF_DTYPE), [6], 5, amaf, [-1, 1], 1)

    rows = run(1)
    assert list(rows["action"]) == list(range(7))
    assert list(rows["created"]) == [0, 0, 0, 0, 0, 1, 1]
    sib = rows[6]
    assert sib["amaf_visits"] == 1, "O replaying the sibling's action later should count as an AMAF match"
    assert list(sib["amaf_score"]) == [-1, 1]
    assert rows[5]["amaf_visits"] == 0 and (rows["amaf_visits"][:5] == 0).all()
    rows = run(0)
    assert (rows["amaf_visits"] == 0).all(), "the sibling's own mover playing that action is not an AMAF match for the root's option"


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_update_amaf_counts_k_aggregated_trials(impl):
    """The aggregated form the GPU path feeds: k = 8 trials of one leaf, 5 of which contain (column 6, player 1) with
    utility sum +3 for player 1 => the sibling's row gets 5 visits, score (-3, +3) -- what 8 calls of the per-trial
    update_amaf add up to."""
    root = _connect4_after([3])
    amaf = np.zeros((2, 7), dtype=o.STAT_DTYPE)
    amaf[1, 6] = (5, 3)
    amaf[1, 2] = (4, -2)  # column 2 is not a created child: ignored
    if impl == "oracle":
        rows = o.kat_update_amaf(o.CONNECT4, root, [6], 5, amaf, [-2, 2], 8)
    else:
        import mcts_b200 as mb
        from mcts_b200 import search as ps

        rows = ps.kat_update_amaf(mb.Game.CONNECT4, root.view(mb.LEAF_DTYPE), [6], 5, amaf, [-2, 2], 8)
    assert rows[6]["amaf_visits"] == 5 and list(rows[6]["amaf_score"]) == [-3, 3]
    assert rows[2]["amaf_visits"] == 0 and rows[5]["amaf_visits"] == 0


def test_grave_with_decisive_mast_rollouts_picks_a_legal_action_oracle():
    """ttt_strategies.rs:367-392 (RaveMastDm = Compose<Rave, DecisiveMove<EpsilonGreedy<Mast>>>, 300 iterations,
    expand_threshold 0, use_transpositions, 4 tree workers) on Connect4: four descents in flight per batch stand in for
    the four workers.  The GPU driver runs the same case in tests/test_gpu_search.py."""
    root = o.initial_state(o.CONNECT4)
    best, children, _ = o.search(o.CONNECT4, root, iterations=300, batch_leaves=4, k=1, policy=o.DECISIVE_MAST_WIN, select=2,
                                 q_init=0, expand_threshold=0, use_transpositions=True, grave_threshold=700, seed=5)
    assert best in o.generate_actions(o.CONNECT4, root)
    assert int(children["visits"].sum()) == 300  # expand_threshold 0: every descent passes through a root child
