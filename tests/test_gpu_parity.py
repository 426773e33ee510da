"""GPU parity: the CUDA path (through the C ABI, libmcts_b200.so) against the CPU oracle, bit-exact.

Every playout is a pure function of (leaf state, seed, leaf id, playout id), so parity is checked per
playout: identical move trace, identical end record (depth, outcome, end type), identical final state,
and identical aggregated statistics.  Integer work: the bar is bit-exact.
"""
import numpy as np
import pytest

import oracle_lib as o

pytestmark = pytest.mark.gpu

mb = pytest.importorskip("mcts_b200")

GAMES = [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11]
MAX_TRACE = {o.CONNECT4: 42, o.BREAKTHROUGH: 256, o.KNIGHTTHROUGH: 300, o.OTHELLO: 128, o.HEX11: 121}
MID_PLIES = {o.CONNECT4: [0, 8, 16], o.BREAKTHROUGH: [0, 10, 30], o.KNIGHTTHROUGH: [0, 10, 30], o.OTHELLO: [0, 10, 50], o.HEX11: [0, 10, 60]}


def _leaves(game, n_per=6, seed=11):
    parts = [o.random_positions(game, seed + p, p, n_per if p else 1) for p in MID_PLIES[game]]
    return np.concatenate(parts)


def _random_mast(game, seed, density=0.6):
    rng = np.random.default_rng(seed)
    na = o.num_actions(game)
    mast = np.zeros((2, na), dtype=o.STAT_DTYPE)
    visited = rng.random((2, na)) < density
    visits = rng.integers(1, 40, size=(2, na))
    score = (rng.integers(0, 41, size=(2, na)) * 2 - 40) * visits // 40
    mast["visits"] = np.where(visited, visits, 0)
    mast["score"] = np.where(visited, np.clip(score, -visits, visits), 0)
    return mast


def _strategy(policy, eps=0.1):
    return {o.UNIFORM: mb.Uniform(), o.EPS_MAST: mb.EpsilonGreedyMast(eps), o.DECISIVE_WIN: mb.DecisiveMoveWin(),
            o.DECISIVE_WINLOSS: mb.DecisiveMove("win_loss"), o.DECISIVE_WINLOSSDRAW: mb.DecisiveMove("win_loss_draw"),
            o.DECISIVE_ANTI: mb.DecisiveMove("anti_decisive"), o.DECISIVE_MAST_WIN: mb.DecisiveMoveMast(eps, "win"),
            o.DECISIVE_MAST_WINLOSS: mb.DecisiveMoveMast(eps, "win_loss"), o.EPS_NST: mb.EpsilonGreedyNst(eps)}[policy]


def _compare(game, policy, leaves, k, *, seed, max_depth=0, mast=None, eps=0.1, leaf_id_base=0, amaf=False, delta=False, devices=1):
    mt = MAX_TRACE[game] if not max_depth else max_depth
    ref = o.playout_batch(
        game, leaves, k, policy=policy, eps=eps, max_depth=max_depth, seed=seed, leaf_id_base=leaf_id_base, mast=mast,
        want_trace=True, max_trace=mt, want_final=True, want_amaf=amaf, want_mast_delta=delta, threads=o.host_threads(),
    )
    with mb.GpuRollout(mb.Game(game), _strategy(policy, eps), seed=seed, max_playout_depth=max_depth, devices=devices) as g:
        got = g.simulate_many(
            leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_trace=True, want_final=True, max_trace=mt,
            want_amaf=amaf, want_mast_delta=delta, leaf_id_base=leaf_id_base,
        )
    name = o.GAME_NAMES[game]
    assert (got.records["depth"] == ref["records"]["depth"]).all(), name
    assert (got.records["outcome"] == ref["records"]["outcome"]).all(), name
    assert (got.records["end_type"] == ref["records"]["end_type"]).all(), name
    assert (got.trace == ref["trace"]).all(), name
    assert got.final.tobytes() == ref["final"].tobytes(), name
    assert got.results.tobytes() == ref["results"].tobytes(), name
    if amaf:
        assert got.amaf.tobytes() == ref["amaf"].tobytes(), name
    if delta:
        assert got.mast_delta.tobytes() == ref["mast_delta"].tobytes(), name
    return got, ref


@pytest.mark.parametrize("game", GAMES)
def test_uniform_playouts_bit_exact(game):
    leaves = _leaves(game)
    got, _ = _compare(game, o.UNIFORM, leaves, 96, seed=0xC0FFEE + game)
    r = got.results
    assert ((r["wins"][:, 0] + r["wins"][:, 1] + r["draws"]) == 96).all()


@pytest.mark.parametrize("game", GAMES)
def test_white_to_move_leaves(game):
    """Leaves after an odd number of plies: the second player (White / P1) moves first in every playout."""
    odd = {o.CONNECT4: [1, 7, 15], o.BREAKTHROUGH: [1, 11, 31], o.KNIGHTTHROUGH: [1, 9, 21], o.OTHELLO: [1, 11, 49], o.HEX11: [1, 31, 75]}[game]
    leaves = np.concatenate([o.random_positions(game, 40 + p, p, 4) for p in odd])
    assert (leaves["turn"] == 1).all()
    _compare(game, o.UNIFORM, leaves, 64, seed=0xBEEF + game, amaf=game in (o.HEX11, o.CONNECT4))
    if game in (o.BREAKTHROUGH, o.KNIGHTTHROUGH):
        _compare(game, o.EPS_MAST, leaves, 64, seed=3, mast=_random_mast(game, 5, 0.7), max_depth=200, delta=True)
        _compare(game, o.DECISIVE_WIN, leaves, 64, seed=4)


@pytest.mark.parametrize("game", GAMES)
def test_uniform_ragged_k_and_leaf_id_base(game):
    # k not a multiple of 32, a single leaf, a non-zero leaf id base
    leaves = _leaves(game, n_per=1)[-1:]
    _compare(game, o.UNIFORM, leaves, 37, seed=5, leaf_id_base=1000)
    _compare(game, o.UNIFORM, leaves, 1, seed=6)


@pytest.mark.parametrize("game", [o.KNIGHTTHROUGH, o.BREAKTHROUGH, o.CONNECT4])
def test_depth_cutoff(game):
    leaves = _leaves(game, n_per=3)
    got, _ = _compare(game, o.UNIFORM, leaves, 40, seed=77, max_depth=9)
    assert got.results["cutoffs"].sum() > 0


def test_terminal_and_stuck_leaves():
    # a won Connect4 leaf, a full-board draw-less leaf is impossible to script cheaply; use won + stuck
    won = o.initial_state(o.CONNECT4)
    for c in [0, 1, 0, 1, 0, 1, 0]:
        won = o.apply(o.CONNECT4, won, c)
    mixed = np.concatenate([won, o.initial_state(o.CONNECT4)])
    got, _ = _compare(o.CONNECT4, o.UNIFORM, mixed, 33, seed=1)
    assert got.results[0]["wins"][0] == 33 and got.results[0]["sum_depth"] == 0
    # Breakthrough: side to move has no action but the game is not terminal => draw with depth 0
    stuck = o.make_state([(1 << 8) | (1 << 1), 1 << 0], turn=0, flag=0)
    got, _ = _compare(o.BREAKTHROUGH, o.UNIFORM, stuck, 5, seed=2)
    assert got.results[0]["draws"] == 5
    # Othello: double-pass terminal leaf and a leaf whose mover must pass
    dp = o.make_state([1 << 0, 1 << 63], turn=1, flag=1)
    _compare(o.OTHELLO, o.UNIFORM, dp, 4, seed=3)
    must_pass = o.make_state([1 << 0, 1 << 63], turn=0, flag=0)
    _compare(o.OTHELLO, o.UNIFORM, must_pass, 4, seed=3)


@pytest.mark.parametrize("game", [o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.CONNECT4, o.OTHELLO, o.HEX11])
@pytest.mark.parametrize("eps", [0.1, 0.0, 1.0])
def test_eps_mast_bit_exact(game, eps):
    leaves = _leaves(game, n_per=3)
    for density in (0.0, 0.5, 1.0):  # empty table (all ties), mixed, fully visited
        mast = _random_mast(game, 42 + int(density * 10), density)
        _compare(game, o.EPS_MAST, leaves, 40, seed=9 + int(eps * 10), mast=mast, eps=eps, max_depth=200, delta=True)


def test_mast_delta_feeds_next_batch():
    """Two batches with the table updated in between (backprop.rs:857-870), GPU and oracle in lockstep."""
    game = o.BREAKTHROUGH
    leaves = _leaves(game, n_per=2)
    mast = np.zeros((2, o.num_actions(game)), dtype=o.STAT_DTYPE)
    for it in range(3):
        got, ref = _compare(game, o.EPS_MAST, leaves, 64, seed=100 + it, mast=mast, max_depth=200, delta=True)
        mast["visits"] += got.mast_delta["visits"]
        mast["score"] += got.mast_delta["score"]
    assert mast["visits"].sum() > 0


@pytest.mark.parametrize("game", [o.HEX11, o.CONNECT4, o.OTHELLO])
def test_amaf_tables_bit_exact(game):
    leaves = _leaves(game, n_per=2)
    got, _ = _compare(game, o.UNIFORM, leaves, 48, seed=31, amaf=True)
    # occurrences == plies: every ply contributes one (action, player) visit
    assert got.amaf["visits"].sum() == got.results["sum_depth"].sum()


@pytest.mark.parametrize("game", [o.OTHELLO, o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.HEX11])
def test_decisive_win_bit_exact(game):
    leaves = _leaves(game, n_per=4)
    if game == o.OTHELLO:  # late positions, where the 64th-disc case can occur
        leaves = np.concatenate([leaves, o.random_positions(game, 5, 56, 8)])
    _compare(game, o.DECISIVE_WIN, leaves, 64, seed=17)


@pytest.mark.parametrize("game", [o.OTHELLO, o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.HEX11])
@pytest.mark.parametrize("mode", [o.DECISIVE_WINLOSS, o.DECISIVE_WINLOSSDRAW])
def test_decisive_other_modes_bit_exact(game, mode):
    """WinLoss / WinLossDraw restated literally in the oracle; the product routes them to the Win kernel."""
    leaves = _leaves(game, n_per=3)
    if game == o.OTHELLO:
        leaves = np.concatenate([leaves, o.random_positions(game, 6, 57, 8)])
    if game == o.CONNECT4:  # nearly full boards: the drawing last stone
        leaves = np.concatenate([leaves, o.random_positions(game, 7, 38, 8)])
    _compare(game, mode, leaves, 48, seed=23)


@pytest.mark.parametrize("game", [o.OTHELLO, o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.HEX11])
def test_anti_decisive_bit_exact(game):
    """DecisiveMoveMode::AntiDecisive (simulate.rs:689-735): the oracle runs the reference's two literal passes
    (apply + terminal_status per action, then per reply); the kernels use closed forms per game.  The mode takes
    the FIRST safe action, so playouts of one leaf differ little: coverage comes from many leaves."""
    plies = {o.CONNECT4: range(0, 40, 3), o.BREAKTHROUGH: range(0, 60, 4), o.KNIGHTTHROUGH: range(0, 40, 3),
             o.OTHELLO: list(range(0, 56, 5)) + [56, 57, 58, 59, 60], o.HEX11: range(0, 110, 9)}[game]
    leaves = np.concatenate([o.random_positions(game, 300 + p, p, 24 if p else 1) for p in plies])
    # (Knightthrough playouts have no depth bound of their own: statistics tables need max_playout_depth)
    got, _ = _compare(game, o.DECISIVE_ANTI, leaves, 8, seed=31, amaf=game != o.OTHELLO, max_depth=300 if game == o.KNIGHTTHROUGH else 0)
    r = got.results
    assert ((r["wins"][:, 0] + r["wins"][:, 1] + r["draws"]) == 8).all()
    # white-to-move leaves as well (odd ply counts are in the ranges above) and a depth cutoff
    _compare(game, o.DECISIVE_ANTI, leaves[::5], 4, seed=32, max_depth=9)


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11])
@pytest.mark.parametrize("policy", [o.DECISIVE_MAST_WIN, o.DECISIVE_MAST_WINLOSS])
def test_decisive_move_over_mast_bit_exact(game, policy):
    """DecisiveMove<EpsilonGreedy<Mast>> (the `ucb1_tuned_dm_mast` and `rave` families): the decisive scan, then the
    epsilon-greedy MAST pick; traces, records, final states, results and the MAST delta."""
    leaves = _leaves(game, n_per=4)
    if game == o.CONNECT4:
        leaves = np.concatenate([leaves, o.random_positions(game, 9, 30, 8)])
    for density, eps in ((0.0, 0.1), (0.6, 0.1), (1.0, 0.0)):
        mast = _random_mast(game, 7 + int(density * 10), density)
        got, ref = _compare(game, policy, leaves, 48, seed=41, mast=mast, eps=eps, max_depth=200, delta=True)
    if game != o.OTHELLO:  # the decisive scan changes picks: not the plain MAST playouts
        plain = o.playout_batch(game, leaves, 48, policy=o.EPS_MAST, eps=0.0, max_depth=200, seed=41, mast=mast)
        assert plain["results"].tobytes() != ref["results"].tobytes()


def _random_replies(game, leaves, seed, density):
    """A reply table made of plausible replies: for a sample of (player, preceding action) keys, an action that was
    legal for that player somewhere in uniform playouts of `leaves` (so that replies are often, not always, legal)."""
    rng = np.random.default_rng(seed)
    na = o.num_actions(game)
    ref = o.playout_batch(game, leaves, 8, seed=seed, want_trace=True, max_trace=MAX_TRACE[game], max_depth=MAX_TRACE[game])
    table = np.zeros((2, na), dtype=np.uint16)
    aid = (lambda a: (a >> 8) * 64 + (a & 0xFF)) if game in (o.BREAKTHROUGH, o.KNIGHTTHROUGH) else (lambda a: a)
    for li in range(len(leaves)):
        for j in range(8):
            d = int(ref["records"][li * 8 + j]["depth"])
            tr = ref["trace"][li * 8 + j]
            for t in range(1, d):
                if rng.random() < density:
                    p = (int(leaves[li]["turn"]) + t) & 1
                    table[p, aid(int(tr[t - 1]))] = int(tr[t]) + 1
    return table


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11])
@pytest.mark.parametrize("eps", [0.1, 0.0])
def test_last_good_reply_bit_exact(game, eps):
    """EpsilonGreedy<Lgr<Uniform>> (simulate.rs:936-1033): traces, records and final states per playout, and the
    batch's reply-table updates (backprop.rs:908-923) as last-winning-occurrence stamps."""
    leaves = _leaves(game, n_per=5)
    # playout()'s prev_action: the move that led to the leaf -- some move of the player who is NOT to move
    prevs = _random_replies(game, leaves, 3, 1.0)
    by_player = [prevs[p][prevs[p] > 0] for p in range(2)]  # prevs[p] holds actions made by player p (+1)
    for i in range(len(leaves)):
        other = by_player[1 - int(leaves[i]["turn"])]
        leaves["prev"][i] = 0 if i % 3 == 0 else other[i % len(other)]
    depth = 200 if game == o.KNIGHTTHROUGH else 0
    for density, k in ((0.0, 33), (0.3, 64), (1.0, 20)):
        table = _random_replies(game, leaves, 5, density)
        mt = MAX_TRACE[game]
        ref = o.playout_batch(game, leaves, k, policy=o.EPS_LGR, eps=eps, seed=77, max_depth=depth, replies=table,
                              want_reply_updates=True, want_trace=True, max_trace=mt, want_final=True, threads=o.host_threads())
        with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyLgr(eps), seed=77, max_playout_depth=depth) as g:
            g.set_replies(table)
            got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, want_trace=True, want_final=True, max_trace=mt)
        assert (got.records["depth"] == ref["records"]["depth"]).all()
        assert (got.records["outcome"] == ref["records"]["outcome"]).all()
        assert (got.trace == ref["trace"]).all()
        assert got.final.tobytes() == ref["final"].tobytes()
        assert got.results.tobytes() == ref["results"].tobytes()
        assert (got.reply_updates["order"] == ref["reply_order"]).all()
        assert (got.reply_updates["action_plus1"] == ref["reply_action"]).all()
        assert (got.last_win == ref["last_win"]).all()
        if density == 1.0 and eps == 0.0:
            uni = o.playout_batch(game, leaves, k, policy=o.UNIFORM, seed=77, max_depth=depth)
            assert got.results.tobytes() != uni["results"].tobytes()  # replies were actually played


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11])
@pytest.mark.parametrize("eps", [0.1, 0.0])
def test_nst_bit_exact(game, eps):
    """EpsilonGreedy<Nst> (simulate.rs:852-930): per-playout traces, records and final states, the unigram (MAST) delta
    and the bigram delta (backprop.rs:885-899), for bigram tables at several backoff thresholds."""
    leaves = _leaves(game, n_per=5)
    prevs = _random_replies(game, leaves, 3, 1.0)
    by_player = [prevs[p][prevs[p] > 0] for p in range(2)]
    for i in range(len(leaves)):
        other = by_player[1 - int(leaves[i]["turn"])]
        leaves["prev"][i] = 0 if i % 3 == 0 else other[i % len(other)]
    depth = 200 if game == o.KNIGHTTHROUGH else 0
    mt = MAX_TRACE[game]
    # statistics of uniform playouts as the tables: realistic, with many pairs above and below the threshold
    base = o.playout_batch(game, leaves, 96, seed=13, max_depth=depth, want_mast_delta=True, want_bigram_delta=True, threads=o.host_threads())
    mast, bigram = base["mast_delta"], base["bigram_delta"]
    for thr, table, k in ((5, bigram, 48), (1, bigram, 33), (0, bigram, 16), (5, None, 16)):
        ref = o.playout_batch(game, leaves, k, policy=o.EPS_NST, eps=eps, seed=78, max_depth=depth, mast=mast, bigram=table,
                              nst_threshold=thr, want_mast_delta=True, want_bigram_delta=True, want_trace=True, max_trace=mt,
                              want_final=True, threads=o.host_threads())
        with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyNst(eps, thr), seed=78, max_playout_depth=depth) as g:
            g.set_bigrams(table)
            got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_trace=True, want_final=True, max_trace=mt,
                                  want_mast_delta=True)
        name = (o.GAME_NAMES[game], thr)
        assert (got.records["depth"] == ref["records"]["depth"]).all(), name
        assert (got.records["outcome"] == ref["records"]["outcome"]).all(), name
        assert (got.trace == ref["trace"]).all(), name
        assert got.final.tobytes() == ref["final"].tobytes(), name
        assert got.results.tobytes() == ref["results"].tobytes(), name
        assert got.mast_delta.tobytes() == ref["mast_delta"].tobytes(), name
        assert got.bigram_delta.tobytes() == ref["bigram_delta"].tobytes(), name
        if table is not None and thr == 1 and eps == 0.0:
            plain = o.playout_batch(game, leaves, k, policy=o.EPS_MAST, eps=eps, seed=78, max_depth=depth, mast=mast)
            assert got.results.tobytes() != plain["results"].tobytes()  # the bigram scores changed picks


@pytest.mark.parametrize("game", GAMES)
@pytest.mark.parametrize("policy", [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WIN, o.DECISIVE_ANTI, o.DECISIVE_MAST_WIN])
def test_lane_refill_paths_bit_exact(game, policy, monkeypatch):
    """Bench-size batches run 512-playout grabs: a lane retires a playout and pulls the next one of its task at a window
    boundary.  Small batches never do (32-playout grabs), so this forces the grab size and checks every per-playout and
    aggregate output again, with tasks cut at leaf boundaries (k = 200 is not a multiple of 32)."""
    monkeypatch.setenv("MR_GRAB_MIN", "512")
    monkeypatch.setenv("MR_GRAB_MAX", "512")
    leaves = _leaves(game, n_per=2)
    depth = 200 if game == o.KNIGHTTHROUGH else 0
    mast = _random_mast(game, 3) if policy in (o.EPS_MAST, o.DECISIVE_MAST_WIN) else None
    _compare(game, policy, leaves, 200, seed=91, max_depth=depth, mast=mast, amaf=True, delta=True)


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.OTHELLO])
@pytest.mark.parametrize("policy", [o.EPS_LGR, o.EPS_NST])
def test_context_policies_on_long_tasks(game, policy, monkeypatch):
    """Many playouts per leaf: lanes retire playouts and pull new ones inside one task, and every new playout must start
    from the LEAF's preceding action again (not from the last move of the playout the lane just finished).  Small batches
    are cut into 32-playout grabs (one playout per lane); the launch knob forces the 512-playout grabs of bench-size
    batches."""
    monkeypatch.setenv("MR_GRAB_MIN", "512")
    monkeypatch.setenv("MR_GRAB_MAX", "512")
    leaves = _leaves(game, n_per=1)
    prevs = _random_replies(game, leaves, 3, 1.0)
    by_player = [prevs[p][prevs[p] > 0] for p in range(2)]
    for i in range(len(leaves)):
        other = by_player[1 - int(leaves[i]["turn"])]
        leaves["prev"][i] = other[i % len(other)]
    k, mt = 1500, MAX_TRACE[game]
    base = o.playout_batch(game, leaves, 64, seed=13, want_mast_delta=True, want_bigram_delta=True, threads=o.host_threads())
    kw = dict(policy=policy, eps=0.1, seed=5, want_trace=True, max_trace=mt, threads=o.host_threads())
    if policy == o.EPS_LGR:
        table = _random_replies(game, leaves, 5, 1.0)
        ref = o.playout_batch(game, leaves, k, replies=table, want_reply_updates=True, **kw)
        with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyLgr(0.1), seed=5) as g:
            g.set_replies(table)
            got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, want_trace=True, max_trace=mt)
        assert (got.reply_updates["order"] == ref["reply_order"]).all()
        assert (got.reply_updates["action_plus1"] == ref["reply_action"]).all()
    else:
        ref = o.playout_batch(game, leaves, k, mast=base["mast_delta"], bigram=base["bigram_delta"], nst_threshold=1,
                              want_bigram_delta=True, **kw)
        with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyNst(0.1, 1), seed=5) as g:
            g.set_bigrams(base["bigram_delta"])
            got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=base["mast_delta"], want_trace=True, max_trace=mt)
        assert got.bigram_delta.tobytes() == ref["bigram_delta"].tobytes()
    assert (got.trace == ref["trace"]).all()
    assert got.results.tobytes() == ref["results"].tobytes()


def _random_replies2(game, leaves, seed, density):
    """A 2-ply reply table of plausible entries: (own previous action, opponent's answer) -> the action that followed in
    uniform playouts of `leaves`."""
    rng = np.random.default_rng(seed)
    ns = o.num_slots(game)
    ref = o.playout_batch(game, leaves, 8, seed=seed, want_trace=True, max_trace=MAX_TRACE[game], max_depth=MAX_TRACE[game])
    table = np.zeros((2, ns, ns), dtype=np.uint16)
    for li in range(len(leaves)):
        for j in range(8):
            d = int(ref["records"][li * 8 + j]["depth"])
            tr = ref["trace"][li * 8 + j]
            for t in range(2, d):
                if rng.random() < density:
                    p = (int(leaves[li]["turn"]) + t) & 1
                    table[p, o.action_slot(game, p, int(tr[t - 2])), o.action_slot(game, 1 - p, int(tr[t - 1]))] = int(tr[t]) + 1
    return table


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11])
@pytest.mark.parametrize("big_grabs,policy", [(False, o.EPS_LGR2), (True, o.EPS_LGR2), (False, o.EPS_LGR2_MAST), (True, o.EPS_LGR2_MAST)])
def test_lgrf2_bit_exact(game, big_grabs, policy, monkeypatch):
    """EpsilonGreedy<Lgr2<Lgr<Uniform>>> (simulate.rs:1036-1142): per-playout traces, records and final states; and both
    reply tables after the batch -- the LGR-1 table by last write, the 2-ply table with its forgetting rule
    (backprop.rs:926-966) through the insert / resolve protocol -- against the oracle's LITERAL trial-by-trial updates."""
    if big_grabs:
        monkeypatch.setenv("MR_GRAB_MIN", "512")
        monkeypatch.setenv("MR_GRAB_MAX", "512")
    leaves = _leaves(game, n_per=3)
    prevs = _random_replies(game, leaves, 3, 1.0)
    by_player = [prevs[p][prevs[p] > 0] for p in range(2)]
    for i in range(len(leaves)):
        turn = int(leaves[i]["turn"])
        other, own = by_player[1 - turn], by_player[turn]
        leaves["prev"][i] = 0 if i % 4 == 0 else other[i % len(other)]
        leaves["prev2"][i] = 0 if i % 3 == 0 else own[(5 * i) % len(own)]
    depth = 200 if game == o.KNIGHTTHROUGH else 0
    mt = MAX_TRACE[game]
    with_mast = policy == o.EPS_LGR2_MAST  # Lgr2<Lgr<Mast>>: the greedy MAST pick as the last fallback (strategy.rs:158-171)
    mast = _random_mast(game, 8) if with_mast else None
    for eps, density, k in ((0.1, 0.5, 150 if big_grabs else 40), (0.0, 1.0, 33), (0.1, 0.0, 20)):
        t1 = _random_replies(game, leaves, 5, density)
        t2 = _random_replies2(game, leaves, 6, density)
        ref = o.playout_batch(game, leaves, k, policy=policy, eps=eps, seed=79, max_depth=depth, replies=t1, replies2=t2, mast=mast,
                              want_final_tables=True, want_trace=True, max_trace=mt, want_final=True, want_mast_delta=with_mast,
                              threads=o.host_threads())
        strat = mb.EpsilonGreedyLgr2Mast(eps) if with_mast else mb.EpsilonGreedyLgr2(eps)
        with mb.GpuRollout(mb.Game(game), strat, seed=79, max_playout_depth=depth) as g:
            g.set_replies(t1)
            g.set_replies2(t2)
            got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_trace=True, want_final=True, max_trace=mt,
                                  want_mast_delta=with_mast)
            removed = g.resolve_replies2(got.reply2_inserts, t2)
        if with_mast:
            assert got.mast_delta.tobytes() == ref["mast_delta"].tobytes()
        name = (o.GAME_NAMES[game], eps, density)
        assert (got.records["depth"] == ref["records"]["depth"]).all(), name
        assert (got.records["outcome"] == ref["records"]["outcome"]).all(), name
        assert (got.trace == ref["trace"]).all(), name
        assert got.final.tobytes() == ref["final"].tobytes(), name
        assert got.results.tobytes() == ref["results"].tobytes(), name
        new1 = np.where(got.reply_updates["order"] > 0, got.reply_updates["action_plus1"], t1)
        assert (new1 == ref["replies_final"]).all(), name
        new2 = mb.GpuRollout.apply_replies2(t2, got.reply2_inserts, removed)
        assert (new2 == ref["replies2_final"]).all(), name
        lost = [(ref["records"]["outcome"].reshape(len(leaves), k) == (o.WINNER_P1 if p == 0 else o.WINNER_P0)) for p in range(2)]
        for p in range(2):
            want = np.array([(np.nonzero(row)[0].max() + 1) if row.any() else 0 for row in lost[p]])
            assert (got.last_lose[:, p] == want).all(), name
        if density == 1.0:
            assert removed.any() and (new2 != t2).any()  # the forgetting rule fired and inserts happened
            plain = o.playout_batch(game, leaves, k, policy=o.EPS_LGR, eps=eps, seed=79, max_depth=depth, replies=t1)
            assert plain["results"].tobytes() != ref["results"].tobytes()  # 2-ply replies (and MAST picks) were actually played


def test_reference_policy_kats_on_the_gpu():
    """The reference's own policy KATs (mcts-tests/tests/decisive_move.rs:10-61, ttt_strategies.rs:916-1075), set up on
    Connect4 as in tests/test_oracle_golden.py, answered by the CUDA path directly."""
    def play(cols):
        st = mb.initial_state(mb.Game.CONNECT4)
        for c in cols:
            st = mb.apply(mb.Game.CONNECT4, st, c)
        return st

    def first_moves(strategy, state, n=64, setup=None, **kw):
        with mb.GpuRollout(mb.Game.CONNECT4, strategy, seed=5, max_playout_depth=1) as g:
            if setup:
                setup(g)
            r = g.simulate_many(state, n, want_trace=True, max_trace=1, **kw)
        return set(int(a) for a in r.trace[:, 0]), r

    block = play([6, 0, 6, 1, 5, 2])  # White threatens d1, Black has no win: the block is the only anti-decisive move
    assert first_moves(mb.DecisiveMove("anti_decisive"), block)[0] == {3}
    assert len(first_moves(mb.Uniform(), block, n=256)[0]) == 7
    win = play([0, 6, 1, 6, 2, 6])  # Black wins at d1 while White threatens g4: the win comes first
    assert first_moves(mb.DecisiveMove("anti_decisive"), win)[0] == {3}
    assert first_moves(mb.DecisiveMoveWin(), win)[0] == {3}
    # NST backoff: unigram favours column 1, the bigram after column 2 favours column 0
    s = mb.initial_state(mb.Game.CONNECT4)
    s["prev"] = 3
    mast = np.zeros((2, 7), dtype=mb.STAT_DTYPE)
    mast["visits"][0] = 10
    mast["score"][0] = [0, 10, -10, -10, -10, -10, -10]
    bigram = np.zeros((2, 7, 7), dtype=mb.STAT_DTYPE)
    bigram["visits"][0, 2, 0], bigram["score"][0, 2, 0] = 2, 2
    bigram["visits"][0, 2, 1], bigram["score"][0, 2, 1] = 2, 0
    assert first_moves(mb.EpsilonGreedyNst(0.0, 5), s, setup=lambda g: g.set_bigrams(bigram), stats=mast)[0] == {1}
    bigram["visits"][0, 2, 0], bigram["score"][0, 2, 0] = 5, 5
    bigram["visits"][0, 2, 1] = 5
    assert first_moves(mb.EpsilonGreedyNst(0.0, 5), s, setup=lambda g: g.set_bigrams(bigram), stats=mast)[0] == {0}
    s["prev"] = 0
    assert first_moves(mb.EpsilonGreedyNst(0.0, 1), s, setup=lambda g: g.set_bigrams(bigram), stats=mast)[0] == {1}
    # NST table population: a two-ply playout puts its one pair into the second mover's table only
    with mb.GpuRollout(mb.Game.CONNECT4, mb.EpsilonGreedyNst(0.1), seed=7, max_playout_depth=2) as g:
        r = g.simulate_many(mb.initial_state(mb.Game.CONNECT4), 1, stats=np.zeros((2, 7), dtype=mb.STAT_DTYPE), want_trace=True, max_trace=2)
    a0, a1 = int(r.trace[0][0]), int(r.trace[0][1])
    assert r.bigram_delta["visits"][1, a0, a1] == 1 and r.bigram_delta["visits"][1].sum() == 1 and r.bigram_delta["visits"][0].sum() == 0


def test_reference_shaped_playout_call():
    """SimulateStrategy::playout(state, max_playout_depth, stats, prev_action, rng) -> Trial"""
    game = o.CONNECT4
    s = o.initial_state(game)
    ref = o.playout_batch(game, s, 3, seed=99, leaf_id_base=7, want_trace=True, max_trace=42, want_final=True)
    with mb.GpuRollout(mb.Game.CONNECT4, mb.Uniform()) as g:
        t = g.playout(s.view(mb.LEAF_DTYPE), rng=(99, 7, 2))
    rec = ref["records"][2]
    assert t.depth == rec["depth"] and t.end_type == mb.EndType.NATURAL_END
    assert [a for a, _ in t.actions] == [int(x) for x in ref["trace"][2][: t.depth]]
    assert [p for _, p in t.actions] == [i & 1 for i in range(t.depth)]
    assert t.state.tobytes() == ref["final"][2:3].tobytes()
    assert int(t.terminal) == rec["outcome"]


def test_double_buffered_slots_and_errors():
    game = o.HEX11
    a, b = _leaves(game, n_per=2), _leaves(game, n_per=3, seed=5)
    with mb.GpuRollout(mb.Game.HEX11, seed=3) as g:
        g.submit(a.view(mb.LEAF_DTYPE), 32, slot=0)
        g.submit(b.view(mb.LEAF_DTYPE), 32, slot=1, leaf_id_base=500)
        with pytest.raises(mb.RolloutError) as e:
            g.submit(a.view(mb.LEAF_DTYPE), 32, slot=0)
        assert "MR_ERR_BUSY" in str(e.value)
        r1, r0 = g.wait(1), g.wait(0)
        assert r0.results.tobytes() == o.playout_batch(game, a, 32, seed=3)["results"].tobytes()
        assert r1.results.tobytes() == o.playout_batch(game, b, 32, seed=3, leaf_id_base=500)["results"].tobytes()
        assert g.kernel_launches() == 2
    with mb.GpuRollout(mb.Game.HEX11, mb.EpsilonGreedyMast()) as g:
        with pytest.raises(mb.RolloutError) as e:  # a MAST policy without a MAST snapshot
            g.simulate_many(a.view(mb.LEAF_DTYPE), 4)
        assert "MR_ERR_INVALID" in str(e.value)
    with mb.GpuRollout(mb.Game.KNIGHTTHROUGH, mb.Uniform()) as g:
        with pytest.raises(mb.RolloutError) as e:  # unbounded playouts cannot carry a MAST delta
            g.simulate_many(mb.initial_state(mb.Game.KNIGHTTHROUGH), 4, want_mast_delta=True)
        assert "MR_ERR_INVALID" in str(e.value)


def test_table_policy_error_paths():
    """The table-policy entry points fail loudly on misuse: a table of another game, reading results before mr_wait or
    from a batch of another policy, a merged LGRF-2 candidate that names no action of the player."""
    import ctypes as C

    from mcts_b200._lib import REPLY_UPDATE_DTYPE, ptr

    leaves = _leaves(o.CONNECT4, n_per=2).view(mb.LEAF_DTYPE)
    mast = _random_mast(o.CONNECT4, 1)
    with mb.GpuRollout(mb.Game.CONNECT4, mb.EpsilonGreedyNst(0.1), seed=1) as g:
        ns_bt = o.num_slots(o.BREAKTHROUGH)
        other = np.zeros((2, ns_bt, ns_bt), dtype=mb.STAT_DTYPE)
        assert g.lib.mr_set_bigrams(g._h, int(mb.Game.BREAKTHROUGH), ptr(other), 5) == 0
        with pytest.raises(mb.RolloutError) as e:  # the bigram table belongs to another game
            g.simulate_many(leaves, 8, stats=mast)
        assert "MR_ERR_INVALID" in str(e.value)
        g.set_bigrams(None)
        with pytest.raises(mb.RolloutError):  # an NST batch needs the unigram (MAST) table
            g.simulate_many(leaves, 8)
        out = np.zeros((2, 7, 7), dtype=mb.STAT_DTYPE)
        assert g.lib.mr_bigram_delta(g._h, 0, ptr(out)) != 0  # no completed NST batch on the slot
        g.submit(leaves, 8, stats=mast)
        assert g.lib.mr_bigram_delta(g._h, 0, ptr(out)) != 0  # still in flight
        r = g.wait()
        assert r.bigram_delta is not None and int(r.bigram_delta["visits"].sum()) > 0
    with mb.GpuRollout(mb.Game.CONNECT4, mb.EpsilonGreedyLgr(0.1), seed=1) as g:
        g.simulate_many(leaves, 8)
        ins = np.zeros((2, 7, 7), dtype=REPLY_UPDATE_DTYPE)
        assert g.lib.mr_reply2_inserts(g._h, 0, ptr(ins), None) != 0  # an LGR-1 batch holds no 2-ply inserts
    with mb.GpuRollout(mb.Game.CONNECT4, mb.EpsilonGreedyLgr2(0.1), seed=1) as g:
        r = g.simulate_many(leaves, 8)
        bad = r.reply2_inserts.copy()
        bad["order"][0, 0, 0] = 1 << 13
        bad["action_plus1"][0, 0, 0] = 9  # column 8 does not exist
        with pytest.raises(mb.RolloutError) as e:
            g.resolve_replies2(bad)
        assert "names no action" in str(e.value)
        assert g.resolve_replies2(r.reply2_inserts).shape == (2, 7, 7)  # and the slot can still be resolved afterwards
        assert mb.num_actions(mb.Game.CONNECT4) == 7 and C.sizeof(C.c_void_p) == 8


@pytest.mark.parametrize("game", GAMES)
def test_full_size_properties(game):
    """At bench size the oracle is too slow to replay everything; check size-independent properties and
    replay a random sample of leaves bit-exactly."""
    n, k = 1024, 1024
    with mb.GpuRollout(mb.Game(game), seed=2024) as g:
        leaves = g.random_positions(MID_PLIES[game][1], n, seed=1)
        assert leaves.tobytes() == o.random_positions(game, 1, MID_PLIES[game][1], n).tobytes()
        r = g.simulate_many(leaves, k)
        r2 = g.simulate_many(leaves, k)
    res = r.results
    assert res.tobytes() == r2.results.tobytes()  # idempotent: no hidden state between batches
    assert ((res["wins"][:, 0].astype(np.int64) + res["wins"][:, 1] + res["draws"]) == k).all()
    assert res["cutoffs"].sum() == 0
    if game == o.HEX11:
        assert res["draws"].sum() == 0  # Hex has no draws
    if game in (o.BREAKTHROUGH, o.KNIGHTTHROUGH):
        assert res["draws"].sum() <= n * k // 1000  # only "no actions left" endings (simulate.rs:112-116), rare
    sample = np.random.default_rng(0).choice(n, size=6, replace=False)
    for i in sample:
        ref = o.playout_batch(game, leaves[i : i + 1].view(o.STATE_DTYPE), k, seed=2024, leaf_id_base=int(i), threads=o.host_threads())
        assert ref["results"][0].tobytes() == res[i].tobytes(), (o.GAME_NAMES[game], i)


@pytest.mark.parametrize("game", [o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.CONNECT4, o.OTHELLO, o.HEX11])
@pytest.mark.parametrize("eps", [0.1, 0.0])
def test_production_variants_match_the_oracle(game, eps):
    """The kernel variants a search actually launches (MAST delta only, MAST delta + AMAF: no trace, no final state)
    against the oracle's tables: every other parity test runs the traced variant, which is a different instantiation.
    Tables and results must come out bit for bit the same, for table densities from "all ties" to "fully visited", a warm
    table (scores of exactly 1 on the winning moves) and enough playouts per lane for the per-lane logs to be drained
    more than once."""
    leaves = _leaves(game, n_per=3)
    warm = o.playout_batch(game, leaves, 64, seed=5, max_depth=200, want_mast_delta=True, threads=o.host_threads())["mast_delta"]
    tables = [_random_mast(game, 7, 0.0), _random_mast(game, 8, 0.5), _random_mast(game, 9, 1.0), warm]
    for t, mast in enumerate(tables):
        k = 2400 if t == 3 else 160
        for amaf in (False, True):
            ref = o.playout_batch(game, leaves, k, policy=o.EPS_MAST, eps=eps, max_depth=200, seed=77 + t, mast=mast,
                                  want_amaf=amaf, want_mast_delta=True, threads=o.host_threads())
            with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyMast(eps), seed=77 + t, max_playout_depth=200) as g:
                got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_amaf=amaf, want_mast_delta=True)
            name = (o.GAME_NAMES[game], t, amaf)
            assert got.results.tobytes() == ref["results"].tobytes(), name
            assert got.mast_delta.tobytes() == ref["mast_delta"].tobytes(), name
            if amaf:
                assert got.amaf.tobytes() == ref["amaf"].tobytes(), name


def test_headline_workload_at_full_size():
    """BASELINE.json configs[1] at its bench size (4096 leaves x 4096 playouts, epsilon-MAST, depth 200): conservation
    properties over the whole batch, the MAST delta's checksum, and bit-exact replay of sampled leaves by the oracle."""
    game, n, k, depth = o.BREAKTHROUGH, 4096, 4096, 200
    leaves = o.random_positions(game, 2, 20, n)
    mast = o.playout_batch(game, leaves, 16, seed=999, max_depth=depth, want_mast_delta=True, threads=o.host_threads())["mast_delta"]
    with mb.GpuRollout(mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), seed=31337, max_playout_depth=depth) as g:
        r = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_mast_delta=True)
    res = r.results
    assert ((res["wins"][:, 0].astype(np.int64) + res["wins"][:, 1] + res["draws"]) == k).all()
    # every ply is one (action, player) occurrence; every decided playout scores +1 for the winner's plies, -1 for the loser's
    assert int(r.mast_delta["visits"].sum()) == int(res["sum_depth"].sum())
    assert abs(int(r.mast_delta["score"].sum())) <= int(res["sum_depth"].sum())
    sample = np.random.default_rng(1).choice(n, size=4, replace=False)
    for i in sample:
        ref = o.playout_batch(game, leaves[i : i + 1], k, policy=o.EPS_MAST, eps=0.1, max_depth=depth, seed=31337, leaf_id_base=int(i),
                              mast=mast, threads=o.host_threads())
        assert ref["results"][0].tobytes() == res[i].tobytes(), i


@pytest.mark.parametrize("policy", [o.EPS_LGR, o.EPS_NST, o.EPS_LGR2])
def test_table_policies_at_large_size(policy):
    """The context policies on a batch large enough for 512-playout grabs, lane refills and table flushes (1024 leaves x
    1024 playouts): conservation properties over the whole batch and bit-exact replay of sampled leaves by the oracle
    (a leaf's playouts depend only on the frozen tables, so the oracle can replay one leaf with its leaf id)."""
    game, n, k, depth = o.BREAKTHROUGH, 1024, 1024, 200
    leaves = o.random_positions(game, 2, 20, n)
    tr = o.playout_batch(game, leaves[:64], 2, seed=4, want_trace=True, max_trace=2, max_depth=depth)
    first = {}  # first move seen for each colour: a plausible preceding action for leaves of the other colour
    for i in range(64):
        first.setdefault(int(leaves[i]["turn"]), int(tr["trace"][2 * i][0]))
    for i in range(n):
        if i % 5 and (1 - int(leaves[i]["turn"])) in first:
            leaves["prev"][i] = first[1 - int(leaves[i]["turn"])] + 1
    base = o.playout_batch(game, leaves[:128], 64, seed=999, max_depth=depth, want_mast_delta=True, want_bigram_delta=True,
                           threads=o.host_threads())
    t1 = _random_replies(game, leaves[:32], 5, 1.0)
    t2 = _random_replies2(game, leaves[:32], 6, 1.0)
    strat = {o.EPS_LGR: mb.EpsilonGreedyLgr(0.1), o.EPS_NST: mb.EpsilonGreedyNst(0.1, 3), o.EPS_LGR2: mb.EpsilonGreedyLgr2(0.1)}[policy]
    with mb.GpuRollout(mb.Game(game), strat, seed=4242, max_playout_depth=depth) as g:
        g.set_replies(t1)
        g.set_replies2(t2)
        g.set_bigrams(base["bigram_delta"])
        r = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=base["mast_delta"] if policy == o.EPS_NST else None,
                            want_mast_delta=policy == o.EPS_NST)
        removed = g.resolve_replies2(r.reply2_inserts, t2) if policy == o.EPS_LGR2 else None
    res = r.results
    assert ((res["wins"][:, 0].astype(np.int64) + res["wins"][:, 1] + res["draws"]) == k).all()
    plies = int(res["sum_depth"].sum())
    if policy == o.EPS_NST:
        assert int(r.mast_delta["visits"].sum()) == plies
        # every ply with a predecessor is one bigram occurrence: all plies but the first ply of playouts whose leaf has none
        # (Breakthrough leaves after 20 plies are not terminal: every playout has a first ply)
        first_plies_without_context = int((leaves["prev"] == 0).sum()) * k
        assert int(r.bigram_delta["visits"].sum()) == plies - first_plies_without_context
    else:
        assert (r.reply_updates["order"] > 0).any() and (r.last_win > 0).all()
    if policy == o.EPS_LGR2:
        assert (r.reply2_inserts["order"] > 0).any() and removed.any() and (r.last_lose > 0).any()
    kw = {o.EPS_LGR: dict(replies=t1), o.EPS_NST: dict(mast=base["mast_delta"], bigram=base["bigram_delta"], nst_threshold=3),
          o.EPS_LGR2: dict(replies=t1, replies2=t2)}[policy]
    for i in np.random.default_rng(2).choice(n, size=3, replace=False):
        ref = o.playout_batch(game, leaves[i : i + 1], k, policy=policy, eps=0.1, max_depth=depth, seed=4242, leaf_id_base=int(i),
                              threads=o.host_threads(), **kw)
        assert ref["results"][0].tobytes() == res[i].tobytes(), i


def test_win_rate_agrees_with_independent_cpu_rollouts():
    """Root-child win rates from GPU playouts vs CPU oracle playouts drawn with a DIFFERENT seed agree
    within a Wilson 99.9% interval (mcts-bench/src/tournament.rs:98-111)."""
    game = o.CONNECT4
    root = o.initial_state(game)
    children = np.concatenate([o.apply(game, root, c) for c in range(7)])
    k = 20000
    with mb.GpuRollout(mb.Game.CONNECT4, seed=1) as g:
        got = g.simulate_many(children.view(mb.LEAF_DTYPE), k).results
    ref = o.playout_batch(game, children, k, seed=2, threads=o.host_threads())["results"]
    z = 3.29
    for c in range(7):
        p1, p2 = got[c]["wins"][0] / k, ref[c]["wins"][0] / k
        se = np.sqrt(p1 * (1 - p1) / k + p2 * (1 - p2) / k)
        assert abs(p1 - p2) < z * se + 1e-9, (c, p1, p2)


def wilson_interval(successes: float, total: int, z: float):
    """mcts-bench/src/tournament.rs:98-111, restated."""
    if total == 0:
        return 0.0, 1.0
    n = float(total)
    p_hat = successes / n
    z2 = z * z
    denom = 1.0 + z2 / n
    center = p_hat + z2 / (2.0 * n)
    margin = z * np.sqrt(p_hat * (1.0 - p_hat) / n + z2 / (4.0 * n * n))
    return max(0.0, (center - margin) / denom), min(1.0, (center + margin) / denom)


@pytest.mark.parametrize("game,policy,k,max_depth", [
    (o.CONNECT4, o.UNIFORM, 20000, 0),
    (o.BREAKTHROUGH, o.EPS_MAST, 4000, 200),   # the headline policy: epsilon-greedy over a MAST table
    (o.OTHELLO, o.DECISIVE_WIN, 4000, 0),      # BASELINE config 4's rollout policy
])
def test_root_child_win_rates_inside_wilson_intervals_of_cpu_rollouts(game, policy, k, max_depth):
    """north_star's statistical parity.  For every root child, the win rate of 16 k GPU playouts (own seeds; sampling
    error a quarter of the CPU sample's) must lie inside the Wilson 99.99 % interval (z = 3.89; `wilson_interval`,
    mcts-bench/src/tournament.rs:98-111) of k CPU oracle rollouts drawn with DIFFERENT seeds -- for two independent seed
    pairs.  A correct engine trips one of the ~60 intervals with probability < 1 %; the seeds are fixed, so the
    outcome is deterministic."""
    root = o.initial_state(game)
    children = np.concatenate([o.apply(game, root, a) for a in o.generate_actions(game, root)])
    mover = int(root["turn"][0])
    strat = {o.UNIFORM: mb.Uniform(), o.EPS_MAST: mb.EpsilonGreedyMast(0.1), o.DECISIVE_WIN: mb.DecisiveMoveWin()}[policy]
    mast = None
    if policy == o.EPS_MAST:  # one table for both sides: the delta of a uniform batch
        mast = o.playout_batch(game, children, 64, seed=77, max_depth=max_depth, want_mast_delta=True, threads=o.host_threads())["mast_delta"]
    z = 3.89
    for gpu_seed, cpu_seed in [(1, 2), (1001, 2002)]:
        with mb.GpuRollout(mb.Game(game), strat, seed=gpu_seed, max_playout_depth=max_depth) as g:
            got = g.simulate_many(children.view(mb.LEAF_DTYPE), 16 * k, stats=mast).results
        ref = o.playout_batch(game, children, k, policy=policy, eps=0.1, max_depth=max_depth, seed=cpu_seed, mast=mast,
                              threads=o.host_threads())["results"]
        for c in range(len(children)):
            p_gpu = int(got[c]["wins"][mover]) / (16 * k)
            lo, hi = wilson_interval(int(ref[c]["wins"][mover]), k, z)
            assert lo <= p_gpu <= hi, (c, "gpu win rate outside the cpu rollouts' Wilson interval", p_gpu, lo, hi)


def test_randomized_differential_sweep():
    """Seeded random batch shapes: game x policy x flags x (n, k) x leaf id base x depth cutoff, GPU vs oracle,
    every playout's trace, record, final state and every aggregate compared bit for bit."""
    import os

    rng = np.random.default_rng(int(os.environ.get("MR_FUZZ_SEED", "20260101")))  # a longer sweep: MR_FUZZ_CASES=400 MR_FUZZ_SEED=7
    policies = {
        o.CONNECT4: [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WIN, o.DECISIVE_ANTI],
        o.BREAKTHROUGH: [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WIN, o.DECISIVE_ANTI, o.DECISIVE_MAST_WINLOSS],
        o.KNIGHTTHROUGH: [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WINLOSS, o.DECISIVE_ANTI],
        o.OTHELLO: [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WINLOSSDRAW, o.DECISIVE_ANTI],
        o.HEX11: [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WIN, o.DECISIVE_ANTI, o.DECISIVE_MAST_WIN],
    }
    max_plies = {o.CONNECT4: 30, o.BREAKTHROUGH: 50, o.KNIGHTTHROUGH: 30, o.OTHELLO: 56, o.HEX11: 100}
    for case in range(int(os.environ.get("MR_FUZZ_CASES", "40"))):
        game = GAMES[case % 5]
        policy = policies[game][int(rng.integers(len(policies[game])))]
        n, k = int(rng.integers(1, 24)), int(rng.choice([1, 2, 7, 31, 32, 33, 64, 97]))
        plies = int(rng.integers(0, max_plies[game]))
        leaves = o.random_positions(game, 1000 + case, plies, n)
        cutoff = int(rng.choice([0, 0, 5, 40]))
        needs_bound = game == o.KNIGHTTHROUGH
        amaf = bool(rng.integers(2))
        delta = policy in (o.EPS_MAST, o.DECISIVE_MAST_WINLOSS) or (game in (o.BREAKTHROUGH, o.KNIGHTTHROUGH) and bool(rng.integers(2)))
        if (amaf or delta) and needs_bound and cutoff == 0:
            cutoff = 250
        if game == o.HEX11 and policy == o.UNIFORM:
            delta = bool(rng.integers(2))  # uniform + MAST delta routes Hex to the per-ply kernel
        uses_mast = policy in (o.EPS_MAST, o.DECISIVE_MAST_WIN, o.DECISIVE_MAST_WINLOSS)
        mast = _random_mast(game, case, float(rng.random())) if uses_mast else None
        _compare(game, policy, leaves, k, seed=int(rng.integers(1 << 62)), max_depth=cutoff, mast=mast,
                 eps=float(rng.choice([0.0, 0.1, 0.5])), leaf_id_base=int(rng.integers(1 << 20)), amaf=amaf, delta=delta)


def test_randomized_sweep_of_the_production_variants():
    """The same kind of seeded sweep over the kernel variants a search launches -- no trace, no final state: results only,
    + AMAF, + MAST delta, + both -- with the aggregates compared bit for bit against the oracle (the traced sweep above
    runs a different instantiation of every kernel).  MR_FUZZ_CASES / MR_FUZZ_SEED lengthen it."""
    import os

    rng = np.random.default_rng(int(os.environ.get("MR_FUZZ_SEED", "20260202")))
    policies = [o.UNIFORM, o.EPS_MAST, o.DECISIVE_WIN, o.DECISIVE_MAST_WIN]
    max_plies = {o.CONNECT4: 30, o.BREAKTHROUGH: 50, o.KNIGHTTHROUGH: 30, o.OTHELLO: 56, o.HEX11: 100}
    for case in range(int(os.environ.get("MR_FUZZ_CASES", "40"))):
        game = GAMES[case % 5]
        policy = policies[int(rng.integers(len(policies)))]
        n, k = int(rng.integers(1, 40)), int(rng.choice([1, 3, 32, 33, 100, 700, 2100]))
        leaves = o.random_positions(game, 5000 + case, int(rng.integers(0, max_plies[game])), n)
        amaf, delta = bool(rng.integers(2)), bool(rng.integers(2))
        depth = int(rng.choice([0, 6, 60])) if not (amaf or delta) and game != o.KNIGHTTHROUGH else int(rng.choice([7, 60, 250]))
        uses_mast = policy in (o.EPS_MAST, o.DECISIVE_MAST_WIN)
        mast = _random_mast(game, case, float(rng.choice([0.0, 0.3, 0.9, 1.0]))) if uses_mast else None
        eps, seed, base = float(rng.choice([0.0, 0.1, 0.5])), int(rng.integers(1 << 62)), int(rng.integers(1 << 20))
        ref = o.playout_batch(game, leaves, k, policy=policy, eps=eps, max_depth=depth, seed=seed, leaf_id_base=base, mast=mast,
                              want_amaf=amaf, want_mast_delta=delta, threads=o.host_threads())
        with mb.GpuRollout(mb.Game(game), _strategy(policy, eps), seed=seed, max_playout_depth=depth) as g:
            got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_amaf=amaf, want_mast_delta=delta, leaf_id_base=base)
        name = (case, o.GAME_NAMES[game], policy, n, k, depth, amaf, delta)
        assert got.results.tobytes() == ref["results"].tobytes(), name
        if amaf:
            assert got.amaf.tobytes() == ref["amaf"].tobytes(), name
        if delta:
            assert got.mast_delta.tobytes() == ref["mast_delta"].tobytes(), name
