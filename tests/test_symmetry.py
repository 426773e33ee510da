"""Canonical-symmetry transposition keys (MR_SEARCH_SYMMETRY): the reference's own symmetry tests for Othello restated
against BOTH the oracle (oracle/search.c: square-by-square index maps) and the product's host rules
(csrc/host_games.h: word-parallel flips), plus the search-level consequences.

Reference sources:
  games/game-core/src/symmetry.rs:75-84   doc example: index_symmetries(27) = [27, 28, ...] on the 8x8 board
  games/game-core/src/symmetry.rs:113-147 index_symmetries / invert_symmetry
  games/othello/src/lib.rs:1215-1237      test_action_transform_round_trip
  games/othello/src/lib.rs:1243-1286      test_canonical_representation_invariant_under_symmetry
  games/othello/src/lib.rs:1294-1325      test_canonical_representation_respects_symmetry_ply_limit
  games/othello/src/lib.rs:1337-1352      test_othello_sym_search (the table is hit on a small search from the opening)
"""
import numpy as np
import pytest

import oracle_lib as o


def _sym_index(i, k):
    return o.lib().orc_symmetry_index(i, k)


def _sym_invert(i, k):
    return o.lib().orc_symmetry_invert(i, k)


def _image(bits: int, k: int) -> int:
    out = 0
    for i in range(64):
        if (bits >> i) & 1:
            out |= 1 << _sym_index(i, k)
    return out


def _state(black, white, turn=0, flag=0):
    return o.make_state([black, white], turn=turn, flag=flag)


def _product_canon(st):
    import mcts_b200 as mb

    out, sym = mb.canonical_representation(mb.Game.OTHELLO, np.ascontiguousarray(st).view(mb.LEAF_DTYPE))
    return out.view(o.STATE_DTYPE), sym


def test_index_symmetries_reference_example_and_round_trip():
    # symmetry.rs:75-84: cell 27 of the 8x8 board: identity 27, column flip 28
    assert _sym_index(27, 0) == 27 and _sym_index(27, 1) == 28
    # flip_rows: row 3 -> row 4; transpose of (row 3, col 3) is itself; (row 1, col 2) = 10 -> (row 2, col 1) = 17
    assert _sym_index(27, 2) == 35 and _sym_index(27, 3) == 27 and _sym_index(10, 3) == 17
    # othello/lib.rs:1215-1237: invert_action(apply_to_action(a, sym), sym) == a for every square and symmetry
    for k in range(8):
        assert sorted(_sym_index(i, k) for i in range(64)) == list(range(64))
        for i in range(64):
            assert _sym_invert(_sym_index(i, k), k) == i


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_canonical_representation_invariant_under_symmetry(impl):
    """othello/lib.rs:1243-1286: every symmetric image of a reachable state has the same canonical form, turn and pass
    flag untouched; the returned symmetry maps the state onto it."""
    canon_of = (lambda st: o.canonical_representation(o.OTHELLO, st)) if impl == "oracle" else _product_canon
    reachable = [o.initial_state(o.OTHELLO)]
    for m in (19, 18, 20):
        reachable.append(o.apply(o.OTHELLO, reachable[-1], m))
    for st in reachable:
        black, white = int(st["board"][0][0]), int(st["board"][0][1])
        canon, sym = canon_of(st)
        assert (int(canon["board"][0][0]), int(canon["board"][0][1])) == (_image(black, sym), _image(white, sym))
        assert canon["turn"][0] == st["turn"][0] and canon["flag"][0] == st["flag"][0]
        images = [(_image(black, k), _image(white, k)) for k in range(8)]
        assert (int(canon["board"][0][0]), int(canon["board"][0][1])) == min(images)
        assert sym == images.index(min(images))  # the FIRST minimal image (min_by_key)
        for b, w in images:
            c2, _ = canon_of(_state(b, w, turn=int(st["turn"][0]), flag=int(st["flag"][0])))
            assert c2.tobytes() == canon.tobytes()


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_canonical_representation_respects_symmetry_ply_limit(impl):
    """othello/lib.rs:1294-1325: a black-only board with SYMMETRY_PLY_LIMIT + 1 = 21 discs on squares 63, 62, ... is not its
    own minimal image, yet canonical_representation leaves it alone (identity); with 20 discs it is still rotated."""
    canon_of = (lambda st: o.canonical_representation(o.OTHELLO, st)) if impl == "oracle" else _product_canon
    over = sum(1 << (63 - i) for i in range(21))
    assert min(range(8), key=lambda k: (_image(over, k), 0)) != 0  # test setup, as in the reference
    canon, sym = canon_of(_state(over, 0))
    assert sym == 0 and int(canon["board"][0][0]) == over and int(canon["board"][0][1]) == 0
    at_limit = sum(1 << (63 - i) for i in range(20))
    canon, sym = canon_of(_state(at_limit, 0))
    assert sym != 0 and int(canon["board"][0][0]) == _image(at_limit, sym)


def test_oracle_and_product_agree_on_random_positions():
    """Two independent implementations (square-by-square maps / word-parallel flips) on positions from 0 to 40 plies."""
    for plies in (0, 1, 2, 3, 5, 8, 12, 16, 17, 20, 40):
        for st in o.random_positions(o.OTHELLO, 100 + plies, plies, 24):
            st = np.array([st], dtype=o.STATE_DTYPE)
            c_o, s_o = o.canonical_representation(o.OTHELLO, st)
            c_p, s_p = _product_canon(st)
            assert s_o == s_p and c_o.tobytes() == c_p.tobytes(), plies
            discs = bin(int(st["board"][0][0]) | int(st["board"][0][1])).count("1")
            if discs > 20:
                assert s_o == 0 and c_o.tobytes() == st.tobytes()


def test_other_games_have_the_identity():
    import mcts_b200 as mb

    for game in (o.CONNECT4, o.BREAKTHROUGH, o.HEX11):
        st = o.random_positions(game, 3, 6, 1)
        c, s = o.canonical_representation(game, st)
        assert s == 0 and c.tobytes() == st.tobytes()
        c, s = mb.canonical_representation(mb.Game(game), st.view(mb.LEAF_DTYPE))
        assert s == 0 and c.tobytes() == st.tobytes()


def test_symmetric_search_shares_the_four_openings_oracle():
    """othello/lib.rs:1337-1352 test_othello_sym_search restated on the oracle: from the opening position the four legal
    moves lead to ONE position up to symmetry, so with canonical keys they share a node -- every root child then owns the
    same subtree, and far fewer nodes are created than with exact-state keys.  The search stays a legal, deterministic
    search: same seed, same answer; all playouts accounted for."""
    root = o.initial_state(o.OTHELLO)
    kw = dict(iterations=300, batch_leaves=1, k=1, q_init=0, expand_threshold=0, seed=11)
    best_a, ch_a, _ = o.search(o.OTHELLO, root, use_transpositions=True, canonical_symmetry=True, **kw)
    best_b, ch_b, _ = o.search(o.OTHELLO, root, use_transpositions=True, canonical_symmetry=True, **kw)
    assert best_a == best_b and ch_a.tobytes() == ch_b.tobytes()
    assert best_a in (19, 26, 37, 44)
    assert sorted(ch_a["action"].tolist()) == [19, 26, 37, 44]  # the root's own list stays literal
    assert int(ch_a["visits"].sum()) == 300


@pytest.mark.gpu
@pytest.mark.parametrize("batch,k,depth", [(1, 1, 1), (16, 4, 1), (32, 8, 3)])
def test_symmetric_search_bit_exact_against_the_oracle(batch, k, depth):
    """MR_SEARCH_SYMMETRY through mr_search against the oracle's restatement: identical root statistics, sequentially and
    in pipelined batches, from the opening (everything canonicalised) and from a 14-ply position whose subtree crosses
    the 20-disc limit (cached children verified against the real successor)."""
    import mcts_b200 as mb

    roots = [o.initial_state(o.OTHELLO), o.random_positions(o.OTHELLO, 9, 14, 1)]
    for root in roots:
        iters = 1500
        ref_best, ref_ch, ref_plies = o.search(o.OTHELLO, root, iterations=iters, batch_leaves=batch, k=k, policy=o.DECISIVE_WIN,
                                                use_transpositions=True, canonical_symmetry=True, seed=21, pipeline_depth=depth,
                                                threads=o.host_threads())
        with mb.GpuSearch(mb.Game.OTHELLO, mb.DecisiveMoveWin(), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k,
                          use_transpositions=True, canonical_symmetry=True, seed=21, pipeline_depth=depth) as s:
            got = s.choose_action(root.view(mb.LEAF_DTYPE))
        assert got.best_action == ref_best
        assert got.children["action"].tolist() == ref_ch["action"].tolist()
        assert got.children["visits"].tolist() == ref_ch["visits"].tolist()
        assert got.children["score"].tolist() == ref_ch["score"].tolist()
        assert got.playout_plies == ref_plies


@pytest.mark.gpu
def test_symmetric_keys_merge_more_nodes():
    """The point of the canonical keys (othello/lib.rs:451-480): from the opening, fewer nodes for the same number of
    descents than with exact-state keys, and fewer still than without a table."""
    import mcts_b200 as mb

    root = o.initial_state(o.OTHELLO).view(mb.LEAF_DTYPE)
    nodes = {}
    for name, kw in (("none", {}), ("exact", dict(use_transpositions=True)), ("canonical", dict(use_transpositions=True, canonical_symmetry=True))):
        with mb.GpuSearch(mb.Game.OTHELLO, mb.Uniform(), max_iterations=4000, batch_leaves=32, num_rollouts_per_leaf=4, seed=3, **kw) as s:
            r = s.choose_action(root)
        assert int(r.children["visits"].sum()) == (4000 - 1) * 4
        nodes[name] = r.nodes
    assert nodes["canonical"] < nodes["exact"] <= nodes["none"], nodes


@pytest.mark.gpu
def test_symmetry_flag_is_refused_with_action_keyed_tables():
    import mcts_b200 as mb

    root = o.initial_state(o.OTHELLO).view(mb.LEAF_DTYPE)
    with mb.GpuSearch(mb.Game.OTHELLO, mb.EpsilonGreedyMast(0.1), max_iterations=64, batch_leaves=8, use_transpositions=True,
                      canonical_symmetry=True, max_playout_depth=64) as s:
        with pytest.raises(mb.RolloutError):
            s.choose_action(root)
