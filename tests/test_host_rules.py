"""The tree driver's host-side rules (mr_generate_actions / mr_apply / mr_terminal_status) against the oracle:
same action lists in the same order, same successor states, same terminal status, along random games."""
import numpy as np
import pytest

import mcts_b200 as mb
import oracle_lib as o


@pytest.mark.parametrize("game", range(5))
def test_host_rules_match_oracle_along_random_games(game):
    rng = np.random.default_rng(7 + game)
    n_games = 40 if game != o.HEX11 else 12
    for g in range(n_games):
        s = o.initial_state(game)
        for ply in range(400):
            st_o = o.terminal_status(game, s)
            assert int(mb.terminal_status(mb.Game(game), s.view(mb.LEAF_DTYPE))) == st_o
            if st_o != o.NOT_TERMINAL:
                break
            acts = o.generate_actions(game, s)
            assert mb.generate_actions(mb.Game(game), s.view(mb.LEAF_DTYPE)) == acts
            if not acts:
                break
            a = acts[int(rng.integers(len(acts)))]
            nxt = o.apply(game, s, a)
            assert mb.apply(mb.Game(game), s.view(mb.LEAF_DTYPE), a).tobytes() == nxt.tobytes()
            s = nxt
