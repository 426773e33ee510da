"""The frozen playout vectors of tests/golden/playouts_v1.npz (written by tests/golden/make_golden.py): the oracle must
still reproduce them on the CPU, and the CUDA path must reproduce them on the GPU -- per playout (trace, record, final
state) and per aggregate (results, AMAF, MAST / bigram deltas, reply-table updates)."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as o

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
_spec = importlib.util.spec_from_file_location("make_golden", GOLDEN_DIR / "make_golden.py")
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN_DIR / "playouts_v1.npz")


def _contexts():
    return list(mg.cases())


def test_fixture_covers_every_game_and_policy(golden):
    keys = {k.rsplit(".", 1)[0] for k in golden.files}
    assert keys == {f"{g}.{p}" for g in o.GAME_NAMES for p in mg.POLICIES}


def test_oracle_reproduces_golden_playouts(golden):
    for game, name, policy, leaves, depth, base, first in _contexts():
        key = f"{o.GAME_NAMES[game]}.{name}"
        assert leaves.tobytes() == golden[f"{key}.leaves"].tobytes(), key
        r = mg.run(game, policy, leaves, depth, base, first)
        for field in ("trace", "records", "results", "final", "amaf", "mast_delta", "bigram_delta", "reply_order", "reply_action",
                      "replies2_final"):
            if f"{key}.{field}" in golden.files:
                assert r[field].tobytes() == golden[f"{key}.{field}"].tobytes(), (key, field)


@pytest.mark.gpu
def test_gpu_reproduces_golden_playouts(golden):
    mb = pytest.importorskip("mcts_b200")
    strategies = {"uniform": mb.Uniform(), "eps_mast": mb.EpsilonGreedyMast(0.1), "decisive_win": mb.DecisiveMoveWin(),
                  "anti_decisive": mb.DecisiveMove("anti_decisive"), "eps_lgr": mb.EpsilonGreedyLgr(0.1),
                  "eps_nst": mb.EpsilonGreedyNst(0.1, 2), "eps_lgr2": mb.EpsilonGreedyLgr2(0.1)}
    for game, name, policy, leaves, depth, base, first in _contexts():
        key = f"{o.GAME_NAMES[game]}.{name}"
        lv = golden[f"{key}.leaves"].view(mb.LEAF_DTYPE)
        with mb.GpuRollout(mb.Game(game), strategies[name], seed=mg.SEED, max_playout_depth=depth) as g:
            kw = dict(want_trace=True, want_final=True, max_trace=mg.MAX_TRACE[game], want_amaf=True, leaf_id_base=17)
            if policy in (o.EPS_MAST, o.EPS_NST):
                kw.update(stats=base["mast_delta"], want_mast_delta=True)
            if policy == o.EPS_NST:
                g.set_bigrams(base["bigram_delta"])
            if policy in (o.EPS_LGR, o.EPS_LGR2):
                g.set_replies(first["replies_final"])
            if policy == o.EPS_LGR2:
                g.set_replies2(first["replies2_final"])
            got = g.simulate_many(lv, mg.K, **kw)
            if policy == o.EPS_LGR2:
                removed = g.resolve_replies2(got.reply2_inserts, first["replies2_final"])
                table2 = mb.GpuRollout.apply_replies2(first["replies2_final"], got.reply2_inserts, removed)
                assert table2.tobytes() == golden[f"{key}.replies2_final"].tobytes(), key
        assert got.trace.tobytes() == golden[f"{key}.trace"].tobytes(), key
        assert got.records.tobytes() == golden[f"{key}.records"].tobytes(), key
        assert got.results.tobytes() == golden[f"{key}.results"].tobytes(), key
        assert got.final.tobytes() == golden[f"{key}.final"].tobytes(), key
        assert got.amaf.tobytes() == golden[f"{key}.amaf"].tobytes(), key
        if f"{key}.mast_delta" in golden.files:
            assert got.mast_delta.tobytes() == golden[f"{key}.mast_delta"].tobytes(), key
        if f"{key}.bigram_delta" in golden.files:
            assert got.bigram_delta.tobytes() == golden[f"{key}.bigram_delta"].tobytes(), key
        if f"{key}.reply_order" in golden.files:
            assert (got.reply_updates["order"] == golden[f"{key}.reply_order"]).all(), key
            assert (got.reply_updates["action_plus1"] == golden[f"{key}.reply_action"]).all(), key
