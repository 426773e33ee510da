"""The reference's stress sweeps at FULL size (mcts-tests/tests/stress.rs), behind an env switch because they take
minutes: ORACLE_STRESS_FULL=1 python -m pytest tests/test_oracle_stress_full.py.  The default suite runs them at 4 % size
(tests/test_oracle_golden.py); the full-size run of this round is recorded in profiles/oracle_stress_full_r02.txt."""
import ctypes as C
import os

import pytest

import oracle_lib as o

pytestmark = pytest.mark.skipif(os.environ.get("ORACLE_STRESS_FULL") != "1", reason="set ORACLE_STRESS_FULL=1 (minutes of CPU)")


def test_othello_symmetry_stress_sweep_100k_games():
    """stress.rs:249-454: 100 000 games, xorshift64* seed 0xdead_beef_cafe_babe, bitboard movegen / flips / apply against the
    restated naive oracle under all eight symmetries."""
    games, plies = C.c_uint64(), C.c_uint64()
    rc = o.lib().orc_stress_othello(100_000, C.byref(games), C.byref(plies))
    assert rc == 0, f"mismatch code {rc} in game {games.value}"
    assert games.value == 100_000 and plies.value > 100_000 * 55
    print(f"othello stress: {games.value} games, {plies.value} plies, no mismatch")


@pytest.mark.parametrize("game", [o.BREAKTHROUGH, o.KNIGHTTHROUGH])
def test_through_array_board_sweep_full(game):
    """stress.rs:485-818: 5 000 games against the array-board oracle, seed 0xcafe_babe_dead_beef (the default suite already
    runs this one at full size; repeated here so that one command records every sweep)."""
    wins, stuck = C.c_uint64(), C.c_uint64()
    rc = o.lib().orc_stress_through(game, 5000, C.byref(wins), C.byref(stuck))
    assert rc == 0
    print(f"{o.GAME_NAMES[game]} stress: 5000 games, {wins.value} goal wins, {stuck.value} stuck")
