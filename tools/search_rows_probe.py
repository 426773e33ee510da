#!/usr/bin/env python
"""The search rows of bench.py (configs 1, 3, 4) on one GPU, for A/B runs of host-side changes."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb
from mcts_b200.search import SELECT_GRAVE, QINIT_INFINITY

def run(name, game, strat, total, b, kk, depth, **kw):
    iters = max(1, total // kk)
    with mb.GpuSearch(game, strat, max_iterations=iters, batch_leaves=b, num_rollouts_per_leaf=kk, pipeline_depth=depth, seed=7, **kw) as s:
        root = mb.initial_state(game)
        s.choose_action(root)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); r = s.choose_action(root); best = min(best, time.perf_counter() - t0)
    print(f"{name:>28} {b}x{kk}: {best*1e3:.3f} ms/move (best of 5), select {r.seconds_select*1e3:.2f} submit {r.seconds_submit*1e3:.2f} wait {r.seconds_gpu*1e3:.2f} backprop {r.seconds_backprop*1e3:.2f}")

run("connect4_uct_100k", mb.Game.CONNECT4, mb.Uniform(), 100_000, 64, 64, 4)
run("connect4_uct_100k", mb.Game.CONNECT4, mb.Uniform(), 100_000, 64, 256, 4)
run("hex11_grave", mb.Game.HEX11, mb.Uniform(), 4096 * 256, 128, 256, 4, select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=700)
run("othello_tt", mb.Game.OTHELLO, mb.DecisiveMoveWin(), 4096 * 256, 128, 256, 4, use_transpositions=True)
run("othello_tt_canonical", mb.Game.OTHELLO, mb.DecisiveMoveWin(), 4096 * 256, 128, 256, 4, use_transpositions=True, canonical_symmetry=True)
run("breakthrough_mast_1thr", mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), 1 << 22, 128, 1024, 4, max_playout_depth=200)
