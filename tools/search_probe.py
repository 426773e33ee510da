#!/usr/bin/env python
"""Search throughput of the GPU-driven tree driver for a few batch shapes (run on the GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb

for game, strat, depth in [(mb.Game.CONNECT4, mb.Uniform(), 0), (mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), 200)]:
    for batch, k, iters, pipe in [(1, 1, 2000, False), (64, 64, 1 << 14, False), (64, 64, 1 << 14, True), (256, 256, 1 << 16, False),
                                  (256, 256, 1 << 16, True), (1024, 1024, 1 << 15, False), (1024, 1024, 1 << 15, True)]:
        with mb.GpuSearch(game, strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k, max_playout_depth=depth,
                          pipelined=pipe, seed=5) as s:
            s.choose_action(mb.initial_state(game))
            t0 = time.perf_counter()
            r = s.choose_action(mb.initial_state(game))
            dt = time.perf_counter() - t0
        print(f"{game.name:>12} B={batch:5d} k={k:5d} iters={iters:6d} pipelined={int(pipe)}: {r.playouts / dt:.3e} playouts/s  {iters / dt:.3e} leaves/s  "
              f"select {r.seconds_select:.3f}s gpu {r.seconds_gpu:.3f}s backprop {r.seconds_backprop:.3f}s nodes {r.nodes} best {r.best_action}")
