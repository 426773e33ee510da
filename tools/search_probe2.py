#!/usr/bin/env python
"""Where one 2^24-playout Breakthrough move spends its time in the tree driver (run on the GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb

game = mb.Game.BREAKTHROUGH
for batch, k, pipe in [(1024, 1024, False), (1024, 1024, True), (4096, 1024, True), (256, 4096, True), (1024, 4096, True), (4096, 4096, True)]:
    iters = (1 << 24) // k
    with mb.GpuSearch(game, mb.EpsilonGreedyMast(0.1), max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k, max_playout_depth=200,
                      pipelined=pipe, seed=5) as s:
        s.choose_action(mb.initial_state(game))
        t0 = time.perf_counter()
        r = s.choose_action(mb.initial_state(game))
        dt = time.perf_counter() - t0
    print(f"B={batch:5d} k={k:5d} iters={iters:6d} pipelined={int(pipe)}: {dt*1e3:7.2f} ms/move {r.playouts / dt:.3e} playouts/s  "
          f"select {r.seconds_select*1e3:.2f} ms  gpu-wait {r.seconds_gpu*1e3:.2f} ms  backprop {r.seconds_backprop*1e3:.2f} ms  nodes {r.nodes} batches {r.batches}")
