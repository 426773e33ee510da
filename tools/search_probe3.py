#!/usr/bin/env python
"""Time split of one move in the tree driver for the Hex GRAVE and Othello decisive + transposition-table configs."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb
from mcts_b200.search import QINIT_INFINITY, SELECT_GRAVE

cfgs = [("hex grave", mb.Game.HEX11, mb.Uniform(), dict(select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=700)),
        ("hex ucb1", mb.Game.HEX11, mb.Uniform(), dict()),
        ("othello dm tt", mb.Game.OTHELLO, mb.DecisiveMoveWin(), dict(use_transpositions=True))]
for name, game, strat, kw in cfgs:
    for batch, k in [(128, 64), (512, 64), (128, 512), (1024, 256)]:
        iters = max(batch, (1 << 18) // k)
        with mb.GpuSearch(game, strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k, pipelined=True, seed=7, **kw) as s:
            s.choose_action(mb.initial_state(game))
            t0 = time.perf_counter()
            r = s.choose_action(mb.initial_state(game))
            dt = time.perf_counter() - t0
        print(f"{name:14s} B={batch:5d} k={k:4d} iters={iters:6d}: {dt*1e3:7.2f} ms/move {r.playouts / dt:.3e} playouts/s  select {r.seconds_select*1e3:.2f}  "
              f"gpu-wait {r.seconds_gpu*1e3:.2f}  backprop {r.seconds_backprop*1e3:.2f} ms  nodes {r.nodes} batches {r.batches}")
