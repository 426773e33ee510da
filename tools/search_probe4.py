#!/usr/bin/env python
"""Search throughput of the tree driver on the BASELINE configs, over batch shape x pipeline depth x searcher threads
(run on the GPU box).  Prints one line per shape with the host-time split the driver reports."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb
from mcts_b200.search import QINIT_INFINITY, SELECT_GRAVE


def run(name, game, strat, total, batch, k, depth, threads=1, reps=3, **kw):
    iters = max(1, total // (k * threads))
    with mb.GpuSearch(game, strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=k, pipeline_depth=depth,
                      seed=7, num_threads=threads, **kw) as s:
        root = mb.initial_state(game)
        s.choose_action(root)
        t0 = time.perf_counter()
        for _ in range(reps):
            r = s.choose_action(root)
        dt = (time.perf_counter() - t0) / reps
    rootv = int(r.children["visits"].sum())
    print(f"{name:>22} B={batch:5d} k={k:5d} depth={depth} thr={threads}: {dt * 1e3:8.3f} ms/move {r.playouts / dt:.3e} playouts/s "
          f"root-child visits {100.0 * rootv / r.playouts:5.1f}% | select {r.seconds_select * 1e3:7.3f} submit {r.seconds_submit * 1e3:7.3f} "
          f"wait {r.seconds_gpu * 1e3:7.3f} backprop {r.seconds_backprop * 1e3:7.3f} ms | nodes {r.nodes} batches {r.batches} best {r.best_action}",
          flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "c4"):
    for batch, k in [(64, 64), (32, 256), (64, 256), (16, 1024), (32, 1024)]:
        for depth in (1, 2, 4):
            run("connect4_uct_100k", mb.Game.CONNECT4, mb.Uniform(), 100_000, batch, k, depth)
    for batch, k in [(64, 256), (32, 1024)]:
        run("connect4_uct_100k", mb.Game.CONNECT4, mb.Uniform(), 100_000, batch, k, 4, threads=4)
if which in ("all", "hex"):
    for batch, k in [(128, 64), (128, 256), (256, 256), (128, 1024)]:
        for depth in (1, 2, 4):
            run("hex11_grave", mb.Game.HEX11, mb.Uniform(), 4096 * k if k <= 256 else 1 << 21, batch, k, depth, select=SELECT_GRAVE,
                q_init=QINIT_INFINITY, grave_threshold=700)
    run("hex11_grave", mb.Game.HEX11, mb.Uniform(), 1 << 22, 128, 256, 4, threads=4, select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=700)
if which in ("all", "bt"):
    for batch, k in [(256, 1024), (128, 4096), (512, 1024)]:
        for depth in (2, 4):
            for thr in (1, 2, 4):
                run("breakthrough_2p24_mast", mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), 1 << 24, batch, k, depth, threads=thr, reps=2,
                    max_playout_depth=200)
    run("breakthrough_2p24_uniform", mb.Game.BREAKTHROUGH, mb.Uniform(), 1 << 24, 256, 1024, 4, threads=4, reps=2, max_playout_depth=200)
if which in ("all", "oth"):
    for depth in (1, 4):
        run("othello_decisive_tt", mb.Game.OTHELLO, mb.DecisiveMoveWin(), 4096 * 64, 128, 64, depth, use_transpositions=True)
        run("othello_decisive_tt", mb.Game.OTHELLO, mb.DecisiveMoveWin(), 4096 * 256, 128, 256, depth, use_transpositions=True)
