#!/usr/bin/env python
"""BASELINE config 5 on one GPU: 2^24 Breakthrough playouts for one move, root-parallel over searcher threads.
usage: config5_probe.py [threads ...]"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb

root = mb.initial_state(mb.Game.BREAKTHROUGH)
for thr in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4]:
    shapes = [(256, 1024, 4), (512, 1024, 4), (256, 2048, 4), (128, 1024, 4)]

    if os.environ.get("SHAPES"):
        shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["SHAPES"].split(",")]
    for bl, kk, depth in shapes:
        iters = int(os.environ.get("TOTAL", str(1 << 24))) // (thr * kk)
        with mb.GpuSearch(mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), max_iterations=iters, batch_leaves=bl, num_rollouts_per_leaf=kk,
                          pipeline_depth=depth, seed=int(os.environ.get("SEED", "7")), num_threads=thr, max_playout_depth=200) as s:
            s.choose_action(root)
            t0 = time.perf_counter()
            for _ in range(3):
                r = s.choose_action(root)
            dt = (time.perf_counter() - t0) / 3
        print(f"threads {thr} batch {bl}x{kk} depth {depth}: {dt*1e3:.2f} ms/move, {r.playouts/dt:.3e} playouts/s, select {r.seconds_select*1e3:.2f} submit {r.seconds_submit*1e3:.2f} wait {r.seconds_gpu*1e3:.2f} backprop {r.seconds_backprop*1e3:.2f} ms")
