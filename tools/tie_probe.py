#!/usr/bin/env python
"""Kernel time of the epsilon-greedy MAST workloads with SYNTHETIC tables of a given density (share of visited actions):
unvisited actions score 1.0, so sparse tables produce ties away from the goal row.  usage: tie_probe.py <game 1|2> [densities...]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import mcts_b200 as mb

game = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dens = [float(x) for x in sys.argv[2:]] or [0.0, 0.5, 0.9, 0.97, 1.0]
g = mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyMast(0.1), seed=12345, max_playout_depth=200)
leaves = g.random_positions(20 if game == 1 else 8, 4096, seed=2)
na = 4096
for d in dens:
    rng = np.random.default_rng(int(d * 1000))
    mast = np.zeros((2, na), dtype=mb.STAT_DTYPE)
    visited = rng.random((2, na)) < d
    visits = rng.integers(20, 4000, size=(2, na))
    score = (rng.random((2, na)) * 2 - 1) * visits * 0.6
    mast["visits"] = np.where(visited, visits, 0)
    mast["score"] = np.where(visited, score.astype(np.int64), 0)
    for i in range(3):
        r = g.simulate_many(leaves, 1024, stats=mast, want_mast_delta=True, leaf_id_base=i * 4096)
    plies = int(r.results["sum_depth"].sum())
    print(f"game {game} density {d:4.2f}: kernel {r.kernel_ms:.3f} ms, {plies / 4096 / 1024:.2f} plies/playout, {plies / r.kernel_ms / 1e6:.2f} Gplies/s")
