#!/usr/bin/env python
"""One workload, a few launches: the command ncu wraps.  usage: profile_one.py <workload-name> [n_leaves] [k] [launches]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import bench
import mcts_b200 as mb

name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
w = next(x for x in [bench.HEADLINE] + bench.OTHERS + bench.PROFILE_ONLY if x["name"] == name)
strat = {0: mb.Uniform(), 1: mb.EpsilonGreedyMast(w["eps"]), 2: mb.DecisiveMoveWin()}[w["policy"]]
g = mb.GpuRollout(mb.Game(w["game"]), strat, seed=12345, max_playout_depth=w["max_depth"])
leaves = g.random_positions(w["plies"], n, seed=w["pos_seed"])
mast = None
if w["policy"] == 1:
    g.strategy = mb.Uniform()
    mast = g.simulate_many(leaves, bench.MAST_WARMUP_K, want_mast_delta=True, seed=999).mast_delta.copy()
    g.strategy = strat
for i in range(launches):
    r = g.simulate_many(leaves, k, stats=mast, want_mast_delta=(w["policy"] == 1), want_amaf=w["amaf"], leaf_id_base=i * n)
plies = int(r.results["sum_depth"].sum())
print(f"{name}: {n}x{k} playouts, {plies} plies, kernel {r.kernel_ms:.3f} ms, {plies / r.kernel_ms / 1e6:.2f} Gplies/s")
