#!/usr/bin/env python
"""End-to-end rate of the headline workload through mr_submit / mr_wait with 1..4 slots in flight (one GPU)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
import mcts_b200 as mb

w = bench.HEADLINE
n, k = 4096, 4096
g = mb.GpuRollout(mb.Game(w["game"]), mb.EpsilonGreedyMast(w["eps"]), seed=12345, max_playout_depth=w["max_depth"])
leaves = g.random_positions(w["plies"], n, seed=w["pos_seed"])
g.strategy = mb.Uniform()
mast = g.simulate_many(leaves, bench.MAST_WARMUP_K, want_mast_delta=True, seed=999).mast_delta.copy()
g.strategy = mb.EpsilonGreedyMast(w["eps"])
kw = dict(stats=mast, want_mast_delta=True, tables_in_slots=True)
for slots in (1, 2, 3, 4):
    def run(count, base):
        for j in range(min(slots - 1, count)):
            g.submit(leaves, k, slot=j % slots, leaf_id_base=(base + j) * n, **kw)
        for i in range(count):
            if i + slots - 1 < count:
                g.submit(leaves, k, slot=(i + slots - 1) % slots, leaf_id_base=(base + i + slots - 1) * n, **kw)
            r = g.wait(i % slots)
        return r
    run(4, 0)
    best = 0
    for rep in range(3):
        t0 = time.perf_counter(); run(20, 100 * (rep + 1)); dt = time.perf_counter() - t0
        best = max(best, n * k * 20 / dt)
    print(f"{slots} slot(s) in flight: {best:.4e} playouts/s ({n * k / best * 1e3:.3f} ms per step)")
