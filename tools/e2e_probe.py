#!/usr/bin/env python
"""Where does the host-side time of one C-ABI batch go?  (run on the GPU box)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mcts_b200 as mb

g = mb.GpuRollout(mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), seed=1, max_playout_depth=200)
leaves = g.random_positions(20, 4096, seed=2)
g.strategy = mb.Uniform()
mast = g.simulate_many(leaves[:64], 64, want_mast_delta=True).mast_delta.copy()
g.strategy = mb.EpsilonGreedyMast(0.1)
for rep in range(5):
    t0 = time.perf_counter(); g.submit(leaves, 4096, stats=mast, want_mast_delta=True, leaf_id_base=rep)
    t1 = time.perf_counter(); r = g.wait()
    t2 = time.perf_counter()
    print(f"submit {1e3*(t1-t0):.2f} ms  wait {1e3*(t2-t1):.2f} ms  kernel {r.kernel_ms:.2f} ms")
