#!/usr/bin/env python
"""Debug probe: one small uniform batch per game through the C ABI, compared with the oracle."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np

import mcts_b200 as mb
import oracle_lib as o

k = int(sys.argv[1]) if len(sys.argv) > 1 else 96
for game in (o.CONNECT4, o.BREAKTHROUGH):
    leaves = np.concatenate([o.initial_state(game), o.random_positions(game, 3, 10, 3)])
    ref = o.playout_batch(game, leaves, k, policy=o.UNIFORM, seed=7, want_trace=True, max_trace=64, want_final=True)
    with mb.GpuRollout(mb.Game(game), mb.Uniform(), seed=7) as g:
        got = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, want_trace=True, want_final=True, max_trace=64)
    print(game, "results equal:", got.results.tobytes() == ref["results"].tobytes(), flush=True)
