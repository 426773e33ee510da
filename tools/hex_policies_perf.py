"""Kernel time of every Hex policy: the fill-then-search kernel (uniform) beside the per-ply kernel (hex_step.cuh)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, mcts_b200 as mb
g0 = mb.GpuRollout(mb.Game.HEX11, mb.Uniform(), seed=1)
leaves = g0.random_positions(30, 4096, seed=3)
mast = g0.simulate_many(leaves, 16, want_mast_delta=True, seed=999).mast_delta.copy()
for name, strat, kw in [("uniform(fill-then-search)", mb.Uniform(), {}), ("uniform(per-ply,+mast delta)", mb.Uniform(), dict(want_mast_delta=True)),
                        ("eps-mast", mb.EpsilonGreedyMast(0.1), dict(stats=mast, want_mast_delta=True)), ("decisive-win", mb.DecisiveMoveWin(), {}),
                        ("anti-decisive", mb.DecisiveMove("anti_decisive"), {}), ("lgr(empty table)", mb.EpsilonGreedyLgr(0.1), {})]:
    g = mb.GpuRollout(mb.Game.HEX11, strat, seed=1)
    for i in range(3): r = g.simulate_many(leaves, 1024, leaf_id_base=i*4096, **kw)
    pl = int(r.results["sum_depth"].sum())
    print(f"hex11 {name}: {r.kernel_ms:.3f} ms, {pl/r.kernel_ms/1e6:.2f} Gplies/s, {4096*1024/r.kernel_ms/1e6:.3f} Gplayouts/s, mean plies {pl/(4096*1024):.1f}")
