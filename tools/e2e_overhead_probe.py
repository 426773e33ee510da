import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, mcts_b200 as mb, bench
w=bench.HEADLINE
g=mb.GpuRollout(mb.Game(w["game"]), mb.EpsilonGreedyMast(0.1), seed=1, max_playout_depth=200)
leaves=g.random_positions(w["plies"],4096,seed=w["pos_seed"])
g.strategy=mb.Uniform(); mast=g.simulate_many(leaves,16,want_mast_delta=True,seed=999).mast_delta.copy(); g.strategy=mb.EpsilonGreedyMast(0.1)
t0=time.perf_counter()
for _ in range(2000): g.set_mast(mast)
print("mr_set_mast: %.1f us per call" % ((time.perf_counter()-t0)/2000*1e6))
import ctypes as C
for k in (4096,):
    for _ in range(3): g.simulate_many(leaves,k,stats=mast,want_mast_delta=True)
    t0=time.perf_counter()
    for i in range(20): r=g.simulate_many(leaves,k,stats=mast,want_mast_delta=True,leaf_id_base=i)
    dt=(time.perf_counter()-t0)/20
    print("e2e step %.3f ms, kernel %.3f ms" % (dt*1e3, r.kernel_ms))
