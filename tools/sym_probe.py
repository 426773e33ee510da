import sys, time
sys.path.insert(0, "/root/repo")
import mcts_b200 as mb
root = mb.initial_state(mb.Game.OTHELLO)
for name, kw in (("exact", dict(use_transpositions=True)), ("canonical", dict(use_transpositions=True, canonical_symmetry=True)), ("none", {})):
    with mb.GpuSearch(mb.Game.OTHELLO, mb.DecisiveMoveWin(), max_iterations=4096, batch_leaves=128, num_rollouts_per_leaf=256, pipeline_depth=4, seed=7, **kw) as s:
        s.choose_action(root)
        t0 = time.perf_counter()
        for _ in range(3):
            r = s.choose_action(root)
        dt = (time.perf_counter() - t0) / 3
    print(name, f"{dt*1e3:.3f} ms/move", "nodes", r.nodes, "select", f"{r.seconds_select*1e3:.3f}", "depth", r.max_tree_depth)
