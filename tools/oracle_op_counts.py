#!/usr/bin/env python
"""Operation counts of the reference's arithmetic per ply, measured by running the INSTRUMENTED oracle
(oracle/_build/liboracle_ops.so, -DORC_COUNT_OPS) on the bench workloads ("re-derive ops/ply by
instrumenting the oracle with op counters, never from the kernel's own SASS").  Runs on the CPU; writes
profiles/oracle_ops_r02.json, which bench.py reads for the roofline numerators it prints.

Two numbers per workload:
  ops_per_ply_reference_algorithm : what the reference's OWN algorithm executes (per-piece loops, materialised action
                                    lists, a flood per ply, apply-and-test decisive scans), 32-bit-instruction weights
                                    of oracle_lib.OP_WEIGHTS
  ops_per_ply_model_v1            : SURVEY 8(d)'s bit-parallel restatement, with its two workload-dependent terms taken
                                    from the counters: MAST score evaluations per ply (8 ops each) and Hex flood
                                    iterations per ply (116 ops each)
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench
import oracle_lib as o

N_LEAVES, K = 128, 32
out = {"weights": o.OP_WEIGHTS, "sample": f"{N_LEAVES} leaves x {K} playouts per workload, single thread", "workloads": {}}
for w in bench.all_workloads():
    leaves = o.random_positions(w["game"], w["pos_seed"], w["plies"], N_LEAVES)
    mast = None
    if w["policy"] == 1:
        mast = o.playout_batch(w["game"], leaves, bench.MAST_WARMUP_K, seed=999, max_depth=w["max_depth"], want_mast_delta=True)["mast_delta"]
    r = o.count_ops(w["game"], leaves, K, policy=w["policy"], eps=w["eps"], max_depth=w["max_depth"], seed=4242, mast=mast)
    c, plies = r["counts"], r["plies"]
    entry = {
        "plies": plies,
        "mean_plies": plies / (N_LEAVES * K),
        "counts_per_ply": {k: v / plies for k, v in c.items() if k != "ply"},
        "ops_per_ply_reference_algorithm": r["ops_per_ply"],
        "score_evals_per_ply": c["score_eval"] / plies,
        "flood_iters_per_ply": c["flood_iter"] / plies,
        "mean_legal_actions": c["list_store"] / max(1, plies),
    }
    entry["ops_per_ply_model_v1"] = bench.model_ops_per_ply(w, entry)
    out["workloads"][w["name"]] = entry
    print(f"{w['name']:>40}: reference-algorithm {entry['ops_per_ply_reference_algorithm']:8.1f} ops/ply, model v1 {entry['ops_per_ply_model_v1']:7.1f}, "
          f"mean plies {entry['mean_plies']:6.2f}, legal {entry['mean_legal_actions']:5.1f}, score evals/ply {entry['score_evals_per_ply']:5.2f}, floods/ply {entry['flood_iters_per_ply']:4.2f}")
(ROOT / "profiles" / "oracle_ops_r02.json").write_text(json.dumps(out, indent=1) + "\n")
