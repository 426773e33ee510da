#!/usr/bin/env python
"""Builds profiles/ncu_counters_r02.json from the `ncu --set full` summaries of the bench workloads
(gpurun_out/ncusum_<workload>.txt, written by tools/ncu_full.sh at bench size) and the plies of the profiled launch
(gpurun_out/plainfull_<workload>.log).  bench.py multiplies these per-ply instruction counts -- a property of the
workload and the seed, not of the run -- with the live plies/s to report measured issue utilisation beside the model."""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "gpurun_out"
out_path = ROOT / "profiles" / "ncu_counters_r02.json"
out = json.loads(out_path.read_text()) if out_path.exists() else {}
for f in sorted(G.glob("ncusum_*.txt")):
    name = f.stem[len("ncusum_"):]
    vals = {}
    for line in f.read_text().splitlines():
        m = re.match(r"\s+(\S+)\s+([0-9.eE+-]+)\s*(\S*)", line)
        if m:
            v = float(m.group(2))
            unit = m.group(3)
            if m.group(1).startswith("dram__bytes"):
                v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
            if m.group(1) == "gpu__time_duration.sum":
                v *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1.0)
            vals[m.group(1)] = v
    log = G / f"plainfull_{name}.log"
    m = re.search(r"(\d+)x(\d+) playouts, (\d+) plies", log.read_text()) if log.exists() else None
    if not m or "smsp__inst_executed.sum" not in vals:
        print(f"skip {name}: incomplete", file=sys.stderr)
        continue
    plies = int(m.group(3))
    warp_inst = vals["smsp__inst_executed.sum"]
    lanes = vals["smsp__thread_inst_executed_per_inst_executed.ratio"]
    out[name] = {
        "leaves": int(m.group(1)), "playouts_per_leaf": int(m.group(2)), "plies": plies,
        "kernel_ms_under_ncu": vals.get("gpu__time_duration.sum"),
        "warp_inst": warp_inst, "lanes_per_inst": lanes,
        "warp_inst_per_ply": warp_inst / plies, "thread_inst_per_ply": warp_inst * lanes / plies,
        "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": vals.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": vals.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "xu_pipe_pct": vals.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": vals.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": vals.get("launch__registers_per_thread"),
        "dram_bytes": vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0),
        "source": f"profiles/ncu_r02_{name}.txt",
    }
    print(name, json.dumps(out[name]))
out_path.write_text(json.dumps(out, indent=1) + "\n")
