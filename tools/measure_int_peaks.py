#!/usr/bin/env python
"""Measures the integer-issue peaks of the GPU (the roofline denominators of this path) with the
library's own micro-benchmark kernels and writes them as JSON.  Run on the B200 box."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mcts_b200 as mb

KINDS = ["lop3", "iadd3", "shf", "imad", "popc", "lop3+imad", "imad.hi", "imad.wide+lop3", "shf.imm", "isetp+sel"]
out = {}
with mb.GpuRollout(mb.Game.CONNECT4) as g:
    for i, name in enumerate(KINDS):
        ops, mhz = g.measure_int_peak(i)
        out[name] = {"thread_ops_per_s": ops, "per_sm_per_clk_at_max_clock": ops / (148 * mhz * 1e6)}
    out["sm_max_mhz"] = mhz
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1) + "\n")
