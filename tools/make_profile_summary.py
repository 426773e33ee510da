#!/usr/bin/env python
"""Writes profiles/SUMMARY_r01.md from profiles/bench_r01.json, bench_reference_r01.json, int_peaks_r01.json and the
ncu summaries.  Run on the CPU box after copying the round's evidence into profiles/."""
import json, re
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"
b = json.loads((P / "bench_r01.json").read_text())
r = json.loads((P / "bench_reference_r01.json").read_text())
pk = json.loads((P / "int_peaks_r01.json").read_text())


def ncu(name):
    f = P / f"ncu_r01_{name}.txt"
    if not f.exists():
        return {}
    out = {}
    for line in f.read_text().splitlines():
        m = re.match(r"\s+(\S+)\s+([0-9.]+)\s*(\S*)", line)
        if m:
            v = float(m.group(2))
            if m.group(1) == "gpu__time_duration.sum":  # normalise to milliseconds
                v *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(m.group(3), 1.0)
            out[m.group(1)] = v
    return out


L = []
L.append("# Round 1 evidence summary (B200, sm_100a)\n")
L.append("All numbers below come from files in this directory; regenerate with `python tools/make_profile_summary.py`.\n")
L.append("## Headline (BASELINE.json configs[1]): Breakthrough 8x8, epsilon-greedy(0.1) MAST, 4096 leaves x 4096 playouts per step\n")
L.append(f"* device-resident: **{b['value']:.3e} playouts/s**, {b['ms_per_step']:.3f} ms/step, {b['plies_per_s']:.3e} plies/s, {b['gpu_launches']} kernel launches in {b['steps']} steps")
L.append(f"* end to end through the C ABI (host buffers): **{b['e2e']['value']:.3e} playouts/s** (H2D {b['e2e']['h2d_bytes_per_step']} B, D2H {b['e2e']['d2h_bytes_per_step']} B per step)")
L.append(f"* reference arm (`bench.py --impl reference`, C restatement of the reference's CPU path, {r['cpu_baseline']['cores']} threads): {r['value']:.3e} playouts/s"
         f" -> e2e / reference = **{b['e2e']['value'] / r['value']:.0f}x**")
rf = b["roofline"]
L.append(f"* integer-issue roofline: achieved {rf['achieved']:.2f} Tiop/s of the measured LOP3 peak {rf['peak']:.2f} Tiop/s = **{rf['frac']:.2f}** (model ops/ply, see DESIGN.md section 6); DRAM traffic {rf['traffic']} B per launch")
L.append(f"* clocks during the run: {b['clocks']}\n")
L.append("## Per game (same batch shape, device-resident)\n")
L.append("| workload | playouts/s | plies/s | mean plies | ms/step | model-ops roofline frac |")
L.append("|---|---|---|---|---|---|")
for k, v in b["per_game"].items():
    if "playouts_per_s" in v and "plies_per_s" in v:
        L.append(f"| {k} | {v['playouts_per_s']:.3e} | {v['plies_per_s']:.3e} | {v['mean_plies']:.1f} | {v['ms_per_step']:.3f} | {v['int_issue_frac']:.2f} |")
s = b["per_game"].get("connect4_uct_100k_per_move")
c = (b.get("cpu_baseline") or {}).get("connect4_uct_100k_per_move")
if s and c and "seconds_per_move" in s:
    L.append(f"\nConfig 1 (Connect4 UCT, ~100k playouts for one move from the empty board): GPU-driven tree driver {s['seconds_per_move'] * 1e3:.2f} ms per move"
             f" ({s['playouts_per_s']:.3e} playouts/s, best move {s['best_action']}); restated reference loop on one CPU core {c['seconds_per_move'] * 1e3:.0f} ms"
             f" ({c['rollouts_per_s_per_core']:.3e} rollouts/s/core, best move {c['best_action']}).\n")
moves = [("hex11_grave_amaf_search", "Config 3 (Hex 11x11, GRAVE over the in-kernel AMAF tables)"),
         ("othello_decisive_tt_search", "Config 4a (Othello, decisive-move rollouts, transposition table on the host)"),
         ("breakthrough_root_parallel_2p24_per_move", "Config 4b / 5 (Breakthrough, 2^24 playouts for one move, one searcher per GPU, root totals merged by one all-reduce)")]
for key, title in moves:
    v = b["per_game"].get(key)
    if v and "seconds_per_move" in v:
        L.append(f"{title}: {v['playouts_per_move']} playouts in {v['seconds_per_move'] * 1e3:.2f} ms per move ({v['playouts_per_s']:.3e} playouts/s through the tree driver).\n")
import glob
scal = []
for f in sorted(glob.glob(str(P / "bench_r01_n*.json"))):
    d = json.loads(Path(f).read_text().strip().splitlines()[-1])
    scal.append((d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
if scal:
    L.append("## Scaling of the headline workload (weak: 4096 x 4096 playouts per GPU per step, one all-reduce per step)\n")
    L.append("| GPUs | playouts/s | ms/step | e2e playouts/s | vs 1 GPU |")
    L.append("|---|---|---|---|---|")
    L.append(f"| 1 | {b['value']:.3e} | {b['ms_per_step']:.3f} | {b['e2e']['value']:.3e} | 1.00 |")
    for n, v, ms, e in sorted(scal):
        L.append(f"| {n} | {v:.3e} | {ms:.3f} | {e:.3e} | {v / b['value']:.2f} |")
    L.append("")
L.append("## Measured integer peaks (`tools/measure_int_peaks.py`, thread-ops per SM per clock at 1965 MHz)\n")
L.append("| op | /SM/clk | T thread-ops/s |")
L.append("|---|---|---|")
for k, v in pk.items():
    if isinstance(v, dict):
        L.append(f"| {k} | {v['per_sm_per_clk_at_max_clock']:.1f} | {v['thread_ops_per_s'] / 1e12:.2f} |")
L.append("\n## ncu counters of the dominant kernels at bench size (`--set full`, one launch)\n")
L.append("| kernel | duration (ms) | threads/inst (of 32) | issue active % | ALU pipe % | FMA pipe % | XU pipe % | warps active % | regs |")
L.append("|---|---|---|---|---|---|---|---|---|")
for name in ["breakthrough8x8_eps_mast", "breakthrough8x8_uniform", "knightthrough8x8_eps_mast", "connect4_uniform_midgame",
             "hex11_uniform_amaf", "hex11_eps_mast_per_ply", "othello_decisive_win"]:
    n = ncu(name)
    if not n:
        continue
    g = lambda key: n.get(key, float("nan"))
    L.append(f"| {name} | {g('gpu__time_duration.sum'):.3f} | {g('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | "
             f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f} | "
             f"{g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f} | "
             f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('launch__registers_per_thread'):.0f} |")
L.append("\n`launches_r01.csv` is the ncu launch list of `bench.py --steps 3 --warmup 3 --headline-only`: the rollout kernel is the only kernel of a step "
         "besides torch's result memsets (< 0.5 % of the step).  `grab_sweep_r01.txt` and `roll_compare_r01.txt` hold the scheduling and unrolling sweeps "
         "the launch parameters were chosen from.  compute-sanitizer is closed on this pool (the tool answered so), so memory-safety evidence is the "
         "bit-exact parity suite itself.\n")
(P / "SUMMARY_r01.md").write_text("\n".join(L) + "\n")
print("\n".join(L))
