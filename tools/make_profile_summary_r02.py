#!/usr/bin/env python
"""Writes profiles/SUMMARY_r02.md from profiles/bench_r02.json, bench_reference_r02.json, bench_r02_n*.json,
ncu_counters_r02.json and oracle_ops_r02.json.  Run on the CPU box after copying the round's evidence into profiles/."""
import glob
import json
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"
last = lambda f: json.loads((P / f).read_text().strip().splitlines()[-1])
b, r = last("bench_r02.json"), last("bench_reference_r02.json")
cnt = json.loads((P / "ncu_counters_r02.json").read_text())
ops = json.loads((P / "oracle_ops_r02.json").read_text())["workloads"]
L = ["# Round 2 evidence summary (B200, sm_100a)\n",
     "Every number below comes from a file in this directory; regenerate with `python tools/make_profile_summary_r02.py`.\n",
     "## Headline (BASELINE.json configs[1]): Breakthrough 8x8, epsilon-greedy(0.1) MAST, 4096 leaves x 4096 playouts per step\n"]
L.append(f"* device-resident, one launch per step: **{b['value']:.3e} playouts/s**, {b['ms_per_step']:.3f} ms/step, {b['plies_per_s']:.3e} plies/s, "
         f"{b['gpu_launches']} kernel launches in {b['steps']} steps; clocks {b['clocks']}")
L.append(f"* end to end through the C ABI (host buffers, three slots in flight): **{b['e2e']['value']:.3e} playouts/s** "
         f"(H2D {b['e2e']['h2d_bytes_per_step']} B, D2H {b['e2e']['d2h_bytes_per_step']} B per step)")
L.append(f"* reference arm (`bench.py --impl reference`, C restatement of the reference's CPU path, {r['cpu_baseline']['cores']} threads, the SAME "
         f"{r['config']['leaves_per_step_per_gpu']} x {r['config']['playouts_per_leaf']} step: config identical = {r['config'] == b['config']}): "
         f"{r['value']:.3e} playouts/s -> e2e / reference = **{b['e2e']['value'] / r['value']:.0f}x**")
rf = b["roofline"]
m = rf["measured"] or {}
L.append(f"* roofline (integer issue): model v1 {rf['ops_per_ply_model_v1']:.0f} ops/ply (MAST term counted by the instrumented oracle) x plies/s = "
         f"{rf['achieved']:.2f} Tiop/s of the live LOP3 peak {rf['peak']:.2f} = **{rf['frac']:.2f}**; the reference's own algorithm executes "
         f"{rf['ops_per_ply_reference_algorithm']:.0f} ops/ply; the KERNEL executes {m.get('thread_inst_per_ply', float('nan')):.0f} thread-instructions per ply at "
         f"{m.get('lanes_per_inst', float('nan')):.1f} of 32 lanes: thread-issue utilisation **{m.get('thread_issue_frac', float('nan')):.2f}**, issue slots "
         f"{m.get('issue_slot_frac', float('nan')):.2f}, ALU pipe {m.get('alu_pipe_frac_ncu', float('nan')):.2f}, FMA pipe {m.get('fma_pipe_frac_ncu', float('nan')):.2f} (ncu); "
         f"DRAM {rf['traffic']} B per launch\n")
L.append("## Per workload (4096 x 4096 playouts, device-resident): model next to measured\n")
L.append("| workload | playouts/s | plies/s | mean plies | ms/step | model ops/ply | model frac | kernel thread-inst/ply | lanes/inst | thread-issue frac | ALU pipe | regs |")
L.append("|---|---|---|---|---|---|---|---|---|---|---|---|")
rows = dict(b["per_game"])
rows[b["config"]["workload"]] = {"playouts_per_s": b["value"], "plies_per_s": b["plies_per_s"], "mean_plies": b["mean_plies_per_playout"], "ms_per_step": b["ms_per_step"],
                                 "model_ops_per_ply": rf["ops_per_ply_model_v1"], "model_int_issue_frac": rf["frac"], "measured": rf["measured"]}
for k, v in rows.items():
    if "playouts_per_s" not in v:
        continue
    mm = v.get("measured") or {}
    g = lambda key, fmt: format(mm[key], fmt) if key in mm and mm[key] is not None else "-"
    L.append(f"| {k} | {v['playouts_per_s']:.3e} | {v['plies_per_s']:.3e} | {v['mean_plies']:.1f} | {v['ms_per_step']:.3f} | {v['model_ops_per_ply']:.0f} | "
             f"{v['model_int_issue_frac']:.2f} | {g('thread_inst_per_ply', '.0f')} | {g('lanes_per_inst', '.1f')} | {g('thread_issue_frac', '.2f')} | {g('alu_pipe_frac_ncu', '.2f')} | {g('registers', '.0f')} |")
if "hex11_eps_mast" in cnt:
    h = cnt["hex11_eps_mast"]
    L.append(f"\nHex 11x11 epsilon-MAST (fill-then-search since this round, `hex_rollout_kernel<1, 2>`): {h['kernel_ms_under_ncu']:.1f} ms per 2^24 playouts under ncu "
             f"(round 1, per-ply kernel: 53.4 ms), {h['lanes_per_inst']:.1f} lanes per instruction (11.9), {h['registers']:.0f} registers (109), "
             f"{h['thread_inst_per_ply']:.0f} thread-instructions per ply.\n")
L.append("## SURVEY 8(d) opening-depth matrix (playouts/s; plies/s; mean plies per playout)\n")
for base, col in b["depth_matrix"].items():
    cells = []
    for plies, v in col.items():
        cells.append(f"{plies} plies: {v['playouts_per_s']:.2e}; {v['plies_per_s']:.2e}; {v['mean_plies']:.1f}" if "error" not in v else f"{plies} plies: {v['error']}")
    L.append(f"* **{base}** — " + " | ".join(cells))
L.append("\n## Searches through the tree driver (one B200; host time split per move)\n")
L.append("| config | batch x k | depth | threads | ms/move | playouts/s | below a root child | select | submit | wait | backprop (ms) |")
L.append("|---|---|---|---|---|---|---|---|---|---|---|")
for name, v in b["search"].items():
    for x in (v if isinstance(v, list) else [v]):
        if "error" in x:
            L.append(f"| {name} | error: {x['error']} |")
            continue
        L.append(f"| {name} | {x['batch_leaves']} x {x['playouts_per_leaf']} | {x['pipeline_depth']} | {x.get('searcher_threads', x.get('searcher_threads_per_rank'))} | "
                 f"{x['seconds_per_move'] * 1e3:.3f} | {x['playouts_per_s']:.3e} | {x['root_child_visit_share']:.4f} | {x['seconds_select'] * 1e3:.2f} | "
                 f"{x['seconds_submit'] * 1e3:.2f} | {x['seconds_gpu_wait'] * 1e3:.2f} | {x['seconds_backprop'] * 1e3:.2f} |")
c1 = (b.get("cpu_baseline") or {}).get("connect4_uct_100k_per_move")
if c1:
    L.append(f"\nConfig 1 on one CPU core (restated reference loop, k = 1): {c1['seconds_per_move'] * 1e3:.0f} ms per move, {c1['rollouts_per_s_per_core']:.3e} rollouts/s/core.\n")
scal = []
for f in sorted(glob.glob(str(P / "bench_r02_n*.json"))):
    d = json.loads(Path(f).read_text().strip().splitlines()[-1])
    scal.append(d)
if scal:
    L.append("## Scaling of the headline workload (weak: 4096 x 4096 playouts per GPU per step; MAST delta exchanged every step)\n")
    L.append("| GPUs | playouts/s | ms/step | vs N x 1 GPU | e2e playouts/s | e2e vs N x 1 GPU | merge self-checks |")
    L.append("|---|---|---|---|---|---|---|")
    L.append(f"| 1 | {b['value']:.3e} | {b['ms_per_step']:.3f} | 1.000 | {b['e2e']['value']:.3e} | 1.000 | - |")
    for d in sorted(scal, key=lambda d: d["n_gpus"]):
        n = d["n_gpus"]
        L.append(f"| {n} | {d['value']:.3e} | {d['ms_per_step']:.3f} | {d['value'] / (n * b['value']):.3f} | {d['e2e']['value']:.3e} | {d['e2e']['value'] / (n * b['e2e']['value']):.3f} | "
                 f"{'all ok' if all(d.get('self_check', {}).values()) else d.get('self_check')} |")
        for key in ("breakthrough_root_parallel_2p24_per_move", "breakthrough_root_parallel_2p24_one_searcher_per_rank",
                    "breakthrough_root_parallel_2p24_two_searchers_per_rank"):
            s5 = d["search"].get(key, {})
            if "seconds_per_move" in s5:
                one = b["search"].get("breakthrough_root_parallel_2p24_per_move", {}).get("seconds_per_move")
                L.append(f"|  | config 5 (2^24 playouts, one move, {s5['ranks']} ranks x {s5['searcher_threads_per_rank']} searchers): {s5['seconds_per_move'] * 1e3:.2f} ms"
                         + (f" ({one / s5['seconds_per_move']:.2f}x the one-GPU move)" if one else "") + f", {s5['playouts_per_s']:.3e} playouts/s, "
                         f"{s5['root_child_visit_share']:.4f} of the playouts below a root child"
                         + (f"; the search call alone {s5['seconds_search_call_only'] * 1e3:.2f} ms" if "seconds_search_call_only" in s5 else "") + " | | | | | |")
    L.append("")
L.append("## Counted operations of the reference's arithmetic (instrumented oracle, `oracle_ops_r02.json`)\n")
L.append("| workload | reference-algorithm ops/ply | model v1 ops/ply | MAST score evals/ply | flood iterations/ply | list stores/ply |")
L.append("|---|---|---|---|---|---|")
for k in [b["config"]["workload"]] + list(b["per_game"].keys()):
    if k in ops:
        o = ops[k]
        L.append(f"| {k} | {o['ops_per_ply_reference_algorithm']:.0f} | {o['ops_per_ply_model_v1']:.0f} | {o['score_evals_per_ply']:.1f} | {o['flood_iters_per_ply']:.2f} | {o['mean_legal_actions']:.1f} |")
L.append("\n`launches_r02.csv` is the ncu launch list of `bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline`: per step one `rollout_kernel<Breakthrough, 1, 2>` "
         "launch (3.3 ms) plus torch's result / delta memsets (2 us each) and the untimed 256 MiB L2 flush; the e2e arm's launches are the same kernel with the fused epilogue. "
         "`ncu_r02_<workload>.txt` are the `ncu --set full` summaries at bench size, `ncu_r02_breakthrough8x8_eps_mast_lines.txt` the per-source-line breakdown of the headline "
         "kernel, `gputests_2gpu_r02.log` the GPU suite on a two-GPU box (244 passed, none skipped), `oracle_stress_full_r02.txt` the reference's 100 000-game Othello sweep at full size.\n")
(P / "SUMMARY_r02.md").write_text("\n".join(L) + "\n")
print("\n".join(L))
