#!/usr/bin/env python
"""bench.py -- playouts/sec of the MCTS simulation phase on B200 (the BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic leaves: N_LEAVES leaf states x K
playouts each, every playout run to its end on the GPU, per-leaf win/draw/depth statistics and the
MAST table delta reduced in-kernel.  The headline workload is BASELINE.json configs[1]
(Breakthrough 8x8, epsilon-greedy MAST rollouts, batched leaves on one B200); the other games are
timed for a few steps each and reported under "per_game".

  value  : whole-job playouts/s with the leaves already resident in HBM (device-timed, CUDA events
           on the launching stream, max over ranks)
  e2e    : the same metric through the reference-facing C ABI call (mr_submit + mr_wait) with HOST
           buffers: pinned staging, H2D of leaves + MAST snapshot and D2H of results + MAST delta inside
           the timed region
  N > 1  : one process per GPU (torchrun); every rank runs its own batch (weak scaling) and the
           ranks merge the MAST delta and the per-root-child statistics with one NCCL all-reduce per step

`--impl reference` times the reference's CPU playout path instead (the plain-C restatement in oracle/,
all host threads; the Rust reference itself cannot be built in this image) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "playouts/sec per game at 1/2/4/8 B200 vs reference CPU rollouts (cores stated)"
UNIT = "playouts/s"

# (game, policy, opening plies, seed) -- SURVEY.md section 8(d)
HEADLINE = dict(name="breakthrough8x8_eps_mast", game=1, policy=1, plies=20, pos_seed=2, max_depth=200, eps=0.1, amaf=False)
OTHERS = [
    dict(name="connect4_uniform_midgame", game=0, policy=0, plies=12, pos_seed=1, max_depth=0, eps=0.0, amaf=False),
    dict(name="knightthrough8x8_eps_mast", game=2, policy=1, plies=20, pos_seed=2, max_depth=200, eps=0.1, amaf=False),
    # (after 20 random plies a Knightthrough game is almost over: 2.4 plies per playout; 8-ply openings show the per-ply rate)
    dict(name="knightthrough8x8_eps_mast_opening8", game=2, policy=1, plies=8, pos_seed=2, max_depth=200, eps=0.1, amaf=False),
    dict(name="othello_decisive_win", game=3, policy=2, plies=30, pos_seed=4, max_depth=0, eps=0.0, amaf=False),
    dict(name="hex11_uniform_amaf", game=4, policy=0, plies=30, pos_seed=3, max_depth=0, eps=0.0, amaf=True),
    dict(name="breakthrough8x8_uniform", game=1, policy=0, plies=20, pos_seed=2, max_depth=0, eps=0.0, amaf=False),
]
# not part of the bench line: extra workloads tools/profile_one.py can name (ncu captures, A/B timing)
PROFILE_ONLY = [
    dict(name="hex11_eps_mast", game=4, policy=1, plies=30, pos_seed=3, max_depth=0, eps=0.1, amaf=False),
    dict(name="hex11_decisive_win", game=4, policy=2, plies=30, pos_seed=3, max_depth=0, eps=0.0, amaf=False),
    dict(name="othello_eps_mast", game=3, policy=1, plies=30, pos_seed=4, max_depth=0, eps=0.1, amaf=False),
    dict(name="connect4_eps_mast", game=0, policy=1, plies=12, pos_seed=1, max_depth=0, eps=0.1, amaf=False),
]
# ops per ply of a bit-parallel restatement of the reference's arithmetic (model v1)
OPS_PER_PLY = {
    "connect4_uniform_midgame": 125.0,
    "breakthrough8x8_uniform": 160.0,
    "breakthrough8x8_eps_mast": None,  # 160 + 8 * n_legal on the greedy plies; n_legal measured below
    "knightthrough8x8_eps_mast": None,  # 220 + 8 * n_legal on the greedy plies
    "knightthrough8x8_eps_mast_opening8": None,
    "othello_decisive_win": 750.0,
    "hex11_uniform_amaf": 65.0 + 116.0 * 2.9,
}
# DRAM bytes per launch of the headline kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the
# `ncu --set full` capture summarised in profiles/ncu_r01_breakthrough8x8_eps_mast.txt
NCU_TRAFFIC_BYTES = {"breakthrough8x8_eps_mast": 10414848 + 364544}  # profiles/ncu_r01_breakthrough8x8_eps_mast.txt
MAST_WARMUP_K = 16  # uniform playouts per leaf that populate the MAST table before the timed steps
MEAN_LEGAL = {1: 25.7, 2: 47.8}  # BASELINE.md section 4 (oracle-measured mean branching)


def ops_per_ply(w) -> float:
    v = OPS_PER_PLY[w["name"]]
    if v is None:
        base = 160.0 if w["game"] == 1 else 220.0
        v = base + 8.0 * MEAN_LEGAL[w["game"]] * (1.0 - w["eps"])
    return v


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference's playout path
# --------------------------------------------------------------------------------------------


def cpu_sample(w, n_leaves, k, threads, seed):
    """Times oracle playouts for workload `w` on a bounded sample; returns (playouts/s, plies/s)."""
    import oracle_lib as o

    leaves = o.random_positions(w["game"], w["pos_seed"], w["plies"], n_leaves)
    mast = None
    if w["policy"] == 1:
        # a steady-state table: the delta of one uniform batch over the leaves (the same warm-up the GPU arm does)
        warm = o.playout_batch(w["game"], leaves, MAST_WARMUP_K, seed=999, max_depth=w["max_depth"], want_mast_delta=True, threads=threads)
        mast = warm["mast_delta"]
    t0 = time.perf_counter()
    out = o.playout_batch(
        w["game"], leaves, k, policy=w["policy"], eps=w["eps"], max_depth=w["max_depth"], seed=seed, mast=mast,
        want_amaf=w["amaf"], want_mast_delta=(w["policy"] == 1), threads=threads,
    )
    dt = time.perf_counter() - t0
    plies = int(out["results"]["sum_depth"].sum())
    return n_leaves * k / dt, plies / dt, dt


def cpu_sample_size(w, threads, target_s):
    """Bounded sample of the 4096 x 4096 workload that takes about `target_s` seconds on this host."""
    pps, _, _ = cpu_sample(w, 64, 64, threads, 1)
    want = max(4096.0, pps * target_s)
    n = 4096 if want >= 4096 * 64 else 256
    k = int(min(4096, max(32, want // n)))
    return n, k


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import oracle_lib as o

    o.lib()
    threads = o.host_threads()
    w = HEADLINE
    for _ in range(args.warmup):
        cpu_sample(w, 16, 32, threads, 7)
    # bounded sample of the 4096 x 4096 workload: the whole --steps run takes about a minute
    budget_s = float(os.environ.get("BENCH_REFERENCE_SECONDS", "60"))  # whole timed run; the CPU test suite shrinks it
    n_leaves, k = cpu_sample_size(w, threads, max(0.2, budget_s / max(1, args.steps)))
    times, rates = [], []
    for s in range(args.steps):
        pps, _, dt = cpu_sample(w, n_leaves, k, threads, 1000 + s)
        rates.append(pps)
        times.append(dt)
    value = n_leaves * k * len(times) / sum(times)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": workload_config(w, n_leaves, k),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_leaves} leaves x {k} playouts per step of the {w['name']} workload, {threads} pthreads",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(w, n_leaves, k):
    return {
        "workload": w["name"],
        "game": ["connect4", "breakthrough", "knightthrough", "othello", "hex11"][w["game"]],
        "policy": ["uniform", "eps_greedy_mast", "decisive_win"][w["policy"]],
        "epsilon": w["eps"],
        "leaves_per_step_per_gpu": n_leaves,
        "playouts_per_leaf": k,
        "max_playout_depth": w["max_depth"],
        "openings": f"{w['plies']} uniformly random plies from the initial position (seed {w['pos_seed']})",
        "l2_flush": "256 MiB memset between timed steps (untimed); inputs are ~0.2 MB",
    }


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------


def run_gpu(args):
    import torch

    import mcts_b200 as mb
    from mcts_b200 import root_merge

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    else:
        dist = None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_leaves, k = args.leaves, args.playouts
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # a real (non-legacy) stream: the kernels, torch's memsets, NCCL and the timing events all go to it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def time_workload(w, steps, warmup, with_e2e, sampler=None):
        strat = {0: mb.Uniform(), 1: mb.EpsilonGreedyMast(w["eps"]), 2: mb.DecisiveMoveWin()}[w["policy"]]
        g = mb.GpuRollout(mb.Game(w["game"]), strat, devices=[local], seed=12345 + rank, max_playout_depth=w["max_depth"])
        leaves = g.random_positions(w["plies"], n_leaves, seed=w["pos_seed"])
        na = g.na
        mast = None
        if w["policy"] == 1:
            saved, g.strategy = g.strategy, mb.Uniform()
            warm = g.simulate_many(leaves, MAST_WARMUP_K, want_mast_delta=True, seed=999)
            g.strategy = saved
            mast = warm.mast_delta.copy()
            g.set_mast(mast)
        want_delta = w["policy"] == 1
        d_leaves = torch.from_numpy(leaves.view(np.uint8).reshape(-1).copy()).to(dev)
        d_results = torch.zeros(n_leaves * 24, dtype=torch.uint8, device=dev)
        # what the ranks exchange once per step lives in ONE buffer: the MAST delta followed by the per-root-child
        # (visits, score[2]) totals -- one all-reduce instead of two (latency-bound: <= 64 KB)
        n_delta = 2 * na * 2 if want_delta else 0
        d_merge = torch.zeros(n_delta + 48 * 3, dtype=torch.int32, device=dev)
        d_delta = d_merge[:n_delta] if want_delta else None
        d_root = d_merge[n_delta:]
        d_amaf = torch.zeros(n_leaves * 2 * na * 2, dtype=torch.int32, device=dev) if w["amaf"] else None
        launches0 = g.kernel_launches()

        def launch(i):
            d_results.zero_()
            if d_delta is not None:
                d_delta.zero_()
            if d_amaf is not None:
                d_amaf.zero_()
            g.launch_device(
                d_leaves.data_ptr(), d_results.data_ptr(), n_leaves, k, stream=stream.cuda_stream,
                d_amaf=d_amaf.data_ptr() if d_amaf is not None else 0,
                d_mast_delta=d_delta.data_ptr() if d_delta is not None else 0,
                leaf_id_base=root_merge.leaf_id_base(rank, world, i, n_leaves),
            )

        def step(i):
            launch(i)
            if dist is not None:
                # the path's one exchange step: merge MAST delta + per-root-child stats over NVLink (NCCL)
                root_merge.merge_sum_(d_merge)

        if sampler:
            sampler.start()  # before the warm-up: the timed region of this path is tens of milliseconds
        for i in range(warmup):
            step(i)
        barrier()
        evs = []
        launches_before = g.kernel_launches()
        for i in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(warmup + i)
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        launches = g.kernel_launches() - launches_before
        clocks = None
        if sampler:
            # the timed region lasts tens of milliseconds and nvidia-smi reports every ~20 ms: keep the identical load
            # running (untimed) for another 0.4 s so that the clock / throttle-reason record has a dozen samples under load
            t_end = time.perf_counter() + 0.4
            i = 0
            while time.perf_counter() < t_end:
                launch(warmup + steps + i)  # this rank's kernels only: the other ranks do not take part
                i += 1
                if i % 8 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            clocks = sampler.stop()
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(step_ms)
        if dist is not None:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        res = d_results.cpu().numpy().view(mb.RESULT_DTYPE)
        plies_per_step = int(res["sum_depth"].sum())
        done = res["wins"][:, 0].astype(np.int64) + res["wins"][:, 1] + res["draws"]
        assert (done == k).all(), "kernel did not complete every playout"
        out = {
            "total_ms": total_ms,
            "ms_per_step": total_ms / steps,
            "playouts_per_s": world * n_leaves * k * steps / (total_ms * 1e-3),
            "plies_per_s": world * plies_per_step * steps / (total_ms * 1e-3),
            "mean_plies": plies_per_step / (n_leaves * k),
            "launches": launches,
            "clocks": clocks,
        }
        if with_e2e:
            # the call a user makes: host buffers in, host results out, through the C ABI
            for i in range(2):
                g.simulate_many(leaves, k, stats=mast, want_mast_delta=want_delta, want_amaf=w["amaf"], leaf_id_base=i)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                r = g.simulate_many(leaves, k, stats=mast, want_mast_delta=want_delta, want_amaf=w["amaf"],
                                    leaf_id_base=root_merge.leaf_id_base(rank, world, i, n_leaves))
                if dist is not None and want_delta:
                    root_merge.merge_mast_delta(r.mast_delta, device=dev)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            h2d = n_leaves * 40 + (2 * na * 2 if mast is not None else 0)  # leaves + u16 rank table
            d2h = n_leaves * 24 + (2 * na * 8 if want_delta else 0) + (n_leaves * 2 * na * 8 if w["amaf"] else 0)
            out["e2e"] = {"value": world * n_leaves * k * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        out["int_peak"] = g.measure_int_peak(0)
        g.close()
        return out

    sampler = ClockSampler(local) if rank == 0 else None
    head = time_workload(HEADLINE, args.steps, args.warmup, True, sampler)

    per_game = {}
    if not args.headline_only:
        for w in OTHERS:
            try:
                r = time_workload(w, max(2, min(args.steps, 3)), 3, False)
                per_game[w["name"]] = {
                    "playouts_per_s": r["playouts_per_s"], "plies_per_s": r["plies_per_s"], "mean_plies": r["mean_plies"],
                    "ms_per_step": r["ms_per_step"],
                    "int_issue_frac": r["plies_per_s"] * ops_per_ply(w) / (world * r["int_peak"][0]),
                }
            except Exception as e:  # a game that fails must not hide the headline
                per_game[w["name"]] = {"error": str(e)}

    # BASELINE.json config 1 through the whole search loop: Connect4 UCT (UCB1, c = sqrt 2, expand_threshold 1),
    # uniform rollouts, ~100k playouts for one move from the empty board, GPU rollouts behind the tree driver
    if not args.headline_only:
        try:
            with mb.GpuSearch(mb.Game.CONNECT4, mb.Uniform(), max_iterations=1600, batch_leaves=64, num_rollouts_per_leaf=64,
                              pipelined=True, seed=7, devices=[local]) as srch:
                srch.choose_action(mb.initial_state(mb.Game.CONNECT4))
                t0 = time.perf_counter()
                reps = 5
                for _ in range(reps):
                    sr = srch.choose_action(mb.initial_state(mb.Game.CONNECT4))
                dt = (time.perf_counter() - t0) / reps
            per_game["connect4_uct_100k_per_move"] = {
                "playouts_per_move": sr.playouts, "seconds_per_move": dt, "playouts_per_s": sr.playouts / dt,
                "best_action": sr.best_action, "batch_leaves": 64, "playouts_per_leaf": 64, "pipelined": True,
            }
        except Exception as e:
            per_game["connect4_uct_100k_per_move"] = {"error": str(e)}

    # BASELINE.json configs 2-3 through the search loop (one move each, this rank's GPU): Hex 11x11 GRAVE over the
    # in-kernel AMAF tables; Othello with decisive-move rollouts and the transposition table on the host
    if not args.headline_only:
        from mcts_b200.search import QINIT_INFINITY, SELECT_GRAVE

        searches = {
            "hex11_grave_amaf_search": dict(game=mb.Game.HEX11, strat=mb.Uniform(), kw=dict(select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=700)),
            "othello_decisive_tt_search": dict(game=mb.Game.OTHELLO, strat=mb.DecisiveMoveWin(), kw=dict(use_transpositions=True)),
        }
        for name, cfg in searches.items():
            try:
                with mb.GpuSearch(cfg["game"], cfg["strat"], max_iterations=4096, batch_leaves=128, num_rollouts_per_leaf=64,
                                  pipelined=True, seed=7, devices=[local], **cfg["kw"]) as srch:
                    srch.choose_action(mb.initial_state(cfg["game"]))
                    t0 = time.perf_counter()
                    reps = 3
                    for _ in range(reps):
                        sr = srch.choose_action(mb.initial_state(cfg["game"]))
                    dt = (time.perf_counter() - t0) / reps
                per_game[name] = {
                    "playouts_per_move": sr.playouts, "seconds_per_move": dt, "playouts_per_s": sr.playouts / dt, "best_action": sr.best_action,
                    "nodes": sr.nodes, "batch_leaves": 128, "playouts_per_leaf": 64, "pipelined": True,
                }
            except Exception as e:
                per_game[name] = {"error": str(e)}

    # BASELINE.json config 4: Breakthrough 8x8, 2^24 playouts for ONE move, the search root-sharded over the ranks (one
    # searcher per GPU with its own seed and tree, search/parallel.rs:56-115), root-child totals merged with one NCCL
    # all-reduce.  Strong scaling: every rank runs 2^24 / N playouts.
    if not args.headline_only:
        try:
            total, bl, kk = 1 << 24, 256, 1024  # 2^14 / N leaf selections per searcher, 64 / N batches
            iters = max(1, total // (world * kk))
            root = mb.initial_state(mb.Game.BREAKTHROUGH)
            with mb.GpuSearch(mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), max_iterations=iters, batch_leaves=bl,
                              num_rollouts_per_leaf=kk, max_playout_depth=200, pipelined=True, seed=1000 + rank, devices=[local]) as srch:
                mb.choose_action_root_parallel(srch, root, device=dev, draw=3)
                barrier()
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    action, mv, _ms, lres = mb.choose_action_root_parallel(srch, root, device=dev, draw=3)
                barrier()
                dt = (time.perf_counter() - t0) / reps
            if dist is not None:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            per_game["breakthrough_root_parallel_2p24_per_move"] = {
                "playouts_per_move": world * iters * kk, "seconds_per_move": dt, "playouts_per_s": world * iters * kk / dt,
                "merged_root_child_visits": int(mv.sum()), "best_action": int(action),
                "searchers": world, "iterations_per_searcher": iters, "batch_leaves": bl, "playouts_per_leaf": kk,
                "scaling": "strong", "merge": "one all-reduce of the root-child (visits, score) totals per move",
            }
        except Exception as e:
            per_game["breakthrough_root_parallel_2p24_per_move"] = {"error": str(e)}

    peak_ops, clock_mhz = head["int_peak"]
    achieved = head["plies_per_s"] / world * ops_per_ply(HEADLINE)
    line = {
        "metric": METRIC,
        "value": head["playouts_per_s"],
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": workload_config(HEADLINE, n_leaves, k),
        "clocks": head["clocks"],
        "e2e": head["e2e"],
        "gpu_launches": head["launches"],
        "plies_per_s": head["plies_per_s"],
        "mean_plies_per_playout": head["mean_plies"],
        "roofline": {
            "bound": "int32_issue",
            "achieved": achieved / 1e12,
            "peak": peak_ops / 1e12,
            "unit": "Tiop/s",
            "frac": achieved / peak_ops,
            "traffic": NCU_TRAFFIC_BYTES.get(HEADLINE["name"]),
            "traffic_unit": "bytes of DRAM per launch (ncu), for the record: the kernel is not HBM-bound",
            "note": "achieved = plies/s per GPU x ops/ply of the reference's arithmetic (SURVEY 8d model v1: %.0f ops/ply); "
            "peak = LOP3 issue rate measured live on this GPU (mr_measure_int_peak); the path moves <0.01 B/ply so HBM is not the bound"
            % ops_per_ply(HEADLINE),
        },
        "per_game": per_game,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib as o

        o.lib()
        threads = o.host_threads()
        cn, ck = cpu_sample_size(HEADLINE, threads, 12.0)
        pps, plies_s, dt = cpu_sample(HEADLINE, cn, ck, threads, 4242)
        line["cpu_baseline"] = {
            "value": pps, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{cn} leaves x {ck} playouts of the {HEADLINE['name']} workload, {threads} pthreads, {dt:.1f} s",
        }
        if not args.headline_only:
            # config 1 on the CPU: 100 000 UCT iterations (k = 1) from the empty Connect4 board, one thread,
            # the reference's own "rollouts/sec/core" definition (search/core.rs:395-398)
            t0 = time.perf_counter()
            best, _, _ = o.search(o.CONNECT4, o.initial_state(o.CONNECT4), iterations=100000, seed=7)
            dts = time.perf_counter() - t0
            line["cpu_baseline"]["connect4_uct_100k_per_move"] = {
                "rollouts_per_s_per_core": 100000 / dts, "seconds_per_move": dts, "cores": 1, "best_action": best}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--leaves", type=int, default=4096)
    ap.add_argument("--playouts", type=int, default=4096)
    ap.add_argument("--headline-only", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
