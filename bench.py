#!/usr/bin/env python
"""bench.py -- playouts/sec of the MCTS simulation phase on B200 (the BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic leaves: N_LEAVES leaf states x K
playouts each, every playout run to its end on the GPU, per-leaf win/draw/depth statistics and the
MAST table delta reduced in-kernel.  The headline workload is BASELINE.json configs[1]
(Breakthrough 8x8, epsilon-greedy MAST rollouts, batched leaves on one B200); the other games, SURVEY
8(d)'s opening-depth matrix and the search configurations are timed for a few steps each and reported
under "per_game" / "depth_matrix" / "search".

  value  : whole-job playouts/s with the leaves already resident in HBM (device-timed, CUDA events
           on the launching stream, max over ranks)
  e2e    : the same metric through the reference-facing C ABI calls (mr_submit + mr_wait over three of the
           ABI's four slots, as the header describes) with HOST buffers: staging copy, H2D of the leaves,
           the MAST snapshot in the kernel parameters and the results + MAST delta written back to host
           memory inside the timed region
  N > 1  : one process per GPU (torchrun); every rank runs its own batch (weak scaling) and the ranks
           merge the MAST delta with one NCCL all-reduce per step, issued on a second stream so that the
           exchange of step i-1 runs while step i's kernel fills the GPU; after the timed region the
           merge is CHECKED against the gathered per-rank tables and the run fails if they disagree

`--impl reference` times the reference's CPU playout path instead (the plain-C restatement in oracle/,
a fixed number of host threads; the Rust reference itself cannot be built in this image) on the same
config: 4096 leaves x 4096 playouts per step.
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
GAME_NAMES = ["connect4", "breakthrough", "knightthrough", "othello", "hex11"]
POLICY_NAMES = ["uniform", "eps_greedy_mast", "decisive_win"]
CPU_THREADS_CAP = 16  # the CPU arm uses min(16, host cores) pthreads on every box, and says so


def _w(name, game, policy, plies, pos_seed, max_depth=0, eps=0.0, amaf=False):
    return dict(name=name, game=game, policy=policy, plies=plies, pos_seed=pos_seed, max_depth=max_depth, eps=eps, amaf=amaf)


# (game, policy, opening plies, seed) -- SURVEY.md section 8(d)
HEADLINE = _w("breakthrough8x8_eps_mast", 1, 1, 20, 2, 200, 0.1)
OTHERS = [
    _w("connect4_uniform_midgame", 0, 0, 12, 1),
    _w("knightthrough8x8_eps_mast", 2, 1, 20, 2, 200, 0.1),
    # (after 20 random plies a Knightthrough game is almost over: 2.4 plies per playout; 8-ply openings show the per-ply rate)
    _w("knightthrough8x8_eps_mast_opening8", 2, 1, 8, 2, 200, 0.1),
    _w("othello_decisive_win", 3, 2, 30, 4),
    _w("hex11_uniform_amaf", 4, 0, 30, 3, amaf=True),
    _w("breakthrough8x8_uniform", 1, 0, 20, 2),
]
# SURVEY 8(d) input matrix: every game at four opening depths (0 = the initial position)
DEPTHS = {
    "connect4_uniform": (0, 0, 1, [0, 8, 12, 16], 0, 0.0, False),
    "breakthrough8x8_eps_mast": (1, 1, 2, [0, 10, 20, 30], 200, 0.1, False),
    "breakthrough8x8_uniform": (1, 0, 2, [0, 10, 20, 30], 0, 0.0, False),
    "knightthrough8x8_eps_mast": (2, 1, 2, [0, 10, 20, 30], 200, 0.1, False),
    "othello_decisive_win": (3, 2, 4, [0, 10, 30, 50], 0, 0.0, False),
    "hex11_uniform_amaf": (4, 0, 3, [0, 10, 30, 60], 0, 0.0, True),
}
# not part of the bench line: extra workloads tools/profile_one.py can name (ncu captures, A/B timing)
PROFILE_ONLY = [
    _w("hex11_eps_mast", 4, 1, 30, 3, 0, 0.1),
    _w("hex11_decisive_win", 4, 2, 30, 3),
    _w("othello_eps_mast", 3, 1, 30, 4, 0, 0.1),
    _w("connect4_eps_mast", 0, 1, 12, 1, 0, 0.1),
]
MAST_WARMUP_K = 16  # uniform playouts per leaf that populate the MAST table before the timed steps


def depth_matrix_workloads():
    out = []
    for base, (game, policy, seed, depths, max_depth, eps, amaf) in DEPTHS.items():
        for p in depths:
            out.append(_w(f"{base}_p{p}", game, policy, p, seed, max_depth, eps, amaf) | {"base": base})
    return out


def all_workloads():
    return [HEADLINE] + OTHERS + depth_matrix_workloads()


# ---- roofline numerators -----------------------------------------------------------------------------------
# ops per ply of a bit-parallel restatement of the reference's arithmetic (model v1).  Its two
# workload-dependent terms -- MAST score evaluations per ply (8 ops each) and Hex flood6 iterations per ply (116 ops
# each) -- are MEASURED by the instrumented oracle (tools/oracle_op_counts.py -> profiles/oracle_ops_r02.json); the
# constants are the survey's.
MODEL_BASE = {0: 125.0, 1: 160.0, 2: 220.0, 3: 750.0, 4: 65.0}


def model_ops_per_ply(w, counted) -> float:
    v = MODEL_BASE[w["game"]]
    if w["policy"] == 1:
        v += 8.0 * counted["score_evals_per_ply"]
    if w["game"] == 4:
        v += 116.0 * counted["flood_iters_per_ply"]
    return v


def load_json(name):
    p = ROOT / "profiles" / name
    try:
        return json.loads(p.read_text())
    except (OSError, ValueError):
        return None


def ops_entry(w):
    tab = load_json("oracle_ops_r02.json")
    if tab and w["name"] in tab["workloads"]:
        return tab["workloads"][w["name"]]
    # not measured: the survey's own assumptions (initial-position branching, 2.9 flood iterations)
    return {"score_evals_per_ply": {1: 25.7, 2: 47.8}.get(w["game"], 10.0) * (1.0 - w["eps"]), "flood_iters_per_ply": 2.9,
            "ops_per_ply_reference_algorithm": None}


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


def workload_config(w, n_leaves, k):
    return {
        "workload": w["name"],
        "game": GAME_NAMES[w["game"]],
        "policy": POLICY_NAMES[w["policy"]],
        "epsilon": w["eps"],
        "leaves_per_step_per_gpu": n_leaves,
        "playouts_per_leaf": k,
        "max_playout_depth": w["max_depth"],
        "openings": f"{w['plies']} uniformly random plies from the initial position (seed {w['pos_seed']})",
        "l2_flush": "256 MiB memset between timed steps (untimed); inputs are ~0.2 MB",
    }


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference's playout path
# --------------------------------------------------------------------------------------------


def cpu_threads():
    import oracle_lib as o

    return min(CPU_THREADS_CAP, o.host_threads())


def cpu_sample(w, n_leaves, k, threads, seed):
    """Times oracle playouts for workload `w`; returns (playouts/s, plies/s, seconds)."""
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


def cpu_step_shape(w, threads, n_leaves, k, max_step_s):
    """The GPU arm's own batch shape whenever one step of it fits `max_step_s` seconds on this host (it does from ~8
    cores up); otherwise fewer playouts per leaf, and the line's config says so."""
    pps, _, _ = cpu_sample(w, 64, 64, threads, 1)
    if n_leaves * k / pps <= max_step_s:
        return n_leaves, k
    return n_leaves, int(max(32, min(k, pps * max_step_s // n_leaves)))


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import oracle_lib as o

    o.lib()
    threads = cpu_threads()
    w = HEADLINE
    for _ in range(args.warmup):
        cpu_sample(w, 16, 32, threads, 7)
    # the GPU arm's config: 4096 leaves x 4096 playouts per step (~2 s per step on 16 threads)
    budget_s = float(os.environ.get("BENCH_REFERENCE_SECONDS", "240"))  # whole timed run; the CPU test suite shrinks it
    n_leaves, k = cpu_step_shape(w, threads, args.leaves, args.playouts, budget_s / max(1, args.steps))
    times = []
    for s in range(args.steps):
        _, _, dt = cpu_sample(w, n_leaves, k, threads, 1000 + s)
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
            "sample": f"{n_leaves} leaves x {k} playouts per step of the {w['name']} workload, {threads} pthreads "
                      f"(generic-board C port of the reference's CPU path: oracle/; min({CPU_THREADS_CAP}, host cores) threads)",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------


def run_gpu(args):
    import torch

    import mcts_b200 as mb
    from mcts_b200 import root_merge
    from mcts_b200.rollout import num_slots

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

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_leaves, k = args.leaves, args.playouts
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # a real (non-legacy) stream: the kernels, torch's memsets and the timing events all go to it; the exchange of the
    # previous step runs on a second one
    stream = torch.cuda.Stream(device=dev)
    comm_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    self_check = {}

    def time_workload(w, steps, warmup, with_e2e, sampler=None, exchange=False):
        strat = {0: mb.Uniform(), 1: mb.EpsilonGreedyMast(w["eps"]), 2: mb.DecisiveMoveWin()}[w["policy"]]
        g = mb.GpuRollout(mb.Game(w["game"]), strat, devices=[local], seed=12345 + rank, max_playout_depth=w["max_depth"])
        leaves = g.random_positions(w["plies"], n_leaves, seed=w["pos_seed"])
        ns = num_slots(g.game)
        mast = None
        if w["policy"] == 1:
            saved, g.strategy = g.strategy, mb.Uniform()
            warm = g.simulate_many(leaves, MAST_WARMUP_K, want_mast_delta=True, seed=999)
            g.strategy = saved
            mast = warm.mast_delta.copy()
            g.set_mast(mast)
        want_delta = w["policy"] == 1
        exchange = exchange and want_delta and dist is not None
        d_leaves = torch.from_numpy(leaves.view(np.uint8).reshape(-1).copy()).to(dev)
        d_results = torch.zeros(n_leaves * 24, dtype=torch.uint8, device=dev)
        # what the ranks exchange once per step: the batch's MAST delta in SLOT space, [2][num_slots] x (visits, score)
        # = 3 KB for Breakthrough; two buffers, so that the all-reduce of step i-1 can run under step i's kernel
        d_delta = [torch.zeros(2 * ns * 2, dtype=torch.int32, device=dev) for _ in range(2)] if want_delta else None
        d_amaf = torch.zeros(n_leaves * 2 * ns * 2, dtype=torch.int32, device=dev) if w["amaf"] else None
        kernel_done = [torch.cuda.Event(), torch.cuda.Event()]
        comm_done = [torch.cuda.Event(), torch.cuda.Event()]

        def launch(i):
            d_results.zero_()
            if d_delta is not None:
                d_delta[i % 2].zero_()
            if d_amaf is not None:
                d_amaf.zero_()
            g.launch_device(
                d_leaves.data_ptr(), d_results.data_ptr(), n_leaves, k, stream=stream.cuda_stream,
                d_amaf=d_amaf.data_ptr() if d_amaf is not None else 0,
                d_mast_delta=d_delta[i % 2].data_ptr() if d_delta is not None else 0,
                leaf_id_base=root_merge.leaf_id_base(rank, world, i, n_leaves), tables_in_slots=True,
            )

        def exchange_step(j):
            """The path's one exchange step for batch j: merge its MAST delta over NVLink (NCCL), on the second stream."""
            comm_stream.wait_event(kernel_done[j % 2])
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(d_delta[j % 2])
                comm_done[j % 2].record(comm_stream)

        def step(i, first):
            # the exchange of batch i-1 is issued FIRST: its small kernel starts at once and runs while this step's
            # persistent kernel fills the rest of the GPU; the step ends when both are done
            if exchange and i > first:
                exchange_step(i - 1)
            if exchange and i > first + 1:
                stream.wait_event(comm_done[i % 2])  # buffer i % 2 still holds batch i-2 until its all-reduce is done
            launch(i)
            kernel_done[i % 2].record(stream)
            if exchange and i > first:
                stream.wait_event(comm_done[(i - 1) % 2])

        if sampler:
            sampler.start()  # before the warm-up: the timed region of this path is tens of milliseconds
        for i in range(warmup):
            step(i, 0)
        if exchange and warmup > 0:
            exchange_step(warmup - 1)
            stream.wait_event(comm_done[(warmup - 1) % 2])
        barrier()
        evs = []
        launches_before = g.kernel_launches()
        for i in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(warmup + i, warmup)
            e1.record(stream)
            evs.append((e0, e1))
        if exchange:  # the last batch's exchange has nothing to hide under: timed on its own
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            exchange_step(warmup + steps - 1)
            stream.wait_event(comm_done[(warmup + steps - 1) % 2])
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
        total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
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
            "exchange": ("one %d-byte NCCL all-reduce (MAST delta, slot space) per step on a second stream, overlapping the next step's kernel; "
                         "the last one timed on its own" % (2 * ns * 8)) if exchange else None,
        }
        if exchange:
            # correctness record of the merge: one more batch, its LOCAL delta gathered from every rank, the all-reduced
            # table must equal the sum of the gathered ones
            torch.cuda.synchronize()
            launch(warmup + steps + 1000)
            local_delta = d_delta[(warmup + steps + 1000) % 2].clone()
            gathered = torch.empty(world * local_delta.numel(), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gathered, local_delta)
            merged = local_delta.clone()
            dist.all_reduce(merged)
            ok = bool((gathered.view(world, -1).sum(0, dtype=torch.int64) == merged.to(torch.int64)).all().item()) and int(merged.abs().sum().item()) > 0
            self_check["mast_delta_allreduce_equals_sum_of_rank_tables"] = ok
        if with_e2e:
            # the calls a host makes: host buffers in, host results out, through the C ABI, double-buffered over two slots
            kw = dict(stats=mast, want_mast_delta=want_delta, want_amaf=w["amaf"], tables_in_slots=True)

            def e2e_pass(count, base):
                # three slots: while batch i's tables are merged over the ranks (a blocking NCCL round trip whose kernel
                # only gets an SM when batch i+1's persistent kernel drains) batch i+2 is already queued behind i+1
                for j in range(min(2, count)):
                    g.submit(leaves, k, slot=j % 3, leaf_id_base=root_merge.leaf_id_base(rank, world, base + j, n_leaves), **kw)
                merged, pending = None, None
                for i in range(count):
                    if i + 2 < count:
                        g.submit(leaves, k, slot=(i + 2) % 3, leaf_id_base=root_merge.leaf_id_base(rank, world, base + i + 2, n_leaves), **kw)
                    r = g.wait(i % 3)
                    if dist is not None and want_delta:
                        # the exchange of batch i is enqueued now and collected one step later: the table a batch plays
                        # from is a batch stale anyway, and the ranks do not run in lock-step through a blocking collective
                        if pending is not None:
                            merged = pending.result()
                        pending = root_merge.merge_table_i32_async(r.mast_delta, device=dev)
                if pending is not None:
                    merged = pending.result()
                return r, merged

            e2e_pass(3, 5000)
            barrier()
            t0 = time.perf_counter()
            r, merged = e2e_pass(steps, 6000)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            done = r.results["wins"][:, 0].astype(np.int64) + r.results["wins"][:, 1] + r.results["draws"]
            assert (done == k).all(), "e2e batch did not complete every playout"
            h2d = n_leaves * 40 + 1536  # leaves + the kernel parameter block (holds the MAST rank planes)
            d2h = n_leaves * 24 + (2 * ns * 8 if want_delta else 0) + (n_leaves * 2 * ns * 8 if w["amaf"] else 0)
            out["e2e"] = {"value": world * n_leaves * k * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "how": "mr_submit / mr_wait with host buffers, three slots in flight" + ("; MAST delta merged over ranks by one NCCL all-reduce per step, collected one step later" if dist is not None and want_delta else "")}
            if dist is not None and want_delta:
                # the merged table every rank ends with must be the sum of the ranks' own last tables
                mine = torch.from_numpy(np.ascontiguousarray(r.mast_delta).view(np.int32).reshape(-1).copy()).to(dev)
                gathered = torch.empty(world * mine.numel(), dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(gathered, mine)
                want = gathered.view(world, -1).sum(0, dtype=torch.int64).cpu().numpy()
                self_check["e2e_merged_mast_delta_equals_sum_of_rank_tables"] = bool((want == merged.view(np.int32).reshape(-1).astype(np.int64)).all())
        out["int_peak"] = g.measure_int_peak(0)
        g.close()
        return out

    sampler = ClockSampler(local) if rank == 0 else None
    head = time_workload(HEADLINE, args.steps, args.warmup, True, sampler, exchange=True)

    per_game, matrix = {}, {}
    if not args.headline_only:
        few = max(2, min(args.steps, 3))
        for w in OTHERS:
            try:
                r = time_workload(w, few, 3, False)
                oe = ops_entry(w)
                per_game[w["name"]] = {
                    "playouts_per_s": r["playouts_per_s"], "plies_per_s": r["plies_per_s"], "mean_plies": r["mean_plies"],
                    "ms_per_step": r["ms_per_step"],
                    "model_ops_per_ply": model_ops_per_ply(w, oe),
                    "model_int_issue_frac": r["plies_per_s"] * model_ops_per_ply(w, oe) / (world * r["int_peak"][0]),
                    "measured": measured_counters(w["name"], r["plies_per_s"] / world, r["int_peak"]),
                }
            except Exception as e:  # a game that fails must not hide the headline
                per_game[w["name"]] = {"error": str(e)}
        for w in depth_matrix_workloads():
            try:
                r = time_workload(w, 2, 3, False)
                matrix.setdefault(w["base"], {})[str(w["plies"])] = {
                    "playouts_per_s": r["playouts_per_s"], "plies_per_s": r["plies_per_s"], "mean_plies": r["mean_plies"], "ms_per_step": r["ms_per_step"]}
            except Exception as e:
                matrix.setdefault(w["base"], {})[str(w["plies"])] = {"error": str(e)}

    search = {}
    if not args.headline_only:
        from mcts_b200.search import QINIT_INFINITY, SELECT_GRAVE

        def time_search(game, strat, total, batch, kk, depth, threads=1, reps=3, **kw):
            iters = max(1, total // (kk * threads))
            with mb.GpuSearch(game, strat, max_iterations=iters, batch_leaves=batch, num_rollouts_per_leaf=kk, pipeline_depth=depth,
                              seed=7, devices=[local], num_threads=threads, **kw) as srch:
                root = mb.initial_state(game)
                srch.choose_action(root)
                t0 = time.perf_counter()
                for _ in range(reps):
                    sr = srch.choose_action(root)
                dt = (time.perf_counter() - t0) / reps
            return {
                "playouts_per_move": sr.playouts, "seconds_per_move": dt, "playouts_per_s": sr.playouts / dt, "best_action": sr.best_action,
                "batch_leaves": batch, "playouts_per_leaf": kk, "pipeline_depth": depth, "searcher_threads": threads, "nodes": sr.nodes,
                "root_child_visit_share": float(sr.children["visits"].sum()) / max(1, sr.playouts),
                "seconds_select": sr.seconds_select, "seconds_submit": sr.seconds_submit, "seconds_gpu_wait": sr.seconds_gpu,
                "seconds_backprop": sr.seconds_backprop,
            }

        # BASELINE.json config 1: Connect4 UCT (UCB1, c = sqrt 2, expand_threshold 1), uniform rollouts, ~100k playouts for one
        # move from the empty board; config 3: Hex 11x11 GRAVE over the in-kernel AMAF tables; config 4: Othello with
        # decisive-move rollouts and the transposition table on the host -- each at a few (batch, playouts per leaf) shapes
        plans = {
            "connect4_uct_100k_per_move": [(mb.Game.CONNECT4, mb.Uniform(), 100_000, b, kk, 4, 1, {}) for b, kk in [(64, 64), (64, 256), (32, 1024)]],
            "hex11_grave_amaf_search": [(mb.Game.HEX11, mb.Uniform(), 4096 * kk, 128, kk, 4, thr, dict(select=SELECT_GRAVE, q_init=QINIT_INFINITY, grave_threshold=700))
                                        for kk, thr in [(64, 1), (256, 1), (1024, 1), (256, 4)]],
            "othello_decisive_tt_search": [(mb.Game.OTHELLO, mb.DecisiveMoveWin(), 4096 * kk, 128, kk, 4, 1, dict(use_transpositions=True)) for kk in (64, 256)],
            # the same with the table keyed as the reference keys it: canonical D4 orientation up to 20 discs (othello/lib.rs:474-480)
            "othello_decisive_tt_canonical_symmetry_search": [(mb.Game.OTHELLO, mb.DecisiveMoveWin(), 4096 * 256, 128, 256, 4, 1,
                                                               dict(use_transpositions=True, canonical_symmetry=True))],
        }
        for name, shapes in plans.items():
            rows = []
            for game, strat, total, b, kk, depth, thr, kw in shapes:
                try:
                    rows.append(time_search(game, strat, total * thr if name.startswith("hex") and thr > 1 else total, b, kk, depth, thr, **kw))
                except Exception as e:
                    rows.append({"error": str(e), "batch_leaves": b, "playouts_per_leaf": kk})
            search[name] = rows

        # BASELINE.json config 5: Breakthrough 8x8, 2^24 playouts for ONE move, root-sharded (search/parallel.rs:56-115): two
        # searcher threads per GPU, one GPU per rank, every searcher its own seed and tree; root-child totals merged
        # in-library over a rank's searchers and by one NCCL all-reduce over the ranks.  Strong scaling: 2^24 / (ranks x
        # searchers) playouts per searcher.  MAST and uniform rollouts.
        # searchers per rank: two while a searcher's share is large (1 or 2 ranks), one from 4 ranks on, where a share is
        # <= 4 096 descents and a second searcher only adds skew in front of the move's all-reduce (measured both ways:
        # the third entry is the other count)
        thr_main = 2 if world <= 2 else 1
        other = ("breakthrough_root_parallel_2p24_one_searcher_per_rank", 1) if thr_main == 2 else ("breakthrough_root_parallel_2p24_two_searchers_per_rank", 2)
        for label, strat, thr in [("breakthrough_root_parallel_2p24_per_move", mb.EpsilonGreedyMast(0.1), thr_main),
                                  ("breakthrough_root_parallel_2p24_uniform_per_move", mb.Uniform(), thr_main),
                                  (other[0], mb.EpsilonGreedyMast(0.1), other[1])]:
            try:
                # (batch shape from tools/config5_probe.py on one B200: 2 searchers x 128 x 1024 at depth 4 7.2-9 ms per
                # move depending on the box's host, 256 x 1024 9.6-10.5 ms; 3 searchers x 128 x 1024 at depth 2 are faster
                # still but 2^24 is not a multiple of 3.  With 8 ranks a searcher's share is 1 024 descents: the third
                # entry shows one searcher per rank.)
                total, bl, kk = 1 << 24, 128, 1024
                iters = max(1, total // (world * thr * kk))
                root = mb.initial_state(mb.Game.BREAKTHROUGH)
                with mb.GpuSearch(mb.Game.BREAKTHROUGH, strat, max_iterations=iters, batch_leaves=bl, num_rollouts_per_leaf=kk,
                                  max_playout_depth=200, pipeline_depth=4, seed=1000 + rank, devices=[local], num_threads=thr) as srch:
                    mb.choose_action_root_parallel(srch, root, device=dev, draw=3)
                    barrier()
                    t0 = time.perf_counter()
                    reps = 5
                    for _ in range(reps):
                        action, mv, _ms, lres = mb.choose_action_root_parallel(srch, root, device=dev, draw=3)
                    barrier()
                    dt = max_over_ranks((time.perf_counter() - t0) / reps)
                    # where a move's time goes outside the tree driver: the search call alone, timed rank by rank without
                    # the merge (the ranks are not synchronised between these calls)
                    t1 = time.perf_counter()
                    for _ in range(reps):
                        srch.choose_action(root)
                    dt_search = max_over_ranks((time.perf_counter() - t1) / reps)
                    barrier()
                entry = {
                    "playouts_per_move": world * thr * iters * kk, "seconds_per_move": dt, "playouts_per_s": world * thr * iters * kk / dt,
                    "merged_root_child_visits": int(mv.sum()), "root_child_visit_share": float(mv.sum()) / (world * thr * iters * kk),
                    "best_action": int(action), "ranks": world, "searcher_threads_per_rank": thr, "iterations_per_searcher": iters,
                    "batch_leaves": bl, "playouts_per_leaf": kk, "pipeline_depth": 4, "scaling": "strong",
                    "seconds_search_call_only": dt_search,
                    "seconds_select": lres.seconds_select, "seconds_submit": lres.seconds_submit, "seconds_gpu_wait": lres.seconds_gpu,
                    "seconds_backprop": lres.seconds_backprop,
                    "merge": "root-child (visits, score) totals: summed over a rank's searchers in-library, one all-reduce over ranks per move",
                }
                if dist is not None:  # the merged totals must be the sum of the ranks' own
                    mine = torch.from_numpy(lres.children["visits"].astype(np.int64)).to(dev)
                    gathered = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
                    dist.all_gather_into_tensor(gathered, mine)
                    ok = bool((gathered.view(world, -1).sum(0).cpu().numpy() == mv).all())
                    self_check[f"{label}: merged root-child visits == sum of rank totals"] = ok
                search[label] = entry
            except Exception as e:
                search[label] = {"error": str(e)}

    peak_ops, clock_mhz = head["int_peak"]
    oe = ops_entry(HEADLINE)
    model_ops = model_ops_per_ply(HEADLINE, oe)
    achieved = head["plies_per_s"] / world * model_ops
    cnt = load_json("ncu_counters_r02.json") or {}
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
        "exchange": head["exchange"],
        "roofline": {
            "bound": "int32_issue",
            "achieved": achieved / 1e12,
            "peak": peak_ops / 1e12,
            "unit": "Tiop/s",
            "frac": achieved / peak_ops,
            "traffic": (cnt.get(HEADLINE["name"]) or {}).get("dram_bytes"),
            "traffic_unit": ("bytes of DRAM per launch (ncu, one capture; it varies from a few MB to ~100 MB with what the L2 happens to evict); ~0.26 MB of it is "
                             "algorithmic (leaves in, results + MAST delta out), the rest is write-back of the per-lane u16 playout logs (2 B per ply, "
                             "re-written in place after every drain) -- at most ~30 GB/s of the 7.7 TB/s, not a bound"),
            "ops_per_ply_model_v1": model_ops,
            "ops_per_ply_reference_algorithm": oe.get("ops_per_ply_reference_algorithm"),
            "measured": measured_counters(HEADLINE["name"], head["plies_per_s"] / world, head["int_peak"]),
            "note": "achieved = plies/s per GPU x ops/ply of SURVEY 8(d) model v1 (a bit-parallel restatement of the reference's arithmetic; its MAST "
                    "term uses the score evaluations per ply COUNTED by the instrumented oracle, profiles/oracle_ops_r02.json); peak = LOP3 issue rate "
                    "measured live on this GPU (mr_measure_int_peak).  The model is a numerator of convenience, not a ceiling: `measured` holds what the "
                    "kernel really executes (ncu instruction counts of this workload x the live plies/s) against the chip's issue peaks, and "
                    "ops_per_ply_reference_algorithm what the reference's own per-piece algorithm executes.  The path moves <0.01 B/ply: HBM is not the bound.",
        },
        "per_game": per_game,
        "depth_matrix": matrix,
        "search": search,
    }
    if dist is not None:
        line["self_check"] = self_check
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib as o

        o.lib()
        threads = cpu_threads()
        cn, ck = cpu_step_shape(HEADLINE, threads, n_leaves, k, 30.0)
        pps, plies_s, dt = cpu_sample(HEADLINE, cn, ck, threads, 4242)
        line["cpu_baseline"] = {
            "value": pps, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{cn} leaves x {ck} playouts of the {HEADLINE['name']} workload, {threads} pthreads (min({CPU_THREADS_CAP}, host cores)), {dt:.1f} s; "
                      "generic-board C port of the reference's CPU path (oracle/)",
        }
        # ops/ply of the reference's arithmetic, counted live by the instrumented oracle on a small sample of the workload
        leaves = o.random_positions(HEADLINE["game"], HEADLINE["pos_seed"], HEADLINE["plies"], 64)
        warm = o.playout_batch(HEADLINE["game"], leaves, MAST_WARMUP_K, seed=999, max_depth=HEADLINE["max_depth"], want_mast_delta=True)
        oc = o.count_ops(HEADLINE["game"], leaves, 16, policy=HEADLINE["policy"], eps=HEADLINE["eps"], max_depth=HEADLINE["max_depth"], seed=4242, mast=warm["mast_delta"])
        line["cpu_baseline"]["ops_counted"] = {
            "ops_per_ply_reference_algorithm": oc["ops_per_ply"], "score_evals_per_ply": oc["counts"]["score_eval"] / oc["plies"],
            "sample": "64 leaves x 16 playouts, one thread, liboracle_ops.so (-DORC_COUNT_OPS)"}
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
    failed = [kname for kname, ok in self_check.items() if not ok]
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if failed:
        raise SystemExit(f"bench.py: merge self-check FAILED on rank {rank}: {failed}")


def measured_counters(name, plies_per_s_per_gpu, int_peak):
    """What the kernel executes, from the ncu capture of this very workload (profiles/ncu_counters_r02.json: instruction
    counts per ply are a property of the workload and the seed, not of the run) times the LIVE plies/s: the fraction of the
    chip's thread-instruction issue rate (148 SMs x 128 lanes x clock) and of its warp issue slots (148 x 4 x clock) in use."""
    cnt = (load_json("ncu_counters_r02.json") or {}).get(name)
    if not cnt:
        return None
    peak_lop3, clock_mhz = int_peak
    sms = 148
    thread_peak = sms * 128 * clock_mhz * 1e6
    warp_peak = sms * 4 * clock_mhz * 1e6
    return {
        "thread_inst_per_ply": cnt["thread_inst_per_ply"],
        "lanes_per_inst": cnt["lanes_per_inst"],
        "thread_issue_frac": cnt["thread_inst_per_ply"] * plies_per_s_per_gpu / thread_peak,
        "issue_slot_frac": cnt["warp_inst_per_ply"] * plies_per_s_per_gpu / warp_peak,
        "alu_pipe_frac_ncu": cnt["alu_pipe_pct"] / 100.0,
        "fma_pipe_frac_ncu": cnt["fma_pipe_pct"] / 100.0,
        "registers": cnt["registers"],
        "source": cnt["source"],
    }


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
