"""Host mirror of the reference's `Search::choose_action` for the GPU-driven tree driver (mr_search).

Mirrors `SearchConfig` knobs of thomasmarsh/mcts (mcts/src/strategies/mcts/config.rs:291-471):
`max_iterations`, `num_rollouts_per_leaf`, `expand_threshold`, `max_playout_depth`, the select strategy's
exploration constant, `q_init`, and the simulate strategy -- plus `batch_leaves`, the one new knob
(leaves gathered per GPU batch).  The tree itself runs inside libmcts_b200.so (csrc/tree_search.cu).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import EDGE_STATS_DTYPE, LEAF_DTYPE, MAX_ROOT_CHILDREN, ROOT_CHILD_DTYPE, SearchDesc, SearchStats, ptr
from .rollout import Game, GpuRollout, Uniform

SELECT_UCB1, SELECT_UCB1_TUNED, SELECT_GRAVE, SELECT_AMAF = 0, 1, 2, 3
RAVE_HAND_SELECTED, RAVE_THRESHOLD, RAVE_MIN_MSE = 0, 1, 2
QINIT_PARENT, QINIT_INFINITY = 0, 1


@dataclass
class SearchResult:
    best_action: int
    children: np.ndarray  # ROOT_CHILD_DTYPE, in root action order
    playouts: int
    playout_plies: int
    nodes: int
    batches: int
    max_tree_depth: int
    seconds_select: float
    seconds_gpu: float
    seconds_backprop: float
    seconds_submit: float = 0.0


class GpuSearch:
    """`choose_action` with the simulation phase on the GPU."""

    def __init__(self, game: Game, strategy=None, *, max_iterations=10_000, batch_leaves=256, num_rollouts_per_leaf=1,
                 expand_threshold=1, max_playout_depth=0, exploration_constant=math.sqrt(2.0), select=SELECT_UCB1,
                 q_init=QINIT_PARENT, use_transpositions=False, canonical_symmetry=False, pipelined=False, pipeline_depth=0, grave_threshold=700,
                 rave_schedule=RAVE_HAND_SELECTED, rave_param=1000, rave_bias=0.0, amaf_alpha=1.0, seed=0, devices=1, num_threads=1):
        """`pipeline_depth` batches in flight (`pipelined=True` = 2).  `num_threads` > 1 = the reference's root-parallel
        search (SearchConfig.num_threads, search/parallel.rs:56-115): that many independent searchers of
        `max_iterations` iterations each, one host thread and one rollout context per searcher on the same device(s),
        root-child totals merged.  `canonical_symmetry` (implies `use_transpositions`): the table is keyed by Othello's
        canonical D4 orientation up to 20 discs, as the reference keys it (MR_SEARCH_SYMMETRY)."""
        self.rollout = GpuRollout(game, strategy or Uniform(), devices=devices, seed=seed, max_playout_depth=max_playout_depth)
        self.extra_rollouts = [GpuRollout(game, strategy or Uniform(), devices=devices, seed=seed, max_playout_depth=max_playout_depth)
                               for _ in range(max(1, int(num_threads)) - 1)]
        self.desc = SearchDesc(
            int(game), int(self.rollout.strategy.policy), select, q_init, max_iterations, batch_leaves,
            num_rollouts_per_leaf, expand_threshold, max_playout_depth, (1 if use_transpositions else 0) | (2 if pipelined else 0) | (4 if canonical_symmetry else 0), exploration_constant,
            float(getattr(self.rollout.strategy, "epsilon", 0.0)), int(seed) & 0xFFFFFFFFFFFFFFFF,
            grave_threshold, rave_schedule, rave_param, int(getattr(self.rollout.strategy, "backoff_threshold", 5)), float(amaf_alpha),
            float(rave_bias), int(pipeline_depth), 0,
        )

    def close(self):
        self.rollout.close()
        for r in self.extra_rollouts:
            r.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def choose_action(self, state: np.ndarray) -> SearchResult:
        root = np.ascontiguousarray(state, dtype=LEAF_DTYPE).reshape(-1)[:1]
        children = np.zeros(MAX_ROOT_CHILDREN, dtype=ROOT_CHILD_DTYPE)
        best, n = C.c_uint16(), C.c_uint32()
        L = _lib.load()
        if self.extra_rollouts:
            ctxs = [self.rollout] + self.extra_rollouts
            handles = (C.c_void_p * len(ctxs))(*[r._h for r in ctxs])
            sts = (SearchStats * len(ctxs))()
            rc = L.mr_search_root_parallel(handles, len(ctxs), C.byref(self.desc), ptr(root), C.byref(best), ptr(children), C.byref(n), sts)
            self.rollout._check(rc)
            return SearchResult(
                best.value, children[: n.value].copy(), sum(s.playouts for s in sts), sum(s.playout_plies for s in sts),
                sum(s.nodes for s in sts), sum(s.batches for s in sts), max(s.max_tree_depth for s in sts),
                max(s.seconds_select for s in sts), max(s.seconds_gpu for s in sts), max(s.seconds_backprop for s in sts),
                max(s.seconds_submit for s in sts))
        st = SearchStats()
        self.rollout._check(L.mr_search(self.rollout._h, C.byref(self.desc), ptr(root), C.byref(best), ptr(children), C.byref(n), C.byref(st)))
        return SearchResult(best.value, children[: n.value].copy(), st.playouts, st.playout_plies, st.nodes, st.batches,
                            st.max_tree_depth, st.seconds_select, st.seconds_gpu, st.seconds_backprop, st.seconds_submit)


def edge_stats(visits=0, vloss=0, score=(0, 0), sumsq=(0, 0), amaf_visits=0, amaf_score=(0, 0)) -> np.ndarray:
    """One NodeStats / ChildArray row for the known-answer hooks."""
    e = np.zeros(1, dtype=EDGE_STATS_DTYPE)
    e["visits"], e["vloss"], e["score"], e["sumsq"], e["amaf_visits"], e["amaf_score"] = visits, vloss, score, sumsq, amaf_visits, amaf_score
    return e


def select_score(desc: SearchDesc, player: int, current: np.ndarray, child, grave_n: int = 0, grave_q: int = 0) -> float:
    """mr_select_score: the tree driver's selection formula on hand-built statistics (host code, no GPU needed)."""
    return _lib.load().mr_select_score(C.byref(desc), int(player), ptr(current), ptr(child) if child is not None else None, int(grave_n), int(grave_q))


def make_search_desc(game=0, *, select=SELECT_UCB1, q_init=QINIT_PARENT, exploration_constant=math.sqrt(2.0), grave_threshold=700,
                     rave_schedule=RAVE_HAND_SELECTED, rave_param=1000, rave_bias=0.0, amaf_alpha=1.0) -> SearchDesc:
    return SearchDesc(int(game), 0, select, q_init, 1, 1, 1, 1, 0, 0, exploration_constant, 0.0, 0, grave_threshold, rave_schedule,
                      rave_param, 5, float(amaf_alpha), float(rave_bias), 0, 0)


def kat_update_amaf(game: Game, root: np.ndarray, created, processed: int, amaf: np.ndarray, utilities, k: int = 1) -> np.ndarray:
    """mr_kat_update_amaf: update_amaf (backprop.rs:565-617) on a hand-built two-level tree (host code, no GPU needed)."""
    from ._lib import STAT_DTYPE

    root = np.ascontiguousarray(root, dtype=LEAF_DTYPE).reshape(-1)[:1]
    created = np.ascontiguousarray(created, dtype=np.uint16)
    amaf = np.ascontiguousarray(amaf, dtype=STAT_DTYPE)
    util = np.ascontiguousarray(utilities, dtype=np.int64)
    out = np.zeros(MAX_ROOT_CHILDREN, dtype=EDGE_STATS_DTYPE)
    n = C.c_uint32()
    rc = _lib.load().mr_kat_update_amaf(int(game), ptr(root), ptr(created), len(created), int(processed), ptr(amaf), ptr(util), int(k), ptr(out), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"mr_kat_update_amaf failed: {rc}")
    return out[: n.value].copy()


def choose_action_root_parallel(search: GpuSearch, state: np.ndarray, *, device=None, group=None, draw: int = 0):
    """`choose_action_root_parallel` (search/parallel.rs:56-115) with one searcher per RANK instead of per thread: every
    rank runs its own tree on its own GPU (the caller gives each rank its own `seed`, as the reference draws one per
    worker), the per-root-child `(visits, score[2])` totals are summed over ranks with ONE all-reduce (NCCL over NVLink on
    GPUs, gloo in the CPU tests), and every rank picks the child with the most merged visits by `random_best`
    (`draw` in [0, 16 n) is the shared tie-break draw).  Returns (action, merged visits, merged scores, local result).

    Children are reported in root action-list order on every rank (generate_actions is deterministic), so the merge is
    an element-wise sum; unexplored children carry zero visits, like the reference's `is_explored` filter."""
    from . import root_merge
    from .rollout import generate_actions

    local = search.choose_action(state)
    # always reduce one slot per legal root action, so that every rank enters the same collective even when its own root
    # was never expanded (no children reported)
    actions = local.children["action"] if len(local.children) else np.array(generate_actions(search.rollout.game, state), dtype=np.uint16)
    visits = np.zeros(len(actions), dtype=np.int64)
    scores = np.zeros((len(actions), 2), dtype=np.int64)
    if len(local.children):
        visits[:] = local.children["visits"]
        scores[:] = local.children["score"]
    if len(actions) == 0:
        return local.best_action, visits, scores, local
    visits, scores = root_merge.merge_root_stats(visits, scores, device=device, group=group)
    if visits.sum() == 0:  # no rank expanded its root: nothing to choose from
        return local.best_action, visits, scores, local
    idx = root_merge.final_action_index(visits, int(draw) % (16 * len(visits)))
    return int(actions[idx]), visits, scores, local


def flat_monte_carlo(rollout: GpuRollout, state: np.ndarray, samples_per_move: int = 100, max_rollout_depth: int = 100,
                     ucb1: float | None = None, r: int = 0):
    """FlatMonteCarloStrategy::choose_action (mcts/src/strategies/flat_mc.rs:58-151) as ONE GPU batch through the C ABI
    (`mr_flat_monte_carlo`): every root child is a leaf with `samples_per_move` uniform playouts.  The reference's rollout
    loop makes at most `max_rollout_depth` moves and never checks the state reached by the last one, which is exactly a
    playout with max_playout_depth = max_rollout_depth - 1 scored as 0 at the cutoff.  Returns (action, wins per root
    action).  `r` is the combined random_best draw in [0, 16 n)."""
    st = np.ascontiguousarray(state, dtype=LEAF_DTYPE).reshape(-1)[:1]
    wins = np.zeros(MAX_ROOT_CHILDREN, dtype=np.uint32)
    best, n = C.c_uint16(), C.c_uint32()
    rc = rollout.lib.mr_flat_monte_carlo(rollout._h, int(rollout.game), ptr(st), int(samples_per_move), int(max_rollout_depth),
                                         0 if ucb1 is None else 1, 0.0 if ucb1 is None else float(ucb1), int(r), rollout.seed,
                                         C.byref(best), ptr(wins), None, C.byref(n))
    if rc == 1:  # MR_ERR_INVALID
        raise ValueError("flat_monte_carlo on a state without actions (the reference panics here)")
    rollout._check(rc)
    return best.value, wins[: n.value].astype(np.int64)
