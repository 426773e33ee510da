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
from ._lib import LEAF_DTYPE, MAX_ROOT_CHILDREN, ROOT_CHILD_DTYPE, SearchDesc, SearchStats, ptr
from .rollout import Game, GpuRollout, Uniform

SELECT_UCB1, SELECT_UCB1_TUNED = 0, 1
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


class GpuSearch:
    """`choose_action` with the simulation phase on the GPU."""

    def __init__(self, game: Game, strategy=None, *, max_iterations=10_000, batch_leaves=256, num_rollouts_per_leaf=1,
                 expand_threshold=1, max_playout_depth=0, exploration_constant=math.sqrt(2.0), select=SELECT_UCB1,
                 q_init=QINIT_PARENT, seed=0, devices=1):
        self.rollout = GpuRollout(game, strategy or Uniform(), devices=devices, seed=seed, max_playout_depth=max_playout_depth)
        self.desc = SearchDesc(
            int(game), int(self.rollout.strategy.policy), select, q_init, max_iterations, batch_leaves,
            num_rollouts_per_leaf, expand_threshold, max_playout_depth, 0, exploration_constant,
            float(getattr(self.rollout.strategy, "epsilon", 0.0)), int(seed) & 0xFFFFFFFFFFFFFFFF,
        )

    def close(self):
        self.rollout.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def choose_action(self, state: np.ndarray) -> SearchResult:
        root = np.ascontiguousarray(state, dtype=LEAF_DTYPE).reshape(-1)[:1]
        children = np.zeros(MAX_ROOT_CHILDREN, dtype=ROOT_CHILD_DTYPE)
        best, n, st = C.c_uint16(), C.c_uint32(), SearchStats()
        L = _lib.load()
        self.rollout._check(L.mr_search(self.rollout._h, C.byref(self.desc), ptr(root), C.byref(best), ptr(children), C.byref(n), C.byref(st)))
        return SearchResult(best.value, children[: n.value].copy(), st.playouts, st.playout_plies, st.nodes, st.batches,
                            st.max_tree_depth, st.seconds_select, st.seconds_gpu, st.seconds_backprop)
