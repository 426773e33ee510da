"""mcts_b200 -- B200-native rollout engine: a drop-in for the simulation phase of thomasmarsh/mcts.

The compute path is hand-written sm_100a CUDA behind the C ABI in include/mcts_rollout.h
(libmcts_b200.so, built in-tree by mcts_b200.build).  There is no CPU fallback.
"""
from ._lib import LEAF_DTYPE, RECORD_DTYPE, RESULT_DTYPE, STAT_DTYPE, EXPORTS, RolloutError, load
from .search import GpuSearch, SearchResult, choose_action_root_parallel
from .rollout import (
    BatchResult,
    DecisiveMove,
    DecisiveMoveMast,
    DecisiveMoveWin,
    EndType,
    EpsilonGreedyLgr,
    EpsilonGreedyLgr2,
    EpsilonGreedyLgr2Mast,
    EpsilonGreedyMast,
    EpsilonGreedyNst,
    Game,
    GpuRollout,
    MastTable,
    Outcome,
    Policy,
    Trial,
    Uniform,
    apply,
    generate_actions,
    terminal_status,
    canonical_representation,
    initial_state,
    make_leaf,
    num_actions,
)

__all__ = [
    "BatchResult", "DecisiveMove", "DecisiveMoveMast", "DecisiveMoveWin", "EndType", "EpsilonGreedyLgr", "EpsilonGreedyLgr2", "EpsilonGreedyLgr2Mast", "EpsilonGreedyMast", "EpsilonGreedyNst", "Game", "GpuRollout", "MastTable", "Outcome",
    "Policy", "Trial", "Uniform", "initial_state", "make_leaf", "num_actions", "LEAF_DTYPE", "RECORD_DTYPE",
    "RESULT_DTYPE", "STAT_DTYPE", "EXPORTS", "RolloutError", "load", "GpuSearch", "SearchResult", "choose_action_root_parallel", "apply",
    "generate_actions", "terminal_status", "canonical_representation",
]
