"""ctypes binding of libmcts_b200.so -- the ONLY compute path of this package.

There is no CPU or PyTorch fallback: if the CUDA library is missing or no sm_100 device is present,
loading / mr_init fails loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

LIB_PATH = _build.LIB_PATH

MR_OK = 0
ERR_NAMES = {1: "MR_ERR_INVALID", 2: "MR_ERR_CUDA", 3: "MR_ERR_NO_DEVICE", 4: "MR_ERR_BUSY", 5: "MR_ERR_UNSUPPORTED", 6: "MR_ERR_COMM", 7: "MR_ERR_NO_MEMORY"}
ABI_VERSION = 2
NUM_SLOTS = 4

LEAF_DTYPE = np.dtype([("board", "<u8", (4,)), ("turn", "u1"), ("flag", "u1"), ("prev", "<u2"), ("prev2", "<u2"), ("pad", "u1", (2,))])
REPLY_UPDATE_DTYPE = np.dtype([("order", "<u8"), ("action_plus1", "<u2"), ("pad", "<u2", (3,))])
RESULT_DTYPE = np.dtype([("wins", "<u4", (2,)), ("draws", "<u4"), ("cutoffs", "<u4"), ("sum_depth", "<u8")])
STAT_DTYPE = np.dtype([("visits", "<u4"), ("score", "<i4")])
RECORD_DTYPE = np.dtype([("depth", "<u4"), ("outcome", "u1"), ("end_type", "u1"), ("pad", "u1", (2,))])
assert LEAF_DTYPE.itemsize == 40 and RESULT_DTYPE.itemsize == 24 and STAT_DTYPE.itemsize == 8 and RECORD_DTYPE.itemsize == 8


class BatchDesc(C.Structure):
    _fields_ = [
        ("game", C.c_uint32),
        ("policy", C.c_uint32),
        ("n_leaves", C.c_uint32),
        ("playouts_per_leaf", C.c_uint32),
        ("max_playout_depth", C.c_uint32),
        ("flags", C.c_uint32),
        ("seed", C.c_uint64),
        ("leaf_id_base", C.c_uint32),
        ("max_trace", C.c_uint32),
        ("epsilon", C.c_double),
    ]


class SearchDesc(C.Structure):
    _fields_ = [
        ("game", C.c_uint32),
        ("policy", C.c_uint32),
        ("select", C.c_uint32),
        ("q_init", C.c_uint32),
        ("max_iterations", C.c_uint32),
        ("batch_leaves", C.c_uint32),
        ("playouts_per_leaf", C.c_uint32),
        ("expand_threshold", C.c_uint32),
        ("max_playout_depth", C.c_uint32),
        ("flags", C.c_uint32),
        ("exploration_c", C.c_double),
        ("epsilon", C.c_double),
        ("seed", C.c_uint64),
        ("grave_threshold", C.c_uint32),
        ("rave_schedule", C.c_uint32),
        ("rave_param", C.c_uint32),
        ("nst_backoff_threshold", C.c_uint32),
        ("amaf_alpha", C.c_double),
        ("rave_bias", C.c_double),
        ("pipeline_depth", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class SearchStats(C.Structure):
    _fields_ = [
        ("playouts", C.c_uint64),
        ("playout_plies", C.c_uint64),
        ("nodes", C.c_uint64),
        ("batches", C.c_uint32),
        ("max_tree_depth", C.c_uint32),
        ("seconds_select", C.c_double),
        ("seconds_gpu", C.c_double),
        ("seconds_backprop", C.c_double),
        ("seconds_submit", C.c_double),
    ]


ROOT_CHILD_DTYPE = np.dtype([("action", "<u2"), ("pad", "<u2"), ("visits", "<u4"), ("score", "<i8", (2,))])
assert ROOT_CHILD_DTYPE.itemsize == 24
EDGE_STATS_DTYPE = np.dtype([("visits", "<u4"), ("vloss", "<u4"), ("score", "<i8", (2,)), ("sumsq", "<i8", (2,)), ("amaf_visits", "<u4"),
                             ("action", "<u2"), ("created", "<u2"), ("amaf_score", "<i8", (2,))])
assert EDGE_STATS_DTYPE.itemsize == 64
MAX_ROOT_CHILDREN = 256


class RolloutError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


# every symbol include/mcts_rollout.h declares
EXPORTS = [
    "mr_abi_version",
    "mr_num_actions",
    "mr_init",
    "mr_destroy",
    "mr_last_error",
    "mr_num_devices",
    "mr_submit",
    "mr_wait",
    "mr_playout_batch",
    "mr_launch_device",
    "mr_set_mast",
    "mr_set_replies",
    "mr_reply_updates",
    "mr_set_replies2",
    "mr_reply2_inserts",
    "mr_reply2_resolve",
    "mr_num_slots",
    "mr_action_slot",
    "mr_slot_action",
    "mr_set_bigrams",
    "mr_bigram_delta",
    "mr_last_kernel_ms",
    "mr_kernel_launches",
    "mr_set_kernel_timing",
    "mr_allreduce_root",
    "mr_comm_unique_id",
    "mr_comm_init",
    "mr_comm_allreduce_i32",
    "mr_measure_int_peak",
    "mr_search",
    "mr_search_root_parallel",
    "mr_select_score",
    "mr_kat_update_amaf",
    "mr_generate_actions",
    "mr_apply",
    "mr_terminal_status",
    "mr_canonical_representation",
    "mr_flat_monte_carlo",
]

_lib = None


def load() -> C.CDLL:
    """Loads the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not Path(LIB_PATH).exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mcts_b200.build` "
            "(nvcc, sm_100a). mcts_b200 has no CPU fallback."
        )
    L = C.CDLL(str(LIB_PATH))
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    L.mr_abi_version.restype = i32
    L.mr_num_actions.argtypes = [i32]
    L.mr_init.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    L.mr_destroy.argtypes = [vp]
    L.mr_destroy.restype = None
    L.mr_last_error.argtypes = [vp]
    L.mr_last_error.restype = C.c_char_p
    L.mr_num_devices.argtypes = [vp]
    L.mr_submit.argtypes = [vp, i32, C.POINTER(BatchDesc), vp, vp]
    L.mr_wait.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.mr_playout_batch.argtypes = [vp, C.POINTER(BatchDesc), vp, vp, vp, vp, vp, vp, vp, vp]
    L.mr_launch_device.argtypes = [vp, i32, C.POINTER(BatchDesc), vp, vp, vp, vp, vp]
    L.mr_set_mast.argtypes = [vp, i32, vp]
    L.mr_set_replies.argtypes = [vp, i32, vp]
    L.mr_reply_updates.argtypes = [vp, i32, vp, vp]
    L.mr_set_replies2.argtypes = [vp, i32, vp]
    L.mr_reply2_inserts.argtypes = [vp, i32, vp, vp]
    L.mr_reply2_resolve.argtypes = [vp, i32, vp, vp]
    L.mr_num_slots.argtypes = [i32]
    L.mr_action_slot.argtypes = [i32, i32, C.c_uint16]
    L.mr_slot_action.argtypes = [i32, i32, i32]
    L.mr_set_bigrams.argtypes = [vp, i32, vp, u32]
    L.mr_bigram_delta.argtypes = [vp, i32, vp]
    L.mr_last_kernel_ms.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.mr_kernel_launches.argtypes = [vp]
    L.mr_kernel_launches.restype = C.c_uint64
    L.mr_allreduce_root.argtypes = [vp, C.POINTER(vp), i32, C.c_size_t]
    L.mr_comm_unique_id.argtypes = [vp]
    L.mr_comm_init.argtypes = [vp, vp, i32, i32]
    L.mr_comm_allreduce_i32.argtypes = [vp, vp, C.c_size_t, vp]
    L.mr_measure_int_peak.argtypes = [vp, i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mr_generate_actions.argtypes = [i32, vp, vp]
    L.mr_apply.argtypes = [i32, vp, C.c_uint16]
    L.mr_terminal_status.argtypes = [i32, vp]
    L.mr_canonical_representation.argtypes = [i32, vp, vp]
    L.mr_flat_monte_carlo.argtypes = [vp, i32, vp, C.c_uint32, C.c_uint32, i32, C.c_double, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
    L.mr_search.argtypes = [vp, C.POINTER(SearchDesc), vp, C.POINTER(C.c_uint16), vp, C.POINTER(u32), C.POINTER(SearchStats)]
    L.mr_set_kernel_timing.argtypes = [vp, i32]
    L.mr_search_root_parallel.argtypes = [C.POINTER(vp), i32, C.POINTER(SearchDesc), vp, C.POINTER(C.c_uint16), vp, C.POINTER(u32), C.POINTER(SearchStats)]
    L.mr_select_score.argtypes = [C.POINTER(SearchDesc), i32, vp, vp, u32, C.c_int64]
    L.mr_select_score.restype = C.c_double
    L.mr_kat_update_amaf.argtypes = [i32, vp, vp, u32, C.c_uint16, vp, vp, u32, vp, C.POINTER(u32)]
    _lib = L
    return L


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
