"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (mcts_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB_PATH = ORACLE_DIR / "_build" / "liboracle.so"

CONNECT4, BREAKTHROUGH, KNIGHTTHROUGH, OTHELLO, HEX11 = range(5)
GAME_NAMES = ["connect4", "breakthrough", "knightthrough", "othello", "hex11"]
UNIFORM, EPS_MAST, DECISIVE_WIN, DECISIVE_WINLOSS, DECISIVE_WINLOSSDRAW, DECISIVE_ANTI, EPS_LGR, DECISIVE_MAST_WIN, DECISIVE_MAST_WINLOSS, EPS_NST, EPS_LGR2, EPS_LGR2_MAST = range(12)
NOT_TERMINAL, DRAW, WINNER_P0, WINNER_P1 = range(4)
END_TERMINAL, END_TURN_LIMIT, END_NO_ACTIONS = range(3)
OTHELLO_PASS = 64
MAX_ACTIONS = 256

STATE_DTYPE = np.dtype(
    [("board", "<u8", (4,)), ("turn", "u1"), ("flag", "u1"), ("prev", "<u2"), ("prev2", "<u2"), ("pad", "u1", (2,))], align=False
)
RESULT_DTYPE = np.dtype([("wins", "<u4", (2,)), ("draws", "<u4"), ("cutoffs", "<u4"), ("sum_depth", "<u8")])
STAT_DTYPE = np.dtype([("visits", "<u4"), ("score", "<i4")])
RECORD_DTYPE = np.dtype([("depth", "<u4"), ("outcome", "u1"), ("end_type", "u1"), ("pad", "u1", (2,))])
assert STATE_DTYPE.itemsize == 40 and RESULT_DTYPE.itemsize == 24 and RECORD_DTYPE.itemsize == 8


class BatchDesc(C.Structure):
    _fields_ = [
        ("game", C.c_int),
        ("policy", C.c_int),
        ("eps_threshold", C.c_uint64),
        ("max_playout_depth", C.c_uint32),
        ("seed", C.c_uint64),
        ("leaf_id_base", C.c_uint32),
        ("playouts_per_leaf", C.c_uint32),
        ("max_trace", C.c_uint32),
        ("replies", C.c_void_p),
        ("reply_order_out", C.c_void_p),
        ("reply_action_out", C.c_void_p),
        ("last_win_out", C.c_void_p),
        ("bigram", C.c_void_p),
        ("nst_threshold", C.c_uint32),
        ("bigram_delta_out", C.c_void_p),
        ("replies2", C.c_void_p),
        ("replies_final_out", C.c_void_p),
        ("replies2_final_out", C.c_void_p),
    ]


class SearchCfg(C.Structure):
    _fields_ = [
        ("game", C.c_int),
        ("policy", C.c_int),
        ("select", C.c_int),
        ("q_init", C.c_int),
        ("max_iterations", C.c_uint32),
        ("batch_leaves", C.c_uint32),
        ("playouts_per_leaf", C.c_uint32),
        ("expand_threshold", C.c_uint32),
        ("max_playout_depth", C.c_uint32),
        ("threads", C.c_int),
        ("pipeline_depth", C.c_int),
        ("use_transpositions", C.c_int),
        ("exploration_c", C.c_double),
        ("eps_threshold", C.c_uint64),
        ("seed", C.c_uint64),
        ("grave_threshold", C.c_uint32),
        ("rave_schedule", C.c_int),
        ("rave_param", C.c_uint32),
        ("amaf_alpha", C.c_double),
        ("nst_threshold", C.c_uint32),
        ("rave_bias", C.c_double),
    ]


EDGE_STATS_DTYPE = np.dtype([("visits", "<u4"), ("vloss", "<u4"), ("score", "<i8", (2,)), ("sumsq", "<i8", (2,)), ("amaf_visits", "<u4"),
                             ("action", "<u2"), ("created", "<u2"), ("amaf_score", "<i8", (2,))])
assert EDGE_STATS_DTYPE.itemsize == 64
ROOT_CHILD_DTYPE = np.dtype([("action", "<u2"), ("pad", "<u2"), ("visits", "<u4"), ("score", "<i8", (2,))])


def build(force: bool = False) -> Path:
    """Compile the oracle with its own Makefile (gcc).  Idempotent."""
    srcs = list(ORACLE_DIR.glob("*.c")) + list(ORACLE_DIR.glob("*.h")) + [ORACLE_DIR / "Makefile"]
    stale = (not LIB_PATH.exists()) or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.orc_initial_state.argtypes = [i32, vp]
        L.orc_num_actions.argtypes = [i32]
        L.orc_generate_actions.argtypes = [i32, vp, vp]
        L.orc_apply.argtypes = [i32, vp, C.c_uint16]
        L.orc_is_terminal.argtypes = [i32, vp]
        L.orc_winner.argtypes = [i32, vp]
        L.orc_terminal_status.argtypes = [i32, vp]
        L.orc_hex_flood_iterations.restype = u64
        L.orc_c4_generic_random_game.argtypes = [i32, i32, u64, vp, vp]
        L.orc_othello_generate_moves.argtypes = [u64, u64]
        L.orc_othello_generate_moves.restype = u64
        L.orc_othello_get_flips.argtypes = [u64, u64, u64]
        L.orc_othello_get_flips.restype = u64
        L.orc_naive_othello_generate_moves.argtypes = [u64, u64]
        L.orc_naive_othello_generate_moves.restype = u64
        L.orc_naive_othello_get_flips.argtypes = [u64, u64, u64]
        L.orc_naive_othello_get_flips.restype = u64
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_draw.argtypes = [u64, u32, u32, u32, u32]
        L.orc_draw.restype = u32
        L.orc_draw_pair.argtypes = [u64, u32, u32, u32, vp, vp]
        L.orc_playout_batch.argtypes = [C.POINTER(BatchDesc), vp, u32, vp, vp, vp, vp, vp, vp, vp, i32]
        L.orc_random_positions.argtypes = [i32, u64, u32, u32, vp]
        L.orc_stress_othello.argtypes = [u64, vp, vp]
        L.orc_stress_through.argtypes = [i32, u64, vp, vp]
        L.orc_naive_hex_winner.argtypes = [vp, vp, i32]
        L.orc_search.argtypes = [C.POINTER(SearchCfg), vp, vp, vp, vp, vp]
        L.orc_select_score.argtypes = [C.POINTER(SearchCfg), i32, vp, vp, u32, C.c_int64]
        L.orc_select_score.restype = C.c_double
        L.orc_kat_update_amaf.argtypes = [i32, vp, vp, u32, C.c_uint16, vp, vp, u32, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_state(board=(0, 0, 0, 0), turn=0, flag=0) -> np.ndarray:
    s = np.zeros(1, dtype=STATE_DTYPE)
    b = list(board) + [0] * (4 - len(board))
    s["board"][0] = np.array(b, dtype=np.uint64)
    s["turn"] = turn
    s["flag"] = flag
    return s


def initial_state(game: int) -> np.ndarray:
    s = np.zeros(1, dtype=STATE_DTYPE)
    lib().orc_initial_state(game, _p(s))
    return s


def num_actions(game: int) -> int:
    return lib().orc_num_actions(game)


def num_slots(game: int) -> int:
    return lib().orc_num_slots(game)


def action_slot(game: int, player: int, code: int) -> int:
    return lib().orc_action_slot(game, player, code)


def generate_actions(game: int, state: np.ndarray) -> list[int]:
    out = np.zeros(MAX_ACTIONS, dtype=np.uint16)
    n = lib().orc_generate_actions(game, _p(state), _p(out))
    return [int(x) for x in out[:n]]


def apply(game: int, state: np.ndarray, action: int) -> np.ndarray:
    s = state.copy()
    lib().orc_apply(game, _p(s), int(action))
    return s


def is_terminal(game: int, state: np.ndarray) -> bool:
    return bool(lib().orc_is_terminal(game, _p(state)))


def winner(game: int, state: np.ndarray) -> int:
    return lib().orc_winner(game, _p(state))


def terminal_status(game: int, state: np.ndarray) -> int:
    return lib().orc_terminal_status(game, _p(state))


def philox(ctr, key) -> list[int]:
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return [int(x) for x in o]


def draw(seed, leaf_id, playout_id, ply, stream=0) -> int:
    return lib().orc_draw(seed, leaf_id, playout_id, ply, stream)


def draw_pair(seed, leaf_id, playout_id, ply) -> tuple[int, int]:
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_draw_pair(seed, leaf_id, playout_id, ply, C.byref(a), C.byref(b))
    return a.value, b.value


def eps_threshold(eps: float) -> int:
    return max(0, min(1 << 32, int(eps * 4294967296.0)))


def playout_batch(
    game,
    leaves: np.ndarray,
    k: int,
    *,
    policy=UNIFORM,
    eps=0.1,
    max_depth=0,
    seed=0,
    leaf_id_base=0,
    mast: np.ndarray | None = None,
    want_amaf=False,
    want_mast_delta=False,
    want_trace=False,
    max_trace=0,
    want_final=False,
    threads=1,
    replies: np.ndarray | None = None,
    want_reply_updates=False,
    bigram: np.ndarray | None = None,
    nst_threshold=5,
    want_bigram_delta=False,
    replies2: np.ndarray | None = None,
    want_final_tables=False,
    _use_lib=None,
):
    """Runs n_leaves*k playouts.  Returns a dict with results (+ optional amaf / mast_delta / trace /
    records / final)."""
    leaves = np.ascontiguousarray(leaves, dtype=STATE_DTYPE)
    n = len(leaves)
    na = num_actions(game)
    d = BatchDesc(game, policy, eps_threshold(eps), max_depth, seed, leaf_id_base, k, max_trace if want_trace else 0)
    results = np.zeros(n, dtype=RESULT_DTYPE)
    amaf = np.zeros((n, 2, na), dtype=STAT_DTYPE) if want_amaf else None
    delta = np.zeros((2, na), dtype=STAT_DTYPE) if want_mast_delta else None
    trace = np.zeros((n * k, max_trace), dtype=np.uint16) if want_trace else None
    rec = np.zeros(n * k, dtype=RECORD_DTYPE) if (want_trace or want_final) else None
    final = np.zeros(n * k, dtype=STATE_DTYPE) if want_final else None
    reply_order = reply_action = last_win = None
    if replies is not None:
        replies = np.ascontiguousarray(replies, dtype=np.uint16)
        assert replies.shape == (2, na)
        d.replies = replies.ctypes.data
    if want_reply_updates:
        reply_order = np.zeros((2, na), dtype=np.uint64)
        reply_action = np.zeros((2, na), dtype=np.uint16)
        last_win = np.zeros((n, 2), dtype=np.uint32)
        d.reply_order_out, d.reply_action_out, d.last_win_out = reply_order.ctypes.data, reply_action.ctypes.data, last_win.ctypes.data
    bigram_delta = None
    ns = num_slots(game)
    if bigram is not None:
        bigram = np.ascontiguousarray(bigram, dtype=STAT_DTYPE)
        assert bigram.shape == (2, ns, ns)
        d.bigram = bigram.ctypes.data
    d.nst_threshold = nst_threshold
    if want_bigram_delta:
        bigram_delta = np.zeros((2, ns, ns), dtype=STAT_DTYPE)
        d.bigram_delta_out = bigram_delta.ctypes.data
    replies_final = replies2_final = None
    if replies2 is not None:
        replies2 = np.ascontiguousarray(replies2, dtype=np.uint16)
        assert replies2.shape == (2, ns, ns)
        d.replies2 = replies2.ctypes.data
    if want_final_tables:
        replies_final = np.zeros((2, na), dtype=np.uint16)
        replies2_final = np.zeros((2, ns, ns), dtype=np.uint16)
        d.replies_final_out, d.replies2_final_out = replies_final.ctypes.data, replies2_final.ctypes.data
    if mast is not None:
        mast = np.ascontiguousarray(mast, dtype=STAT_DTYPE)
        assert mast.shape == (2, na)
    rc = (_use_lib or lib()).orc_playout_batch(
        C.byref(d), _p(leaves), n, _p(mast), _p(results), _p(amaf), _p(delta), _p(trace), _p(rec), _p(final), threads
    )
    if rc != 0:
        raise RuntimeError(f"orc_playout_batch failed: {rc}")
    return {"results": results, "amaf": amaf, "mast_delta": delta, "trace": trace, "records": rec, "final": final,
            "reply_order": reply_order, "reply_action": reply_action, "last_win": last_win, "bigram_delta": bigram_delta,
            "replies_final": replies_final, "replies2_final": replies2_final}


OP_KINDS = ["logic64", "shift64", "popc64", "ctz64", "list_store", "draw", "score_eval", "ply", "flood_iter"]
# 32-bit integer instructions charged per counted operation (a 64-bit logical / shift operation is two
# 32-bit instructions, a 64-bit population count three, a random word 20 (Philox4x32-10 ~ 80 per four words), one MAST
# score evaluation + comparison 8 (two wide multiplies and a compare per legal move); a list store and the lowest-bit
# search are charged one and three)
OP_WEIGHTS = {"logic64": 2, "shift64": 2, "popc64": 3, "ctz64": 3, "list_store": 1, "draw": 20, "score_eval": 8}
_ops_lib = None


def ops_lib() -> C.CDLL:
    """The counting build of the oracle (liboracle_ops.so, -DORC_COUNT_OPS): same sources, operation counters live."""
    global _ops_lib
    if _ops_lib is None:
        build()
        L = C.CDLL(str(LIB_PATH.with_name("liboracle_ops.so")))
        L.orc_playout_batch.argtypes = lib().orc_playout_batch.argtypes
        L.orc_ops_read.argtypes = [C.c_void_p]
        assert L.orc_ops_enabled() == 1
        _ops_lib = L
    return _ops_lib


def count_ops(game, leaves, k, **kw) -> dict:
    """Operation counts of the oracle (= the reference's arithmetic as its algorithm executes it) over one single-threaded
    batch: raw counts per kind, plies, and `ops_per_ply` = sum(count * OP_WEIGHTS) / plies in 32-bit integer instructions."""
    L = ops_lib()
    L.orc_ops_reset()
    kw["threads"] = 1
    out = playout_batch(game, leaves, k, _use_lib=L, **kw)
    raw = np.zeros(len(OP_KINDS), dtype=np.uint64)
    L.orc_ops_read(_p(raw))
    counts = {nm: int(v) for nm, v in zip(OP_KINDS, raw)}
    plies = counts["ply"]
    assert plies == int(out["results"]["sum_depth"].sum())
    total = sum(counts[nm] * w for nm, w in OP_WEIGHTS.items())
    return {"counts": counts, "plies": plies, "ops_per_ply": total / max(1, plies),
            "legal_per_greedy_ply": None, "flood_iters_per_ply": counts["flood_iter"] / max(1, plies),
            "list_stores_per_ply": counts["list_store"] / max(1, plies)}


def canonical_representation(game, state):
    """(canonical state, symmetry index) by the oracle's square-by-square restatement."""
    st = np.ascontiguousarray(state, dtype=STATE_DTYPE).reshape(-1)[:1]
    out = st.copy()
    sym = lib().orc_canonical_representation(int(game), _p(st), _p(out))
    return out, sym


def search_cfg(game=0, *, iterations=1, batch_leaves=1, k=1, policy=UNIFORM, select=0, q_init=0, expand_threshold=1,
               max_depth=0, c=2.0 ** 0.5, eps=0.1, seed=0, threads=1, use_transpositions=False, grave_threshold=700,
               rave_schedule=0, rave_param=1000, pipelined=False, pipeline_depth=0, amaf_alpha=1.0, nst_threshold=5, rave_bias=0.0, canonical_symmetry=False):
    depth = int(pipeline_depth) if pipeline_depth else (2 if pipelined else 1)
    return SearchCfg(game, policy, select, q_init, iterations, batch_leaves, k, expand_threshold, max_depth, threads,
                     depth, int(bool(use_transpositions)) | (3 if canonical_symmetry else 0), c, eps_threshold(eps), seed, grave_threshold, rave_schedule,
                     rave_param, amaf_alpha, nst_threshold, rave_bias)


def search(game, root, **kw):
    """The restated reference MCTS loop with CPU playouts.  Returns (best_action, children, playout_plies)."""
    cfg = search_cfg(game, **kw)
    root = np.ascontiguousarray(root, dtype=STATE_DTYPE)
    children = np.zeros(256, dtype=ROOT_CHILD_DTYPE)
    best, n, plies = C.c_uint16(), C.c_uint32(), C.c_uint64()
    rc = lib().orc_search(C.byref(cfg), _p(root), C.byref(best), _p(children), C.byref(n), C.byref(plies))
    if rc != 0:
        raise RuntimeError(f"orc_search failed: {rc}")
    return best.value, children[: n.value].copy(), plies.value


def edge_stats(visits=0, vloss=0, score=(0, 0), sumsq=(0, 0), amaf_visits=0, amaf_score=(0, 0)) -> np.ndarray:
    e = np.zeros(1, dtype=EDGE_STATS_DTYPE)
    e["visits"], e["vloss"], e["score"], e["sumsq"], e["amaf_visits"], e["amaf_score"] = visits, vloss, score, sumsq, amaf_visits, amaf_score
    return e


def select_score(cfg: SearchCfg, player: int, current: np.ndarray, child, grave_n=0, grave_q=0) -> float:
    """The oracle's selection formula on hand-built statistics (child None = a child that does not exist yet)."""
    return lib().orc_select_score(C.byref(cfg), player, _p(current), _p(child) if child is not None else None, grave_n, grave_q)


def kat_update_amaf(game, root, created, processed, amaf, utilities, k=1) -> np.ndarray:
    root = np.ascontiguousarray(root, dtype=STATE_DTYPE)
    created = np.ascontiguousarray(created, dtype=np.uint16)
    amaf = np.ascontiguousarray(amaf, dtype=STAT_DTYPE)
    util = np.ascontiguousarray(utilities, dtype=np.int64)
    out = np.zeros(256, dtype=EDGE_STATS_DTYPE)
    n = C.c_uint32()
    rc = lib().orc_kat_update_amaf(game, _p(root), _p(created), len(created), processed, _p(amaf), _p(util), k, _p(out), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"orc_kat_update_amaf failed: {rc}")
    return out[: n.value].copy()


def random_positions(game: int, seed: int, plies: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=STATE_DTYPE)
    lib().orc_random_positions(game, seed, plies, n, _p(out))
    return out


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
