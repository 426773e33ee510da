"""Host-side mirror of the reference's simulation plug-in surface, over the C ABI.

Reference interface mirrored (thomasmarsh/mcts, paths relative to the reference root):
  * `SimulateStrategy<G>::playout(state, max_playout_depth, &TreeStats, prev_action, rng) -> Trial<G>`
    and `backprop_flags()`                       mcts/src/strategies/mcts/simulate.rs:48-155
  * the strategies in scope: `Uniform` (:157-216), `EpsilonGreedy<G, Mast>` (:485-570, :819-848),
    `DecisiveMove<G, Uniform>` Win mode (:573-760)
  * `Trial { actions, state, status.end_type, depth, terminal }`            simulate.rs:15-46
  * `TreeSearch::simulate_many` (k playouts of one leaf)                    search/core.rs:218-258
  * `TreeStats.player_actions` (the MAST table) and its writer             search/shared.rs:85-94, backprop.rs:857-870

The reference is Rust and no Rust toolchain exists in this image, so the host side above the C ABI is
written in Python (and C++ for the tree driver); INTEGRATION.md shows the Rust binding a maintainer
would add.  Everything here runs on the GPU through libmcts_b200.so -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum

import numpy as np

from . import _lib
from ._lib import LEAF_DTYPE, RECORD_DTYPE, REPLY_UPDATE_DTYPE, RESULT_DTYPE, STAT_DTYPE, BatchDesc, RolloutError, ptr


class Game(IntEnum):
    CONNECT4 = 0
    BREAKTHROUGH = 1
    KNIGHTTHROUGH = 2
    OTHELLO = 3
    HEX11 = 4


class Policy(IntEnum):
    UNIFORM = 0
    EPS_MAST = 1
    DECISIVE_WIN = 2
    DECISIVE_WINLOSS = 3
    DECISIVE_WINLOSSDRAW = 4
    DECISIVE_ANTI = 5
    EPS_LGR = 6
    DECISIVE_MAST_WIN = 7
    DECISIVE_MAST_WINLOSS = 8
    EPS_NST = 9
    EPS_LGR2 = 10
    EPS_LGR2_MAST = 11


class Outcome(IntEnum):  # TerminalStatus (mcts/src/game.rs:117-122)
    NOT_TERMINAL = 0
    DRAW = 1
    WINNER_P0 = 2
    WINNER_P1 = 3


class EndType(IntEnum):  # simulate.rs:15-20 (+ the no-actions ending, :112-116)
    NATURAL_END = 0
    TURN_LIMIT = 1
    NO_ACTIONS = 2


WANT_AMAF, WANT_MAST_DELTA, WANT_TRACE, WANT_FINAL_STATE, TABLES_IN_SLOTS = 1, 2, 4, 8, 16

# BackpropFlags bits (mcts/src/strategies/mcts/config.rs:11-30)
BACKPROP_GRAVE, BACKPROP_GLOBAL, BACKPROP_AMAF, BACKPROP_NST, BACKPROP_LGR, BACKPROP_LGR2 = 1, 2, 4, 8, 16, 32

OTHELLO_PASS = 64
_INITIAL = {
    Game.CONNECT4: ((0, 0, 0, 0), 0, 0),
    Game.BREAKTHROUGH: ((0xFFFF000000000000, 0xFFFF, 0, 0), 0, 0),  # breakthrough/lib.rs:64-82
    Game.KNIGHTTHROUGH: ((0xFFFF000000000000, 0xFFFF, 0, 0), 0, 0),
    Game.OTHELLO: (((1 << 28) | (1 << 35), (1 << 27) | (1 << 36), 0, 0), 0, 0),  # othello/lib.rs:10-11
    Game.HEX11: ((0, 0, 0, 0), 0, 0),
}


def make_leaf(board, turn=0, flag=0) -> np.ndarray:
    """One 40-byte leaf record (include/mcts_rollout.h mr_leaf)."""
    s = np.zeros(1, dtype=LEAF_DTYPE)
    b = list(board) + [0] * (4 - len(board))
    s["board"][0] = np.array(b, dtype=np.uint64)
    s["turn"] = turn
    s["flag"] = flag
    return s


def initial_state(game: Game) -> np.ndarray:
    board, turn, flag = _INITIAL[Game(game)]
    return make_leaf(board, turn, flag)


def generate_actions(game: Game, state: np.ndarray) -> list[int]:
    """Game::generate_actions on a leaf record (host rules of the tree driver)."""
    out = np.zeros(256, dtype=np.uint16)
    st = np.ascontiguousarray(state, dtype=LEAF_DTYPE)
    n = _lib.load().mr_generate_actions(int(game), ptr(st), ptr(out))
    return [int(a) for a in out[:n]]


def apply(game: Game, state: np.ndarray, action: int) -> np.ndarray:
    """Game::apply on a leaf record."""
    st = np.ascontiguousarray(state, dtype=LEAF_DTYPE).copy()
    _lib.load().mr_apply(int(game), ptr(st), int(action))
    return st


def terminal_status(game: Game, state: np.ndarray) -> Outcome:
    st = np.ascontiguousarray(state, dtype=LEAF_DTYPE)
    return Outcome(_lib.load().mr_terminal_status(int(game), ptr(st)))


def canonical_representation(game: Game, state: np.ndarray):
    """Game::canonical_representation on a leaf record: (state in its canonical orientation, symmetry index 0..7).  Othello up
    to 20 discs (games/othello/src/lib.rs:618-646); the identity for everything else."""
    st = np.ascontiguousarray(state, dtype=LEAF_DTYPE)
    out = st.copy()
    sym = _lib.load().mr_canonical_representation(int(game), ptr(st), ptr(out))
    if sym < 0:
        raise ValueError("mr_canonical_representation: bad arguments")
    return out, sym


def num_actions(game: Game) -> int:
    return _lib.load().mr_num_actions(int(game))


def num_slots(game: Game) -> int:
    """Size of the compact per-player action index that keys the bigram (NST) tables."""
    return _lib.load().mr_num_slots(int(game))


def action_slot(game: Game, player: int, code: int) -> int:
    return _lib.load().mr_action_slot(int(game), int(player), int(code))


def slot_action(game: Game, player: int, slot: int) -> int:
    return _lib.load().mr_slot_action(int(game), int(player), int(slot))


def slot_action_ids(game: Game) -> np.ndarray:
    """[2][num_slots] dense action id (index into [2][num_actions] tables) of every compact slot, -1 for a slot that
    names no action of that player (a move off the board)."""
    g, ns = Game(game), num_slots(game)
    out = np.full((2, ns), -1, dtype=np.int64)
    through = g in (Game.BREAKTHROUGH, Game.KNIGHTTHROUGH)
    for pl in range(2):
        for sl in range(ns):
            code = slot_action(g, pl, sl)
            if code >= 0:
                out[pl, sl] = (code >> 8) * 64 + (code & 0xFF) if through else code
    return out


def slots_to_dense(game: Game, table: np.ndarray) -> np.ndarray:
    """Expands a [.., 2, num_slots] STAT_DTYPE table (MR_TABLES_IN_SLOTS) to the dense [.., 2, num_actions] form."""
    ids = slot_action_ids(game)
    na = _lib.load().mr_num_actions(int(game))
    out = np.zeros(table.shape[:-1] + (na,), dtype=table.dtype)
    for pl in range(2):
        ok = ids[pl] >= 0
        out[..., pl, ids[pl][ok]] = table[..., pl, ok]
    return out


# ---- the strategy objects: same names and knobs as the reference ----


@dataclass(frozen=True)
class Uniform:
    """simulate.rs:157-216"""

    policy = Policy.UNIFORM

    def backprop_flags(self) -> int:
        return 0


@dataclass(frozen=True)
class EpsilonGreedyMast:
    """EpsilonGreedy<G, Mast> (simulate.rs:485-570 wrapping :819-848); epsilon default 0.1 (:525)."""

    epsilon: float = 0.1
    policy = Policy.EPS_MAST

    def backprop_flags(self) -> int:
        return BACKPROP_GLOBAL  # Mast::backprop_flags (simulate.rs:826-828)


@dataclass(frozen=True)
class DecisiveMoveWin:
    """DecisiveMove<G, Uniform> with DecisiveMoveMode::Win (simulate.rs:573-760)."""

    policy = Policy.DECISIVE_WIN

    def backprop_flags(self) -> int:
        return 0


@dataclass(frozen=True)
class DecisiveMove:
    """DecisiveMove<G, Uniform> with an explicit mode: "win", "win_loss", "win_loss_draw" or "anti_decisive"
    (DecisiveMoveMode, simulate.rs:586-604)."""

    mode: str = "win"

    @property
    def policy(self) -> Policy:
        return {"win": Policy.DECISIVE_WIN, "win_loss": Policy.DECISIVE_WINLOSS, "win_loss_draw": Policy.DECISIVE_WINLOSSDRAW,
                "anti_decisive": Policy.DECISIVE_ANTI}[self.mode]

    def backprop_flags(self) -> int:
        return 0


@dataclass(frozen=True)
class EpsilonGreedyLgr:
    """EpsilonGreedy<G, Lgr<G, Uniform>> (simulate.rs:485-570, 936-1033; strategy.rs:130-143 `Ucb1Lgr`)."""

    epsilon: float = 0.1
    policy = Policy.EPS_LGR

    def backprop_flags(self) -> int:
        return BACKPROP_LGR  # Lgr::backprop_flags (simulate.rs:984-986)


@dataclass(frozen=True)
class EpsilonGreedyLgr2:
    """EpsilonGreedy<G, Lgr2<G, Lgr<G, Uniform>>> (simulate.rs:485-570, 1036-1142; strategy.rs:143-156 `Ucb1Lgr2`, LGRF-2)."""

    epsilon: float = 0.1
    policy = Policy.EPS_LGR2

    def backprop_flags(self) -> int:
        return BACKPROP_LGR2 | BACKPROP_LGR  # Lgr2::backprop_flags (simulate.rs:1101-1103)


@dataclass(frozen=True)
class EpsilonGreedyLgr2Mast:
    """EpsilonGreedy<G, Lgr2<G, Lgr<G, Mast>>> (strategy.rs:158-171 `Ucb1Lgr2Mast`): LGRF-2 -> LGR-1 -> greedy MAST."""

    epsilon: float = 0.1
    policy = Policy.EPS_LGR2_MAST

    def backprop_flags(self) -> int:
        return BACKPROP_LGR2 | BACKPROP_LGR | BACKPROP_GLOBAL


@dataclass(frozen=True)
class EpsilonGreedyNst:
    """EpsilonGreedy<G, Nst> (simulate.rs:485-570, 852-930; strategy.rs:113-126 `Ucb1Nst`)."""

    epsilon: float = 0.1
    backoff_threshold: int = 5  # Nst::default (simulate.rs:875-881)
    policy = Policy.EPS_NST

    def backprop_flags(self) -> int:
        return BACKPROP_GLOBAL | BACKPROP_NST  # Nst::backprop_flags (simulate.rs:894-896)


@dataclass(frozen=True)
class DecisiveMoveMast:
    """DecisiveMove<G, EpsilonGreedy<G, Mast>> (mcts-tune SimulateSpec::DecisiveMoveMast): mode "win" is the
    `ucb1_tuned_dm_mast` family, "win_loss" the `rave` family (family_catalog.rs:470-514)."""

    epsilon: float = 0.1
    mode: str = "win"

    @property
    def policy(self) -> Policy:
        return {"win": Policy.DECISIVE_MAST_WIN, "win_loss": Policy.DECISIVE_MAST_WINLOSS}[self.mode]

    def backprop_flags(self) -> int:
        return BACKPROP_GLOBAL


class MastTable:
    """TreeStats.player_actions as a dense [2][num_actions] (visits, score) table
    (search/shared.rs:88, node.rs:51-55).  `update` is the GLOBAL write of backprop.rs:857-870 applied
    to a batch's aggregated delta."""

    def __init__(self, game: Game):
        self.game = Game(game)
        self.table = np.zeros((2, num_actions(game)), dtype=STAT_DTYPE)

    def update(self, delta: np.ndarray) -> None:
        self.table["visits"] += delta["visits"]
        self.table["score"] += delta["score"]


@dataclass
class Trial:
    """simulate.rs:27-46"""

    actions: list  # [(action code, player index)]
    state: np.ndarray  # final leaf record
    end_type: EndType
    depth: int
    terminal: Outcome  # NOT_TERMINAL for TurnLimit / no-actions endings (scored as a draw)


@dataclass
class BatchResult:
    results: np.ndarray  # RESULT_DTYPE [n_leaves]
    amaf: np.ndarray | None = None  # STAT_DTYPE [n_leaves][2][NA]
    mast_delta: np.ndarray | None = None  # STAT_DTYPE [2][NA]
    trace: np.ndarray | None = None  # u16 [n_leaves*k][max_trace]
    records: np.ndarray | None = None  # RECORD_DTYPE [n_leaves*k]
    final: np.ndarray | None = None  # LEAF_DTYPE [n_leaves*k]
    kernel_ms: float = 0.0
    reply_updates: np.ndarray | None = None  # REPLY_UPDATE_DTYPE [2][NA]  (Policy.EPS_LGR)
    last_win: np.ndarray | None = None  # u32 [n_leaves][2]                (Policy.EPS_LGR)
    bigram_delta: np.ndarray | None = None  # STAT_DTYPE [2][S][S]           (Policy.EPS_NST)
    reply2_inserts: np.ndarray | None = None  # REPLY_UPDATE_DTYPE [2][S][S]   (Policy.EPS_LGR2)
    last_lose: np.ndarray | None = None  # u32 [n_leaves][2]                (Policy.EPS_LGR2)

    def utilities_sum(self) -> np.ndarray:
        """Per leaf, per player: sum of utilities = wins[p] - losses[p] (node.rs:673-681 aggregated)."""
        r = self.results
        k = r["wins"][:, 0].astype(np.int64) + r["wins"][:, 1] + r["draws"]
        w = r["wins"].astype(np.int64)
        return w - (k[:, None] - r["draws"].astype(np.int64)[:, None] - w)


@dataclass
class _Pending:
    desc: BatchDesc
    keep: list = field(default_factory=list)


class GpuRollout:
    """Drop-in for the CPU playout path: `SimulateStrategy` for one game, on one or more B200s.

    `simulate_many` replaces TreeSearch::simulate_many for a whole batch of leaves;
    `playout` is the single-trial form with the reference's argument meaning.
    """

    def __init__(self, game: Game, strategy=None, *, devices: list[int] | int = 1, seed: int = 0, max_playout_depth: int = 0):
        self.lib = _lib.load()
        self.game = Game(game)
        self.strategy = strategy if strategy is not None else Uniform()
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.max_playout_depth = int(max_playout_depth)  # 0 = unlimited (config.rs:440 default usize::MAX)
        if isinstance(devices, int):
            ids, n = None, devices
        else:
            arr = (C.c_int * len(devices))(*devices)
            ids, n = arr, len(devices)
        h = C.c_void_p()
        rc = self.lib.mr_init(ids, n, C.byref(h))
        if rc != 0:
            raise RolloutError(rc, self.lib.mr_last_error(None).decode())
        self._h = h
        self._pending: dict[int, _Pending] = {}
        self.na = self.lib.mr_num_actions(int(self.game))

    # -- lifecycle --
    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.mr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RolloutError(rc, self.lib.mr_last_error(self._h).decode())

    # -- reference-shaped interface --
    def backprop_flags(self) -> int:
        return self.strategy.backprop_flags()

    def _desc(self, n_leaves, k, flags, seed, leaf_id_base, max_trace, max_depth) -> BatchDesc:
        return BatchDesc(
            int(self.game),
            int(self.strategy.policy),
            n_leaves,
            k,
            self.max_playout_depth if max_depth is None else int(max_depth),
            flags,
            self.seed if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF,
            leaf_id_base,
            max_trace,
            float(getattr(self.strategy, "epsilon", 0.0)),
        )

    def submit(
        self,
        leaves: np.ndarray,
        k: int,
        *,
        slot: int = 0,
        stats: MastTable | np.ndarray | None = None,
        want_amaf=False,
        want_mast_delta=False,
        want_trace=False,
        want_final=False,
        max_trace: int = 0,
        seed: int | None = None,
        leaf_id_base: int = 0,
        max_playout_depth: int | None = None,
        tables_in_slots: bool = False,
    ) -> None:
        """`tables_in_slots`: the AMAF / MAST-delta tables come back indexed by compact per-player action slot,
        [.., 2, num_slots(game)], instead of by dense action id (MR_TABLES_IN_SLOTS); `slot_action_ids` converts."""
        leaves = np.ascontiguousarray(leaves, dtype=LEAF_DTYPE).reshape(-1)
        flags = (
            (WANT_AMAF if want_amaf else 0)
            | (WANT_MAST_DELTA if want_mast_delta else 0)
            | (WANT_TRACE if want_trace else 0)
            | (WANT_FINAL_STATE if want_final else 0)
            | (TABLES_IN_SLOTS if tables_in_slots else 0)
        )
        d = self._desc(len(leaves), int(k), flags, seed, leaf_id_base, max_trace if want_trace else 0, max_playout_depth)
        mast = None
        if stats is not None:
            mast = np.ascontiguousarray(stats.table if isinstance(stats, MastTable) else stats, dtype=STAT_DTYPE)
            if mast.shape != (2, self.na):
                raise ValueError(f"MAST table must have shape (2, {self.na})")
        self._check(self.lib.mr_submit(self._h, slot, C.byref(d), ptr(leaves), ptr(mast)))
        self._pending[slot] = _Pending(d, [leaves, mast])

    def wait(self, slot: int = 0) -> BatchResult:
        pend = self._pending.pop(slot, None)
        if pend is None:
            raise RuntimeError(f"slot {slot} has no batch in flight")
        d = pend.desc
        n, k = d.n_leaves, d.playouts_per_leaf
        res = np.zeros(n, dtype=RESULT_DTYPE)
        tl = self.lib.mr_num_slots(int(self.game)) if d.flags & TABLES_IN_SLOTS else self.na
        amaf = np.zeros((n, 2, tl), dtype=STAT_DTYPE) if d.flags & WANT_AMAF else None
        delta = np.zeros((2, tl), dtype=STAT_DTYPE) if d.flags & WANT_MAST_DELTA else None
        trace = np.zeros((n * k, d.max_trace), dtype=np.uint16) if d.flags & WANT_TRACE else None
        rec = np.zeros(n * k, dtype=RECORD_DTYPE) if d.flags & WANT_TRACE else None
        final = np.zeros(n * k, dtype=LEAF_DTYPE) if d.flags & WANT_FINAL_STATE else None
        self._check(self.lib.mr_wait(self._h, slot, ptr(res), ptr(amaf), ptr(delta), ptr(trace), ptr(rec), ptr(final)))
        ms = C.c_float()
        self.lib.mr_last_kernel_ms(self._h, slot, C.byref(ms))
        upd = last = None
        if d.policy in (int(Policy.EPS_LGR), int(Policy.EPS_LGR2), int(Policy.EPS_LGR2_MAST)):
            upd = np.zeros((2, self.na), dtype=REPLY_UPDATE_DTYPE)
            last = np.zeros((n, 2), dtype=np.uint32)
            self._check(self.lib.mr_reply_updates(self._h, slot, ptr(upd), ptr(last)))
        bigram = None
        if d.policy == int(Policy.EPS_NST):
            ns = self.lib.mr_num_slots(int(self.game))
            bigram = np.zeros((2, ns, ns), dtype=STAT_DTYPE)
            self._check(self.lib.mr_bigram_delta(self._h, slot, ptr(bigram)))
        ins2 = lose = None
        if d.policy in (int(Policy.EPS_LGR2), int(Policy.EPS_LGR2_MAST)):
            ns = self.lib.mr_num_slots(int(self.game))
            ins2 = np.zeros((2, ns, ns), dtype=REPLY_UPDATE_DTYPE)
            lose = np.zeros((n, 2), dtype=np.uint32)
            self._check(self.lib.mr_reply2_inserts(self._h, slot, ptr(ins2), ptr(lose)))
        return BatchResult(res, amaf, delta, trace, rec, final, ms.value, upd, last, bigram, ins2, lose)

    def simulate_many(self, leaves: np.ndarray, k: int, **kw) -> BatchResult:
        """k playouts of every leaf (search/core.rs:218-258 for a batch of leaves)."""
        self.submit(leaves, k, slot=0, **kw)
        return self.wait(0)

    def playout(
        self,
        state: np.ndarray,
        max_playout_depth: int | None = None,
        stats: MastTable | np.ndarray | None = None,
        prev_action=None,
        rng: tuple[int, int, int] | None = None,
    ) -> Trial:
        """One playout with the reference's argument meaning (simulate.rs:84-91).

        `prev_action` (compact action code of the move that led to `state`, or None) is the first ply's reply
        context of the LGR / LGRF-2 / NST strategies (simulate.rs select_move); it travels in the leaf record's
        `prev` field (`prev_action_plus1`).  The other strategies ignore it, as in the reference.  `rng` names the
        counter-based stream instead of carrying a SmallRng: (seed, leaf_id, playout_id).
        """
        seed, leaf_id, playout_id = rng if rng is not None else (self.seed, 0, 0)
        if prev_action is not None:
            state = np.array(np.asarray(state).reshape(-1)[:1], dtype=LEAF_DTYPE)  # a copy: the caller's record is untouched
            state["prev"] = int(prev_action) + 1
            state["prev2"] = 0  # a playout's own-previous-move context starts empty (simulate.rs:98)
        depth_cap = self.max_playout_depth if max_playout_depth is None else int(max_playout_depth)
        # playout ids are 0..k-1 inside a batch; run k = playout_id + 1 and keep the last
        max_trace = depth_cap if depth_cap else 512
        r = self.simulate_many(
            state,
            playout_id + 1,
            stats=stats,
            want_trace=True,
            want_final=True,
            max_trace=max_trace,
            seed=seed,
            leaf_id_base=leaf_id,
            max_playout_depth=depth_cap,
        )
        rec = r.records[playout_id]
        depth = int(rec["depth"])
        turn0 = int(np.asarray(state).reshape(-1)[0]["turn"])
        actions = [(int(a), (turn0 + t) & 1) for t, a in enumerate(r.trace[playout_id][:depth])]
        end = EndType(int(rec["end_type"]))
        terminal = Outcome(int(rec["outcome"])) if end == EndType.NATURAL_END else Outcome.NOT_TERMINAL
        return Trial(actions, r.final[playout_id : playout_id + 1].copy(), end, depth, terminal)

    # -- device-resident path and measurement helpers --
    def set_mast(self, stats: MastTable | np.ndarray) -> None:
        mast = np.ascontiguousarray(stats.table if isinstance(stats, MastTable) else stats, dtype=STAT_DTYPE)
        self._check(self.lib.mr_set_mast(self._h, int(self.game), ptr(mast)))

    def set_replies(self, replies: np.ndarray | None) -> None:
        """The Last-Good-Reply table for Policy.EPS_LGR batches: u16 [2][NA], 1 + compact code of player p's reply to
        the preceding action with that dense id, 0 = none (TreeStats.player_replies).  None clears it."""
        if replies is None:
            self._check(self.lib.mr_set_replies(self._h, int(self.game), None))
            return
        tab = np.ascontiguousarray(replies, dtype=np.uint16)
        if tab.shape != (2, self.na):
            raise ValueError(f"reply table must have shape (2, {self.na})")
        self._check(self.lib.mr_set_replies(self._h, int(self.game), ptr(tab)))

    def set_replies2(self, replies2: np.ndarray | None) -> None:
        """The LGRF-2 table for Policy.EPS_LGR2 batches: u16 [2][S][S], S = num_slots(game), indexed [mover][slot of the
        mover's own previous action][slot of the opponent's answer] = 1 + compact code of the reply, 0 = none
        (TreeStats.player_replies2).  None clears it."""
        if replies2 is None:
            self._check(self.lib.mr_set_replies2(self._h, int(self.game), None))
            return
        ns = self.lib.mr_num_slots(int(self.game))
        tab = np.ascontiguousarray(replies2, dtype=np.uint16)
        if tab.shape != (2, ns, ns):
            raise ValueError(f"2-ply reply table must have shape (2, {ns}, {ns})")
        self._check(self.lib.mr_set_replies2(self._h, int(self.game), ptr(tab)))

    def resolve_replies2(self, merged: np.ndarray, table: np.ndarray | None = None, slot: int = 0) -> np.ndarray:
        """Second step of the LGRF-2 update (mr_reply2_resolve): given the last insert per context (the batch's
        `reply2_inserts`, merged with the caller's tree-path windows) and the table being updated, replays the batch and
        returns u8 [2][S][S]: 1 where a later losing playout removes the candidate entry (backprop.rs:961-963)."""
        ns = self.lib.mr_num_slots(int(self.game))
        m = np.ascontiguousarray(merged, dtype=REPLY_UPDATE_DTYPE).copy()
        if m.shape != (2, ns, ns):
            raise ValueError(f"merged inserts must have shape (2, {ns}, {ns})")
        if table is not None:  # contexts without an insert: the entry the table holds now is the candidate
            keep = m["order"] == 0
            m["action_plus1"][keep] = np.asarray(table, dtype=np.uint16)[keep]
        removed = np.zeros((2, ns, ns), dtype=np.uint8)
        self._check(self.lib.mr_reply2_resolve(self._h, slot, ptr(m), ptr(removed)))
        return removed

    @staticmethod
    def apply_replies2(table: np.ndarray, merged: np.ndarray, removed: np.ndarray) -> np.ndarray:
        """New table = none where removed, else the merged insert's action, else the old entry."""
        out = np.where(merged["order"] > 0, merged["action_plus1"], table).astype(np.uint16)
        out[removed != 0] = 0
        return out

    def set_bigrams(self, bigram: np.ndarray | None, backoff_threshold: int | None = None) -> None:
        """TreeStats.player_bigram_actions for Policy.EPS_NST batches: STAT_DTYPE [2][S][S], S = num_slots(game), indexed
        [mover][slot of the preceding action in the other player's numbering][slot of the action].  None = empty."""
        thr = getattr(self.strategy, "backoff_threshold", 5) if backoff_threshold is None else backoff_threshold
        if bigram is None:
            self._check(self.lib.mr_set_bigrams(self._h, int(self.game), None, int(thr)))
            return
        ns = self.lib.mr_num_slots(int(self.game))
        tab = np.ascontiguousarray(bigram, dtype=STAT_DTYPE)
        if tab.shape != (2, ns, ns):
            raise ValueError(f"bigram table must have shape (2, {ns}, {ns})")
        self._check(self.lib.mr_set_bigrams(self._h, int(self.game), ptr(tab), int(thr)))

    def launch_device(
        self, d_leaves: int, d_results: int, n_leaves: int, k: int, *, stream: int = 0, device_index: int = 0,
        d_amaf: int = 0, d_mast_delta: int = 0, seed: int | None = None, leaf_id_base: int = 0,
        max_playout_depth: int | None = None, tables_in_slots: bool = False,
    ) -> None:
        """Launch on caller-owned DEVICE buffers (raw pointers) on `stream` (raw cudaStream_t); nothing is copied.
        With `tables_in_slots` d_amaf / d_mast_delta are [.., 2, num_slots(game)] instead of [.., 2, num_actions]."""
        flags = (WANT_AMAF if d_amaf else 0) | (WANT_MAST_DELTA if d_mast_delta else 0) | (TABLES_IN_SLOTS if tables_in_slots else 0)
        d = self._desc(n_leaves, k, flags, seed, leaf_id_base, 0, max_playout_depth)
        self._check(
            self.lib.mr_launch_device(
                self._h, device_index, C.byref(d), C.c_void_p(d_leaves), C.c_void_p(d_results),
                C.c_void_p(d_amaf) if d_amaf else None, C.c_void_p(d_mast_delta) if d_mast_delta else None,
                C.c_void_p(stream) if stream else None,
            )
        )

    def kernel_launches(self) -> int:
        return int(self.lib.mr_kernel_launches(self._h))

    def measure_int_peak(self, kind: int = 0, device_index: int = 0) -> tuple[float, float]:
        ops, mhz = C.c_double(), C.c_double()
        self._check(self.lib.mr_measure_int_peak(self._h, device_index, kind, C.byref(ops), C.byref(mhz)))
        return ops.value, mhz.value

    def random_positions(self, plies: int, n: int, *, seed: int = 0) -> np.ndarray:
        """`n` synthetic non-terminal positions after `plies` uniformly random plies from the initial
        position, generated ON THE GPU (a depth-cut playout's final state).  Games that end before
        `plies` are replaced by the next playout ids, so the set is deterministic."""
        out = np.zeros(n, dtype=LEAF_DTYPE)
        if plies == 0:
            out[:] = initial_state(self.game)
            return out
        saved = self.strategy
        self.strategy = Uniform()
        try:
            k = max(64, int(n * 1.5) + 16)
            while True:  # the survivors of playout ids 0..k-1 are a prefix of those of a larger k: growing k keeps the set
                r = self.simulate_many(
                    initial_state(self.game), k, want_trace=True, want_final=True, max_trace=1, seed=seed, max_playout_depth=plies
                )
                ok = (r.records["end_type"] == EndType.TURN_LIMIT) & (r.records["depth"] == plies)
                sel = r.final[ok]
                if len(sel) >= n or k >= 256 * n:
                    break
                k *= 4
        finally:
            self.strategy = saved
        if len(sel) < n:
            raise RuntimeError(f"only {len(sel)} of {n} random positions survived {plies} plies; lower plies")
        out[:] = sel[:n]
        return out
