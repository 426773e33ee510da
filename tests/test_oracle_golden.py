"""Pins the CPU oracle (oracle/*.c) against every golden vector / known-answer test the reference's
own tests hold for the playout path.  CPU only.

Each test names the reference file:line whose fixture it replays.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as o

# --------------------------------------------------------------------------------------------
# Philox4x32-10: Random123 known-answer vectors (kat_vectors: philox4x32 10)
# --------------------------------------------------------------------------------------------


def test_philox_random123_kat():
    assert o.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert o.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert o.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09,
        0x94FDCCEB,
        0x5001E420,
        0x24126EA1,
    ]


def test_draw_discipline_counter_layout():
    seed = 0x0123456789ABCDEF
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for ply in range(9):
        # uniform / decisive: one word per ply, four plies per Philox block, stream word 0
        assert o.draw(seed, 7, 11, ply, 0) == o.philox([7, 11, ply >> 2, 0], key)[ply & 3]
        # epsilon-MAST: two words per ply, two plies per block, stream word 1
        block = o.philox([7, 11, ply >> 1, 1], key)
        assert o.draw_pair(seed, 7, 11, ply) == (block[2 * (ply & 1)], block[2 * (ply & 1) + 1])


# --------------------------------------------------------------------------------------------
# Connect4 -- games/connect4/src/lib.rs:468-553
# --------------------------------------------------------------------------------------------


def _c4_play(cols, stop_when_terminal=True, assert_not_terminal=False):
    s = o.initial_state(o.CONNECT4)
    for c in cols:
        if assert_not_terminal:
            assert not o.is_terminal(o.CONNECT4, s), "won before the scripted sequence completed"
        elif stop_when_terminal and o.is_terminal(o.CONNECT4, s):
            break
        assert c in o.generate_actions(o.CONNECT4, s)
        s = o.apply(o.CONNECT4, s, c)
    return s


def test_c4_vertical_win():  # lib.rs:468-479
    s = _c4_play([0, 1, 0, 1, 0, 1, 0])
    assert s["flag"][0] == 1 and o.winner(o.CONNECT4, s) == 0
    assert o.terminal_status(o.CONNECT4, s) == o.WINNER_P0


def test_c4_horizontal_win():  # lib.rs:482-497
    s = _c4_play([0, 4, 1, 4, 2, 4, 3])
    assert s["flag"][0] == 1 and o.winner(o.CONNECT4, s) == 0


def test_c4_diagonal_win():  # lib.rs:506-525
    s = _c4_play([0, 1, 1, 2, 6, 2, 2, 3, 6, 3, 6, 3, 3], assert_not_terminal=True)
    assert s["flag"][0] == 1 and o.winner(o.CONNECT4, s) == 0
    black = int(s["board"][0][0])
    for row, col in [(0, 0), (1, 1), (2, 2), (3, 3)]:
        assert (black >> (row * 7 + col)) & 1
    # the winner keeps the turn (lib.rs:332-336)
    assert s["turn"][0] == 0


def test_c4_initial_and_layout():
    s = o.initial_state(o.CONNECT4)
    assert o.generate_actions(o.CONNECT4, s) == list(range(7))
    s = o.apply(o.CONNECT4, s, 3)  # row 0, col 3 -> bit 3 (idx = row*7 + col, lib.rs:318)
    assert int(s["board"][0][0]) == 1 << 3 and s["turn"][0] == 1
    s = o.apply(o.CONNECT4, s, 3)  # row 1, col 3 -> bit 10
    assert int(s["board"][0][1]) == 1 << 10
    # a full column disappears from the action list
    t = o.initial_state(o.CONNECT4)
    for _ in range(6):
        t = o.apply(o.CONNECT4, t, 0)
    assert o.generate_actions(o.CONNECT4, t) == [1, 2, 3, 4, 5, 6]


def test_c4_draw_exists_on_4x5():  # lib.rs:534-553 (searches seeds 0..2000 for a drawn 4x5 game)
    L = o.lib()
    plies, has_winner = C.c_int(), C.c_int()
    found = False
    for seed in range(2000):
        w = L.orc_c4_generic_random_game(4, 5, seed, C.byref(plies), C.byref(has_winner))
        if not has_winner.value:
            assert w == -1 and plies.value == 20
            found = True
            break
    assert found, "no drawn game found in 2000 random seeds on a 4x5 board"


# --------------------------------------------------------------------------------------------
# Othello -- games/othello/src/lib.rs:771-1052, 1171-1209
# --------------------------------------------------------------------------------------------


def _bits(idx):
    v = 0
    for i in idx:
        v |= 1 << i
    return v


def _d4(i, sym):  # games/game-core/src/symmetry.rs:88-127
    row, col = divmod(i, 8)
    fc = lambda r, c: (r, 7 - c)
    fr = lambda r, c: (7 - r, c)
    tr = lambda r, c: (c, r)
    r, c = row, col
    if sym == 1:
        r, c = fc(r, c)
    elif sym == 2:
        r, c = fr(r, c)
    elif sym == 3:
        r, c = tr(r, c)
    elif sym == 4:
        r, c = fr(*fc(r, c))
    elif sym == 5:
        r, c = tr(*fc(r, c))
    elif sym == 6:
        r, c = tr(*fr(r, c))
    elif sym == 7:
        r, c = tr(*fr(*fc(r, c)))
    return r * 8 + c


def test_d4_doc_example():  # symmetry.rs:81-83
    assert _d4(27, 0) == 27 and _d4(27, 1) == 28


OTHELLO_MOVEGEN_KATS = [  # (black, white, expected) -- lib.rs:771-912
    ([28, 35], [27, 36], [19, 26, 37, 44]),  # initial moves
    ([27], [35], [43]),  # north
    ([43], [35], [27]),  # south
    ([25], [26, 27], [28]),  # east
    ([30], [29, 28], [27]),  # west
    ([19], [28], [37]),  # NE
    ([35], [28], [21]),  # NW
    ([21], [28], [35]),  # SE
    ([37], [28], [19]),  # SW
    ([48], [40], [32]),  # west edge
    ([63], [55], [47]),  # east edge
    ([40, 58], [48, 57], [56]),  # corner multi-direction
    ([40, 58, 42], [48, 57, 49], [56]),  # triple direction
    ([0], [1, 2, 3, 4, 5, 6], [7]),  # long line west
    ([0], [8, 16, 24, 32, 40], [48]),  # long line north
    ([0], [63], []),  # no sandwich possible
]


@pytest.mark.parametrize("black,white,expected", OTHELLO_MOVEGEN_KATS)
def test_othello_movegen_kats_all_symmetries(black, white, expected):
    L = o.lib()
    for sym in range(8):
        p = _bits(_d4(i, sym) for i in black)
        q = _bits(_d4(i, sym) for i in white)
        e = _bits(_d4(i, sym) for i in expected)
        assert L.orc_othello_generate_moves(p, q) == e
        assert L.orc_naive_othello_generate_moves(p, q) == e
        assert e & (p | q) == 0


def test_othello_no_pieces():  # lib.rs:892-904
    L = o.lib()
    assert L.orc_othello_generate_moves(0, _bits([27, 36])) == 0
    assert L.orc_othello_generate_moves(_bits([28, 35]), 0) == 0


def test_othello_initial_position_and_apply():  # lib.rs:916-962
    s = o.initial_state(o.OTHELLO)
    assert int(s["board"][0][0]) == _bits([28, 35]) and int(s["board"][0][1]) == _bits([27, 36])
    assert s["turn"][0] == 0 and not o.is_terminal(o.OTHELLO, s)
    assert o.generate_actions(o.OTHELLO, s) == [19, 26, 37, 44]
    m = o.apply(o.OTHELLO, s, 19)
    black, white = int(m["board"][0][0]), int(m["board"][0][1])
    assert bin(black).count("1") == 4 and (black >> 19) & 1 and (black >> 27) & 1
    assert white == 1 << 36 and m["turn"][0] == 1


def test_othello_pass_handling():  # lib.rs:972-1052
    s = o.make_state([_bits([0]), _bits([63])], turn=0, flag=0)
    assert o.generate_actions(o.OTHELLO, s) == [o.OTHELLO_PASS]
    s = o.apply(o.OTHELLO, s, o.OTHELLO_PASS)
    assert s["flag"][0] == 1 and s["turn"][0] == 1
    assert o.generate_actions(o.OTHELLO, s) == [o.OTHELLO_PASS]
    assert o.is_terminal(o.OTHELLO, s)
    # real move resets last_pass (lib.rs:1009-1052)
    s["board"][0][1] = _bits([26, 27])
    s["board"][0][0] = _bits([25])
    s = o.apply(o.OTHELLO, s, 24)
    assert s["flag"][0] == 0 and s["turn"][0] == 0
    assert int(s["board"][0][0]) == 0 and int(s["board"][0][1]) == _bits([24, 25, 26, 27])


def test_othello_130452_replay():  # lib.rs:1171-1209
    L = o.lib()
    moves = [26, 20, 29, 18, 19, 22, 37, 45, 17, 25, 24, 34, 44, 16, 33, 53, 42, 32, 30, 49, 56, 51, 43]
    s = o.initial_state(o.OTHELLO)
    for ply, mv in enumerate(moves):
        b, w = int(s["board"][0][0]), int(s["board"][0][1])
        p, q = (b, w) if s["turn"][0] == 0 else (w, b)
        assert b & w == 0
        assert (L.orc_othello_generate_moves(p, q) >> mv) & 1, f"ply {ply}: move {mv} not legal"
        assert L.orc_othello_get_flips(p, q, 1 << mv) == L.orc_naive_othello_get_flips(p, q, 1 << mv)
        s = o.apply(o.OTHELLO, s, mv)
    occ = int(s["board"][0][0]) | int(s["board"][0][1])
    assert bin(occ).count("1") == 27
    assert int(s["board"][0][0]) & int(s["board"][0][1]) == 0


def test_othello_symmetry_stress_sweep():
    """mcts-tests/tests/stress.rs:249-454 restated (same xorshift64* stream, seed 0xdead_beef_cafe_babe).
    The reference runs 100 000 games under `cargo test --test stress`; 4 000 here keeps the CPU suite
    short (tests/test_oracle_stress_full.py holds the full-size run behind ORACLE_STRESS_FULL=1; profiles/oracle_stress_full_r02.txt
    records it)."""
    games, plies = C.c_uint64(), C.c_uint64()
    rc = o.lib().orc_stress_othello(4000, C.byref(games), C.byref(plies))
    assert rc == 0, f"mismatch code {rc} in game {games.value}"
    assert plies.value > 4000 * 55


def test_othello_random_games_end_within_200_actions():  # stress.rs:161-241 (2 000 games <= 200 actions)
    init = o.initial_state(o.OTHELLO)
    out = o.playout_batch(o.OTHELLO, init, 2000, seed=99, want_final=True)
    assert out["records"]["depth"].max() <= 200
    assert (out["records"]["end_type"] == o.END_TERMINAL).all()
    # draws have equal disc counts, winners have strictly more (lib.rs:1150-1168)
    for rec, fin in zip(out["records"][:200], out["final"][:200]):
        b = bin(int(fin["board"][0])).count("1")
        w = bin(int(fin["board"][1])).count("1")
        if rec["outcome"] == o.DRAW:
            assert b == w
        elif rec["outcome"] == o.WINNER_P0:
            assert b > w
        else:
            assert w > b


# --------------------------------------------------------------------------------------------
# Breakthrough / Knightthrough -- mcts-tests/tests/stress.rs:485-818, games/breakthrough/src/main.rs:119-126
# --------------------------------------------------------------------------------------------


@pytest.mark.parametrize("game", [o.BREAKTHROUGH, o.KNIGHTTHROUGH])
def test_through_initial_boards(game):
    s = o.initial_state(game)
    assert int(s["board"][0][0]) == 0xFFFF000000000000 and int(s["board"][0][1]) == 0xFFFF
    assert s["turn"][0] == 0 and s["flag"][0] == 0


def test_breakthrough_initial_moves_order():
    # Black (rows 6-7) moves south: src ascending, dst ascending => src-9, src-8, src-7 (lib.rs:131-156)
    acts = o.generate_actions(o.BREAKTHROUGH, o.initial_state(o.BREAKTHROUGH))
    pairs = [(a >> 8, a & 0xFF) for a in acts]
    assert pairs[:5] == [(48, 40), (48, 41), (49, 40), (49, 41), (49, 42)]
    assert len(pairs) == 22 and pairs == sorted(pairs)


def test_knightthrough_initial_moves():
    acts = o.generate_actions(o.KNIGHTTHROUGH, o.initial_state(o.KNIGHTTHROUGH))
    pairs = [(a >> 8, a & 0xFF) for a in acts]
    assert pairs == sorted(pairs)
    # from 48 (row 6, col 0): (4,1)=33, (5,2)=42 ; (7,2)=58 is own
    assert [p for p in pairs if p[0] == 48] == [(48, 33), (48, 42)]


@pytest.mark.parametrize("game", [o.BREAKTHROUGH, o.KNIGHTTHROUGH])
def test_through_oracle_stress_5000(game):
    """stress.rs:638-818 restated at its full size (5 000 games, seed 0xcafe_babe_dead_beef)."""
    goals, stuck = C.c_uint64(), C.c_uint64()
    rc = o.lib().orc_stress_through(game, 5000, C.byref(goals), C.byref(stuck))
    assert rc == 0, f"mismatch code {rc}"
    assert goals.value + stuck.value == 5000


# --------------------------------------------------------------------------------------------
# Hex 11x11 -- gdl/tests/hex_gen_oracle.rs:63-75, 213-248
# --------------------------------------------------------------------------------------------

SIDE, SITES = 11, 121


def _hex_assert_agrees(sites):
    """hex_gen_oracle.rs:140-177 assert_agrees: legal moves, terminal-ness and winner after every move."""
    L = o.lib()
    s = o.initial_state(o.HEX11)
    occ = [np.zeros(SITES, dtype=np.uint8), np.zeros(SITES, dtype=np.uint8)]
    to_move = 0

    def oracle_legal():
        return [i for i in range(SITES) if not occ[0][i] and not occ[1][i]]

    def oracle_winner():
        return L.orc_naive_hex_winner(occ[0].ctypes.data, occ[1].ctypes.data, 1 - to_move)

    assert o.generate_actions(o.HEX11, s) == oracle_legal()
    assert not o.is_terminal(o.HEX11, s)
    for step, site in enumerate(sites):
        occ[to_move][site] = 1
        to_move = 1 - to_move
        s = o.apply(o.HEX11, s, site)
        assert o.generate_actions(o.HEX11, s) == oracle_legal(), f"legal moves disagree after move {step}"
        ow = oracle_winner()
        oterm = ow >= 0 or not oracle_legal()
        assert o.is_terminal(o.HEX11, s) == oterm, f"terminal-ness disagrees after move {step}"
        if oterm:
            assert o.winner(o.HEX11, s) == ow
    return s


def _interleave(p0, p1):
    seq = []
    for i in range(max(len(p0), len(p1))):
        if i < len(p0):
            seq.append(p0[i])
        if i < len(p1):
            seq.append(p1[i])
    return seq


def _filler(used, count):
    return [s for s in range(SITES) if s not in used][:count]


def test_hex_in_progress():  # hex_gen_oracle.rs:213-221
    _hex_assert_agrees([4, 0])
    _hex_assert_agrees([0, 1, 3, 5])


def test_hex_p0_wins_straight_column():  # :223-228
    p0 = [r * SIDE for r in range(SIDE)]
    s = _hex_assert_agrees(_interleave(p0, _filler(p0, len(p0) - 1)))
    assert o.winner(o.HEX11, s) == 0 and o.terminal_status(o.HEX11, s) == o.WINNER_P0


def test_hex_p1_wins_hex_adjacent_diagonal():  # :230-235
    p1 = [r * SIDE + r for r in range(SIDE)]
    s = _hex_assert_agrees(_interleave(_filler(p1, len(p1)), p1))
    assert o.winner(o.HEX11, s) == 1


def test_hex_non_adjacent_diagonal_does_not_connect():  # :237-242
    p1 = [r * SIDE + (SIDE - 1 - r) for r in range(SIDE)]
    s = _hex_assert_agrees(_interleave(_filler(p1, len(p1)), p1))
    assert o.winner(o.HEX11, s) == -1 and not o.is_terminal(o.HEX11, s)


def test_hex_full_board_row_major_fill():  # :244-248
    _hex_assert_agrees(list(range(SITES)))


def test_hex_layout_two_words():
    # idx = row*11 + col stored at word idx/64, bit idx%64, low word first (bitboard/src/board.rs:89-98)
    s = o.apply(o.HEX11, o.initial_state(o.HEX11), 70)
    assert int(s["board"][0][1]) == 1 << 6 and int(s["board"][0][0]) == 0
    s = o.apply(o.HEX11, s, 3)
    assert int(s["board"][0][2]) == 1 << 3


# --------------------------------------------------------------------------------------------
# Playout control flow and policies (pinned by reading code only; SURVEY section 8c "unpinned")
# --------------------------------------------------------------------------------------------


def test_playout_marginals_match_survey_sanity_anchors():
    """uniform playouts from the initial position."""
    anchors = {  # game: (mean plies, P0 win rate, tolerance plies, tolerance rate)
        o.CONNECT4: (21.4, 0.557, 0.4, 0.02),
        o.BREAKTHROUGH: (64.1, 0.522, 1.5, 0.03),
        o.OTHELLO: (60.4, 0.478, 0.6, 0.03),
        o.HEX11: (107.2, 0.505, 1.5, 0.04),
    }
    for game, (plies, p0, tp, tr) in anchors.items():
        k = 4000 if game != o.HEX11 else 1500
        out = o.playout_batch(game, o.initial_state(game), k, seed=12345, threads=o.host_threads())
        r = out["results"][0]
        assert abs(r["sum_depth"] / k - plies) < tp, (o.GAME_NAMES[game], r["sum_depth"] / k)
        assert abs(r["wins"][0] / k - p0) < tr, (o.GAME_NAMES[game], r["wins"][0] / k)
        assert r["wins"][0] + r["wins"][1] + r["draws"] == k


def test_depth_cutoff_is_a_draw():  # simulate.rs:104-108, backprop.rs:686-690
    out = o.playout_batch(o.KNIGHTTHROUGH, o.initial_state(o.KNIGHTTHROUGH), 64, max_depth=10, seed=1, want_final=True)
    r = out["results"][0]
    assert r["cutoffs"] == 64 and r["draws"] == 64 and r["sum_depth"] == 640
    assert (out["records"]["end_type"] == o.END_TURN_LIMIT).all()


def test_no_actions_not_terminal_is_a_draw():  # simulate.rs:112-116; stress.rs:721-731
    # Black to move with a single blocked pawn: straight ahead occupied, both diagonals own/off-board
    black = _bits([8, 1])  # pawn on a2 (8); own piece on b1 (1) blocks the east diagonal... a1 is straight
    white = _bits([0])  # white on a1 blocks the straight move
    s = o.make_state([black, white], turn=0, flag=0)
    acts = o.generate_actions(o.BREAKTHROUGH, s)
    # piece at 1 is on row 0: it has moved "past" the board for Black; shift_south drops it => no moves
    assert acts == []
    out = o.playout_batch(o.BREAKTHROUGH, s, 3, seed=5, want_final=True)
    assert out["results"][0]["draws"] == 3 and out["results"][0]["sum_depth"] == 0
    assert (out["records"]["end_type"] == o.END_NO_ACTIONS).all()


def test_decisive_win_takes_the_win_and_prefers_terminal_loss_quirk():  # simulate.rs:662-687
    # Connect4: Black has three in column 0; decisive policy must always play column 0 and win at depth 1
    s = _c4_play([0, 1, 0, 1, 0, 1])
    out = o.playout_batch(o.CONNECT4, s, 50, policy=o.DECISIVE_WIN, seed=3, want_trace=True, max_trace=4)
    assert (out["records"]["depth"] == 1).all() and (out["records"]["outcome"] == o.WINNER_P0).all()
    assert (out["trace"][:, 0] == 0).all()
    # Othello: with our keyed draws the decisive scan never changes the pick, so the
    # two policies give identical playouts
    pos = o.random_positions(o.OTHELLO, 4, 50, 16)
    a = o.playout_batch(o.OTHELLO, pos, 8, policy=o.UNIFORM, seed=8, want_trace=True, max_trace=80)
    b = o.playout_batch(o.OTHELLO, pos, 8, policy=o.DECISIVE_WIN, seed=8, want_trace=True, max_trace=80)
    assert (a["trace"] == b["trace"]).all() and (a["records"] == b["records"]).all()


def test_mast_empty_table_picks_random_best_start():  # util.rs:147-161 with all scores 1.0
    game = o.BREAKTHROUGH
    na = o.num_actions(game)
    mast = np.zeros((2, na), dtype=o.STAT_DTYPE)
    s = o.initial_state(game)
    out = o.playout_batch(game, s, 32, policy=o.EPS_MAST, eps=0.0, mast=mast, seed=21, want_trace=True, max_trace=1, max_depth=1)
    acts = o.generate_actions(game, s)
    for pid in range(32):
        x0, _ = o.draw_pair(21, 0, pid, 0)
        r = (x0 * 16 * len(acts)) >> 32
        assert out["trace"][pid, 0] == acts[r // 16]


def test_mast_prefers_best_scored_action_and_delta_counts():
    game = o.BREAKTHROUGH
    na = o.num_actions(game)
    s = o.initial_state(game)
    acts = o.generate_actions(game, s)
    mast = np.zeros((2, na), dtype=o.STAT_DTYPE)
    for a in acts:  # every legal move visited with a poor score, one with a perfect score
        aid = (a >> 8) * 64 + (a & 0xFF)
        mast[0, aid] = (10, -2)
    fav = acts[7]
    mast[0, (fav >> 8) * 64 + (fav & 0xFF)] = (10, 9)
    out = o.playout_batch(
        game, s, 40, policy=o.EPS_MAST, eps=0.0, mast=mast, seed=2, want_trace=True, max_trace=1, max_depth=1, want_mast_delta=True
    )
    assert (out["trace"][:, 0] == fav).all()
    d = out["mast_delta"]
    assert d["visits"].sum() == 40 and d[0, (fav >> 8) * 64 + (fav & 0xFF)]["visits"] == 40
    assert d["score"].sum() == 0  # depth cutoff => draw => zero utility


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11])
def test_nst_backs_off_to_mast_and_bigram_delta_sums_to_mast_delta(game):  # simulate.rs:898-930, backprop.rs:885-899
    """Nst with no bigram entry at the threshold IS Mast; and since every playout ply of a leaf that carries its
    preceding action has a predecessor, the bigram delta summed over predecessors equals the MAST delta."""
    depth = 200 if game == o.KNIGHTTHROUGH else 0
    leaves = np.concatenate([o.random_positions(game, 3, 6, 4), o.random_positions(game, 4, 13, 4)])
    uni = o.playout_batch(game, leaves, 8, seed=2, max_depth=depth, want_trace=True, max_trace=1)
    ns, na = o.num_slots(game), o.num_actions(game)
    for i in range(len(leaves)):  # some action of the player who is not to move (only its slot matters here)
        p = 1 - int(leaves[i]["turn"])
        codes = [c for c in range(1 << 14) if o.action_slot(game, p, c) >= 0] if i == 0 else codes_by_player[p]
        if i == 0:
            codes_by_player = {p: codes, 1 - p: [c for c in range(1 << 14) if o.action_slot(game, 1 - p, c) >= 0]}
        leaves["prev"][i] = codes_by_player[p][(7 * i) % len(codes_by_player[p])] + 1
    base = o.playout_batch(game, leaves, 32, seed=5, max_depth=depth, want_mast_delta=True, want_bigram_delta=True)
    mast, bigram = base["mast_delta"], base["bigram_delta"]
    aid = (lambda c: (c >> 8) * 64 + (c & 0xFF)) if game in (o.BREAKTHROUGH, o.KNIGHTTHROUGH) else (lambda c: c)
    for p in range(2):
        col_v, col_s = bigram["visits"][p].sum(axis=0), bigram["score"][p].sum(axis=0)
        for slot in range(ns):
            code = next((c for c in codes_by_player[p] if o.action_slot(game, p, c) == slot), None)
            if code is None:
                assert col_v[slot] == 0
                continue
            assert col_v[slot] == mast["visits"][p, aid(code)] and col_s[slot] == mast["score"][p, aid(code)]
    kw = dict(policy=o.EPS_NST, eps=0.0, seed=9, max_depth=depth, mast=mast, want_trace=True, max_trace=64)
    as_mast = o.playout_batch(game, leaves, 16, **{**kw, "policy": o.EPS_MAST})
    never = o.playout_batch(game, leaves, 16, bigram=bigram, nst_threshold=1 << 30, **kw)
    empty = o.playout_batch(game, leaves, 16, bigram=None, **kw)
    assert (never["trace"] == as_mast["trace"]).all() and (empty["trace"] == as_mast["trace"]).all()
    used = o.playout_batch(game, leaves, 16, bigram=bigram, nst_threshold=1, **kw)
    assert not (used["trace"] == as_mast["trace"]).all()  # the bigram scores changed picks


@pytest.mark.parametrize("game", [o.CONNECT4, o.BREAKTHROUGH, o.KNIGHTTHROUGH, o.OTHELLO, o.HEX11])
def test_random_play_terminates(game):
    """The games' own smoke tests (`random_play::<G>()`, mcts/src/util.rs:227-233; connect4/lib.rs:461, breakthrough/lib.rs:270,
    knightthrough/lib.rs:264, hex-gen/lib.rs:265, othello's equivalent): uniformly random self-play from the initial
    position reaches a terminal state -- here 400 games per game, each ending NaturalEnd with a terminal status, within
    the game's own length bound (Knightthrough, whose knights also jump backwards, within 2000 plies)."""
    bound = {o.CONNECT4: 42, o.BREAKTHROUGH: 224, o.KNIGHTTHROUGH: 2000, o.OTHELLO: 128, o.HEX11: 121}[game]
    r = o.playout_batch(game, o.initial_state(game), 400, seed=17, max_depth=bound + 1, want_final=True, want_trace=True, max_trace=1)
    rec = r["records"]
    assert (rec["end_type"] == o.END_TERMINAL).all()
    assert (rec["depth"] <= bound).all() and (rec["depth"] > 0).all()
    for s in r["final"][:50]:
        assert o.terminal_status(game, np.array([s])) != o.NOT_TERMINAL
    if game == o.HEX11:  # Hex has no draws
        assert (rec["outcome"] != o.DRAW).all()


# ---- the reference's policy KATs (mcts-tests), which it states on TicTacToe -- a game the rollout path does not carry.
# ---- The strategies are game-independent, so the same situations are set up on Connect4 and driven through the oracle's
# ---- playout loop (one ply: max_depth = 1, the trace holds the pick).

def _first_moves(state, policy, n=64, **kw):
    r = o.playout_batch(o.CONNECT4, state, n, policy=policy, seed=5, max_depth=1, want_trace=True, max_trace=1, **kw)
    return set(int(a) for a in r["trace"][:, 0])


def test_anti_decisive_blocks_the_opponents_only_winning_reply():  # mcts-tests/tests/decisive_move.rs:10-37
    """White threatens d1 (a1 b1 c1 on the bottom row); Black has no win of its own, and every move but the block hands
    White that win on the next reply: the block is the unique anti-decisive choice although it wins nothing."""
    s = _c4_play([6, 0, 6, 1, 5, 2])  # Black: g1 g2 f1, White: a1 b1 c1; Black to move
    assert int(s[0]["turn"]) == 0
    assert _first_moves(s, o.DECISIVE_ANTI) == {3}
    assert len(_first_moves(s, o.UNIFORM, n=256)) == 7  # (the uniform policy plays everything here)


def test_anti_decisive_still_prefers_an_immediate_win():  # decisive_move.rs:39-61
    """Black wins at d1 and White separately threatens g4: Black must take its own win rather than block."""
    s = _c4_play([0, 6, 1, 6, 2, 6])  # Black: a1 b1 c1, White: g1 g2 g3
    assert _first_moves(s, o.DECISIVE_ANTI) == {3}
    assert _first_moves(s, o.DECISIVE_WIN) == {3}


def test_nst_bigram_table_populated_by_backprop():  # mcts-tests/tests/ttt_strategies.rs:916-991
    """A two-ply playout has exactly one consecutive pair: it lands in the SECOND mover's table with one visit, keyed by
    (first action, second action), and the first mover's table stays empty."""
    s = o.initial_state(o.CONNECT4)
    mast = np.zeros((2, 7), dtype=o.STAT_DTYPE)
    r = o.playout_batch(o.CONNECT4, s, 1, policy=o.EPS_NST, eps=0.1, seed=7, max_depth=2, mast=mast, want_trace=True, max_trace=2,
                        want_bigram_delta=True)
    assert int(r["records"][0]["depth"]) == 2
    first, second = int(r["trace"][0][0]), int(r["trace"][0][1])
    bg = r["bigram_delta"]
    assert bg["visits"][1, first, second] == 1 and bg["visits"][1].sum() == 1
    assert bg["visits"][0].sum() == 0  # not (mis)attributed to the other player's table


def test_nst_backoff_falls_back_to_unigram_below_threshold():  # ttt_strategies.rs:993-1075
    """Unigram favours column 1 (average 1.0 against 0.0 for column 0); the bigram in the context of the preceding action
    favours column 0 but with too few samples.  (The KAT restricts `available` to the two moves; here the other five
    columns are given the worst possible average instead.)"""
    s = o.initial_state(o.CONNECT4).copy()
    s["prev"] = 2 + 1  # the preceding action: column 2
    mast = np.zeros((2, 7), dtype=o.STAT_DTYPE)
    mast["visits"][0] = 10
    mast["score"][0] = [0, 10, -10, -10, -10, -10, -10]
    bigram = np.zeros((2, 7, 7), dtype=o.STAT_DTYPE)
    bigram["visits"][0, 2, 0], bigram["score"][0, 2, 0] = 2, 2
    bigram["visits"][0, 2, 1], bigram["score"][0, 2, 1] = 2, 0
    kw = dict(eps=0.0, mast=mast, bigram=bigram)
    assert _first_moves(s, o.EPS_NST, nst_threshold=5, **kw) == {1}  # 2 samples < 5: the unigram score decides
    bigram["visits"][0, 2, 0], bigram["score"][0, 2, 0] = 5, 5
    bigram["visits"][0, 2, 1] = 5
    assert _first_moves(s, o.EPS_NST, nst_threshold=5, **kw) == {0}  # at the threshold the bigram score wins
    s["prev"] = 0  # no preceding action (a playout straight from the root): always the unigram score
    assert _first_moves(s, o.EPS_NST, nst_threshold=1, **kw) == {1}
