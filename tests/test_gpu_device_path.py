"""The device-resident entry point (mr_launch_device: caller-owned HBM buffers, caller's stream) and the
multi-GPU context (one process driving several GPUs, host gather) against the host-buffer path / the oracle."""
import numpy as np
import pytest

import oracle_lib as o

pytestmark = pytest.mark.gpu
mb = pytest.importorskip("mcts_b200")
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("game,policy,amaf", [(o.BREAKTHROUGH, o.EPS_MAST, False), (o.HEX11, o.UNIFORM, True), (o.CONNECT4, o.UNIFORM, False)])
def test_launch_device_matches_host_path(game, policy, amaf):
    dev = torch.device("cuda", 0)
    n, k = 48, 160
    leaves = o.random_positions(game, 3, 10, n)
    strat = mb.EpsilonGreedyMast(0.1) if policy == o.EPS_MAST else mb.Uniform()
    depth = 200 if policy == o.EPS_MAST else 0
    with mb.GpuRollout(mb.Game(game), strat, seed=5, max_playout_depth=depth) as g:
        na = g.na
        mast = None
        if policy == o.EPS_MAST:
            mast = o.playout_batch(game, leaves, 16, seed=1, max_depth=depth, want_mast_delta=True)["mast_delta"]
            g.set_mast(mast)
        ref = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_mast_delta=policy == o.EPS_MAST, want_amaf=amaf, leaf_id_base=9)
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            d_leaves = torch.from_numpy(leaves.view(np.uint8).reshape(-1).copy()).to(dev)
            d_res = torch.zeros(n * 24, dtype=torch.uint8, device=dev)
            d_delta = torch.zeros(2 * na * 2, dtype=torch.int32, device=dev) if policy == o.EPS_MAST else None
            d_amaf = torch.zeros(n * 2 * na * 2, dtype=torch.int32, device=dev) if amaf else None
            g.launch_device(d_leaves.data_ptr(), d_res.data_ptr(), n, k, stream=stream.cuda_stream, leaf_id_base=9,
                            d_mast_delta=d_delta.data_ptr() if d_delta is not None else 0,
                            d_amaf=d_amaf.data_ptr() if d_amaf is not None else 0)
        stream.synchronize()
        assert d_res.cpu().numpy().tobytes() == ref.results.tobytes()
        if d_delta is not None:
            assert d_delta.cpu().numpy().tobytes() == ref.mast_delta.tobytes()
        if d_amaf is not None:
            assert d_amaf.cpu().numpy().tobytes() == ref.amaf.tobytes()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
@pytest.mark.parametrize("game,policy", [(o.BREAKTHROUGH, o.EPS_MAST), (o.HEX11, o.UNIFORM), (o.OTHELLO, o.UNIFORM)])
def test_two_gpu_context_is_bit_identical(game, policy):
    """mr_init with two devices: the global playout range is split over the GPUs, results are summed on the host;
    playout ids are global, so every output equals the single-GPU one (and the oracle's)."""
    n, k = 5, 333  # a single leaf's playouts straddle both devices
    leaves = o.random_positions(game, 4, 12, n)
    strat = mb.EpsilonGreedyMast(0.1) if policy == o.EPS_MAST else mb.Uniform()
    depth = 200 if policy == o.EPS_MAST else 0
    mast = o.playout_batch(game, leaves, 8, seed=1, max_depth=depth, want_mast_delta=True)["mast_delta"] if policy == o.EPS_MAST else None
    mt = 256
    outs = []
    for devices in (1, 2):
        with mb.GpuRollout(mb.Game(game), strat, seed=77, max_playout_depth=depth, devices=devices) as g:
            outs.append(g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=mast, want_trace=True, want_final=True, max_trace=mt,
                                        want_mast_delta=policy == o.EPS_MAST, want_amaf=game == o.HEX11))
    a, b = outs
    assert a.results.tobytes() == b.results.tobytes()
    assert (a.trace == b.trace).all() and a.records.tobytes() == b.records.tobytes() and a.final.tobytes() == b.final.tobytes()
    if a.mast_delta is not None:
        assert a.mast_delta.tobytes() == b.mast_delta.tobytes()
    if a.amaf is not None:
        assert a.amaf.tobytes() == b.amaf.tobytes()
    ref = o.playout_batch(game, leaves, k, policy=policy, max_depth=depth, seed=77, mast=mast, threads=o.host_threads())
    assert ref["results"].tobytes() == b.results.tobytes()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
@pytest.mark.parametrize("game", [o.BREAKTHROUGH, o.CONNECT4])
def test_two_gpu_context_table_policies(game):
    """The table-updating policies merge differently across the ctx's devices: NST bigram deltas are summed, LGR / LGRF-2
    insert stamps are max-merged, and the LGRF-2 removals come from a replay on every device.  One and two devices must
    report identical tables."""
    n, k = 6, 400
    leaves = o.random_positions(game, 4, 10, n)
    tr = o.playout_batch(game, leaves, 4, seed=2, want_trace=True, max_trace=4, max_depth=200)["trace"]
    for i in range(n):  # plausible preceding actions: moves of uniform playouts from the same position colour
        j = (i + 1) % n
        if leaves[j]["turn"] != leaves[i]["turn"]:
            leaves["prev"][i] = int(tr[j * 4][0]) + 1
    base = o.playout_batch(game, leaves, 64, seed=3, max_depth=200, want_mast_delta=True, want_bigram_delta=True)
    ns = o.num_slots(game)
    t2 = np.zeros((2, ns, ns), dtype=np.uint16)
    outs = []
    for devices in (1, 2):
        with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyNst(0.1, 1), seed=11, max_playout_depth=200, devices=devices) as g:
            g.set_bigrams(base["bigram_delta"])
            nst = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, stats=base["mast_delta"], want_mast_delta=True)
        with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyLgr2(0.1), seed=11, max_playout_depth=200, devices=devices) as g:
            first = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k)  # fills both tables from nothing
            rem1 = g.resolve_replies2(first.reply2_inserts, t2)
            tab1 = np.where(first.reply_updates["order"] > 0, first.reply_updates["action_plus1"], 0).astype(np.uint16)
            tab2 = mb.GpuRollout.apply_replies2(t2, first.reply2_inserts, rem1)
            g.set_replies(tab1)
            g.set_replies2(tab2)
            second = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, leaf_id_base=100)
            rem2 = g.resolve_replies2(second.reply2_inserts, tab2)
        outs.append((nst, first, rem1, second, rem2))
    (a_nst, a1, a_r1, a2, a_r2), (b_nst, b1, b_r1, b2, b_r2) = outs
    assert a_nst.results.tobytes() == b_nst.results.tobytes() and a_nst.bigram_delta.tobytes() == b_nst.bigram_delta.tobytes()
    assert a_nst.mast_delta.tobytes() == b_nst.mast_delta.tobytes()
    for x, y in ((a1, b1), (a2, b2)):
        assert x.results.tobytes() == y.results.tobytes()
        assert x.reply_updates.tobytes() == y.reply_updates.tobytes() and (x.last_win == y.last_win).all()
        assert x.reply2_inserts.tobytes() == y.reply2_inserts.tobytes() and (x.last_lose == y.last_lose).all()
    assert (a_r1 == b_r1).all() and (a_r2 == b_r2).all()
    assert a_r2.any() and (a2.reply2_inserts["order"] > 0).any()
    # and the single-device tables are the oracle's literal ones
    ref = o.playout_batch(game, leaves, k, policy=o.EPS_LGR2, eps=0.1, seed=11, max_depth=200, want_final_tables=True, threads=o.host_threads())
    assert (mb.GpuRollout.apply_replies2(t2, a1.reply2_inserts, a_r1) == ref["replies2_final"]).all()


@pytest.mark.parametrize("game,policy", [(o.BREAKTHROUGH, o.EPS_MAST), (o.KNIGHTTHROUGH, o.EPS_MAST), (o.HEX11, o.EPS_MAST), (o.HEX11, o.UNIFORM),
                                         (o.CONNECT4, o.UNIFORM)])
def test_tables_in_slot_space_equal_the_dense_tables(game, policy):
    """MR_TABLES_IN_SLOTS: the AMAF tables and the MAST delta indexed by compact per-player action slot hold exactly the
    dense tables' entries (host path and device-resident path), and nothing else."""
    from mcts_b200.rollout import slots_to_dense

    dev = torch.device("cuda", 0)
    n, k = 40, 96
    leaves = o.random_positions(game, 3, 10, n)
    strat = mb.EpsilonGreedyMast(0.1) if policy == o.EPS_MAST else mb.Uniform()
    with mb.GpuRollout(mb.Game(game), strat, seed=5, max_playout_depth=200) as g:
        mast = o.playout_batch(game, leaves, 16, seed=1, max_depth=200, want_mast_delta=True)["mast_delta"] if policy == o.EPS_MAST else None
        kw = dict(stats=mast, want_mast_delta=True, want_amaf=True, leaf_id_base=3)
        dense = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, **kw)
        slots = g.simulate_many(leaves.view(mb.LEAF_DTYPE), k, tables_in_slots=True, **kw)
        ns = mb.rollout.num_slots(g.game)
        assert slots.mast_delta.shape == (2, ns) and slots.amaf.shape == (n, 2, ns)
        assert slots.results.tobytes() == dense.results.tobytes()
        assert slots_to_dense(g.game, slots.mast_delta).tobytes() == dense.mast_delta.tobytes()
        assert slots_to_dense(g.game, slots.amaf).tobytes() == dense.amaf.tobytes()
        assert int(slots.mast_delta["visits"].sum()) == int(dense.mast_delta["visits"].sum()) == int(dense.results["sum_depth"].sum())
        # device-resident path, slot space
        if mast is not None:
            g.set_mast(mast)
        d_leaves = torch.from_numpy(leaves.view(np.uint8).reshape(-1).copy()).to(dev)
        d_res = torch.zeros(n * 24, dtype=torch.uint8, device=dev)
        d_delta = torch.zeros(2 * ns * 2, dtype=torch.int32, device=dev)
        d_amaf = torch.zeros(n * 2 * ns * 2, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        g.launch_device(d_leaves.data_ptr(), d_res.data_ptr(), n, k, leaf_id_base=3, d_mast_delta=d_delta.data_ptr(), d_amaf=d_amaf.data_ptr(),
                        tables_in_slots=True)
        torch.cuda.synchronize()
        assert d_delta.cpu().numpy().tobytes() == slots.mast_delta.tobytes()
        assert d_amaf.cpu().numpy().tobytes() == slots.amaf.tobytes()


def test_wait_returns_while_a_later_slot_is_still_running():
    """Every slot has its own stream and completion signal: mr_wait(slot 0) must come back when slot 0's small batch is
    done, not when the large batch submitted after it on slot 1 is (round 1 waited on one shared stream)."""
    import time

    leaves = mb.initial_state(mb.Game.OTHELLO)
    small = np.repeat(leaves, 32)
    big = np.repeat(leaves, 4096)
    with mb.GpuRollout(mb.Game.OTHELLO, mb.Uniform(), seed=1) as g:
        g.simulate_many(big, 256)  # warm: buffers allocated, kernel loaded
        g.simulate_many(small, 32)
        torch.cuda.synchronize()
        g.submit(small, 32, slot=0)
        g.submit(big, 4096, slot=1)  # 2^24 playouts of 60 plies: ~18 ms
        t0 = time.perf_counter()
        r0 = g.wait(0)
        t_small = time.perf_counter() - t0
        r1 = g.wait(1)
        t_big = time.perf_counter() - t0
    assert int(r0.results["wins"].sum() + r0.results["draws"].sum()) == 32 * 32
    assert int(r1.results["wins"].sum() + r1.results["draws"].sum()) == 4096 * 4096
    assert t_big > 5e-3, t_big
    assert t_small < 0.25 * t_big, (t_small, t_big)


def test_four_slots_in_flight_are_independent_and_ordered_by_slot_only():
    """MR_NUM_SLOTS batches submitted back to back, waited in reverse order: each slot returns its own batch's results,
    bit-identical to the same batch run alone."""
    game = o.BREAKTHROUGH
    leaves = o.random_positions(game, 3, 10, 24)
    mast = o.playout_batch(game, leaves, 16, seed=1, max_depth=200, want_mast_delta=True)["mast_delta"]
    with mb.GpuRollout(mb.Game(game), mb.EpsilonGreedyMast(0.1), seed=5, max_playout_depth=200) as g:
        alone = [g.simulate_many(leaves.view(mb.LEAF_DTYPE), 64 + 32 * s, stats=mast, want_mast_delta=True, leaf_id_base=100 * s) for s in range(4)]
        for rep in range(3):  # the fused epilogue leaves the device buffers clean for the slot's next batch
            for s in range(4):
                g.submit(leaves.view(mb.LEAF_DTYPE), 64 + 32 * s, slot=s, stats=mast, want_mast_delta=True, leaf_id_base=100 * s)
            for s in reversed(range(4)):
                r = g.wait(s)
                assert r.results.tobytes() == alone[s].results.tobytes(), (rep, s)
                assert r.mast_delta.tobytes() == alone[s].mast_delta.tobytes(), (rep, s)
