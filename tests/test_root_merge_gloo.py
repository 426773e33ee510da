"""The N>1 path on CPU: world_size-2 gloo processes exercising the sharding and the root-statistics /
MAST-delta merge that bench.py and the search driver use across GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcts_b200 import root_merge
from mcts_b200._lib import STAT_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        visits = rng.integers(0, 1000, size=22)
        scores = rng.integers(-500, 500, size=(22, 2))
        mv, ms = root_merge.merge_root_stats(visits, scores)
        delta = np.zeros((2, 64), dtype=STAT_DTYPE)
        delta["visits"] = rng.integers(0, 50, size=(2, 64))
        delta["score"] = rng.integers(-50, 50, size=(2, 64))
        md = root_merge.merge_mast_delta(delta)
        # the slot-space table bench.py exchanges every step ([2][192] for Breakthrough), summed as int32
        slots = np.zeros((2, 192), dtype=STAT_DTYPE)
        slots["visits"] = rng.integers(0, 4000, size=(2, 192))
        slots["score"] = rng.integers(-4000, 4000, size=(2, 192))
        msl = root_merge.merge_table_i32(slots)
        assert root_merge.merge_table_i32_async(slots).result().tobytes() == msl.tobytes()  # the deferred form of the same merge
        b, e = root_merge.shard_range(4097, rank, world)
        q.put((rank, visits, scores, mv, ms, delta, md, (b, e), slots, msl))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_root_merge():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v_sum = outs[0][1] + outs[1][1]
    s_sum = outs[0][2] + outs[1][2]
    for o in outs:  # every rank holds the same merged arrays (all-reduce semantics)
        assert (o[3] == v_sum).all() and (o[4] == s_sum).all()
        assert (o[6]["visits"] == outs[0][5]["visits"] + outs[1][5]["visits"]).all()
        assert (o[6]["score"] == outs[0][5]["score"] + outs[1][5]["score"]).all()
        assert o[9].dtype == STAT_DTYPE and o[9].shape == (2, 192)
        assert (o[9]["visits"] == outs[0][8]["visits"] + outs[1][8]["visits"]).all()
        assert (o[9]["score"] == outs[0][8]["score"] + outs[1][8]["score"]).all()
    # shards are contiguous, disjoint and cover everything
    assert outs[0][7][0] == 0 and outs[0][7][1] == outs[1][7][0] and outs[1][7][1] == 4097
    # all ranks pick the same final action from the merged visits
    picks = {root_merge.final_action_index(o[3], 5 * 16 + 3) for o in outs}
    assert len(picks) == 1


def test_shard_range_and_leaf_ids():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            spans = [root_merge.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    ids = {root_merge.leaf_id_base(r, 4, s, 4096) for r in range(4) for s in range(10)}
    assert len(ids) == 40
    with pytest.raises(ValueError):
        root_merge.shard_range(10, 2, 2)


def test_final_action_matches_random_best_semantics():
    v = np.array([5, 9, 9, 1, 9])
    # start at i0 = 3, stride = PRIMES[0] % 5 = 3: visiting order 3,1,4,2,0 -> first maximum is index 1
    assert root_merge.final_action_index(v, 3 * 16 + 0) == 1
    # no process group: merge is the identity
    mv, ms = root_merge.merge_root_stats(v, np.zeros((5, 2)))
    assert (mv == v).all()


class _StubSearch:
    """A searcher whose root totals are a fixed function of its rank (the merge logic needs no GPU)."""

    def __init__(self, rank):
        from types import SimpleNamespace

        self.rank = rank
        self.rollout = SimpleNamespace(game=0)

    def choose_action(self, state):
        from types import SimpleNamespace

        from mcts_b200._lib import ROOT_CHILD_DTYPE

        rng = np.random.default_rng(7 + self.rank)
        ch = np.zeros(7, dtype=ROOT_CHILD_DTYPE)
        ch["action"] = np.arange(7)
        ch["visits"] = rng.integers(0, 500, size=7)
        ch["score"] = rng.integers(-300, 300, size=(7, 2))
        return SimpleNamespace(best_action=int(ch["action"][ch["visits"].argmax()]), children=ch)


def _rp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mcts_b200.search import choose_action_root_parallel

        action, visits, scores, local = choose_action_root_parallel(_StubSearch(rank), None, draw=37)
        q.put((rank, action, visits, scores, local.children))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_root_parallel_choice():
    """choose_action_root_parallel (search/parallel.rs:56-115 with one searcher per rank): every rank ends with the same
    merged totals and the same action, the child with the most MERGED visits."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v_sum = outs[0][4]["visits"].astype(np.int64) + outs[1][4]["visits"]
    s_sum = outs[0][4]["score"] + outs[1][4]["score"]
    for o in outs:
        assert (o[2] == v_sum).all() and (o[3] == s_sum).all()
        assert o[1] == int(v_sum.argmax())  # the maximum is unique here, so the tie-break draw does not matter
    assert outs[0][1] == outs[1][1]
