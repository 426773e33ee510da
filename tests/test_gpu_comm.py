"""mr_comm_*: the C-ABI root-statistics merge over NCCL (one rank per GPU), for hosts without torch.distributed."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
mb = pytest.importorskip("mcts_b200")
torch = pytest.importorskip("torch")


def _rank(rank, world, uid_bytes, q):
    import ctypes as C

    import numpy as np
    import torch

    import mcts_b200 as mb

    torch.cuda.set_device(rank)
    g = mb.GpuRollout(mb.Game.BREAKTHROUGH, devices=[rank])
    L = g.lib
    uid = (C.c_uint8 * 128).from_buffer_copy(uid_bytes)
    rc = L.mr_comm_init(g._h, uid, rank, world)
    if rc != 0:
        q.put((rank, "init failed: " + L.mr_last_error(g._h).decode()))
        return
    buf = torch.arange(48 * 3, dtype=torch.int32, device=f"cuda:{rank}") * (rank + 1)
    torch.cuda.synchronize()
    rc = L.mr_comm_allreduce_i32(g._h, C.c_void_p(buf.data_ptr()), buf.numel(), None)
    if rc != 0:
        q.put((rank, "allreduce failed: " + L.mr_last_error(g._h).decode()))
        return
    torch.cuda.synchronize()
    q.put((rank, buf.cpu().numpy()))
    g.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_root_merge_through_the_c_abi():
    import torch.multiprocessing as mp

    L = mb.load()
    uid = (C.c_uint8 * 128)()
    assert L.mr_comm_unique_id(uid) == 0, L.mr_last_error(None).decode()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, 2, bytes(uid), q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    expect = np.arange(48 * 3, dtype=np.int32) * 3  # (rank 0: x1) + (rank 1: x2)
    for r in range(2):
        assert isinstance(outs[r], np.ndarray), outs[r]
        assert (outs[r] == expect).all()


def _rp_rank(rank, world, port, q):
    import os

    import torch
    import torch.distributed as dist

    import mcts_b200 as mb

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        with mb.GpuSearch(mb.Game.BREAKTHROUGH, mb.EpsilonGreedyMast(0.1), max_iterations=512, batch_leaves=32, num_rollouts_per_leaf=16,
                          max_playout_depth=200, seed=100 + rank, devices=[rank]) as s:
            action, visits, scores, local = mb.choose_action_root_parallel(s, mb.initial_state(mb.Game.BREAKTHROUGH),
                                                                           device=torch.device("cuda", rank), draw=5)
        q.put((rank, action, visits, scores, local.children.copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_root_parallel_search_over_nccl():
    """BASELINE config 4 in miniature: one searcher per GPU, root-child totals merged with one NCCL all-reduce; the merged
    totals are the sum of the two local trees' totals and both ranks choose the same move."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rp_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    v_sum = outs[0][4]["visits"].astype(np.int64) + outs[1][4]["visits"]
    assert (outs[0][4]["visits"] != outs[1][4]["visits"]).any()  # different seeds grew different trees
    for o in outs:
        assert (o[2] == v_sum).all()
        assert (o[3] == outs[0][4]["score"] + outs[1][4]["score"]).all()
    assert outs[0][1] == outs[1][1]
    # every playout below the root is counted once; the first batch of each tree plays from the unexpanded root itself
    assert 2 * (512 - 32) * 16 <= v_sum.sum() <= 2 * 512 * 16
