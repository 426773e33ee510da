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
