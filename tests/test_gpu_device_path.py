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
