"""Multi-GPU plumbing of the rollout path: how leaf batches are sharded over ranks and how the small
per-root-child statistics are merged.

Playouts are independent given (leaf state, Philox key), so the data path needs no collective: each rank
runs its own leaf batches.  What is exchanged, once per search-iteration batch, is what the reference's
root-parallel search merges on the host (search/parallel.rs:94-115): per root child `(visits, score[2])`,
summed over searchers, final move = the child with most merged visits -- plus, for MAST, the batch's table
delta, so that every rank plays the next batch from the same table (backprop.rs:857-870).

One process per GPU; torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) share of `n_items` for `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    return n_items * rank // world, n_items * (rank + 1) // world


def leaf_id_base(rank: int, world: int, step: int, leaves_per_rank: int) -> int:
    """Philox leaf-id base of a rank's batch at `step`: distinct (rank, step) pairs never share a stream."""
    return ((step * world + rank) * leaves_per_rank) & 0xFFFFFFFF


def merge_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place element-wise sum over ranks (a no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def merge_root_stats(visits: np.ndarray, scores: np.ndarray, device=None, group=None) -> tuple[np.ndarray, np.ndarray]:
    """Sums per-root-child visit counts [C] and per-player score sums [C, 2] over ranks.

    Children must be in the same (root action list) order on every rank, which generate_actions guarantees."""
    v = np.ascontiguousarray(visits, dtype=np.int64).reshape(-1)
    s = np.ascontiguousarray(scores, dtype=np.int64).reshape(len(v), -1)
    buf = torch.from_numpy(np.concatenate([v, s.reshape(-1)]))
    if device is not None:
        buf = buf.to(device)
    merge_sum_(buf, group)
    out = buf.cpu().numpy()
    return out[: len(v)].copy(), out[len(v):].reshape(s.shape).copy()


def merge_mast_delta(delta: np.ndarray, device=None, group=None) -> np.ndarray:
    """Sums a [2][num_actions] (visits, score) MAST delta over ranks."""
    flat = np.ascontiguousarray(delta).view(np.int32).reshape(-1).astype(np.int64)
    buf = torch.from_numpy(flat)
    if device is not None:
        buf = buf.to(device)
    merge_sum_(buf, group)
    out = np.zeros(delta.shape, dtype=delta.dtype)
    out.view(np.int32).reshape(-1)[:] = buf.cpu().numpy().astype(np.int32)
    return out


def merge_table_i32(table: np.ndarray, device=None, group=None) -> np.ndarray:
    """Sums a (visits, score) table -- any shape, e.g. the [2][num_slots] MAST delta of MR_TABLES_IN_SLOTS (3 KB for
    Breakthrough) -- over ranks as int32, the element type it has on the device: host -> device -> one all-reduce
    (NCCL over NVLink on GPUs, gloo in the CPU tests) -> host.  Returns a new array of the input's dtype and shape."""
    buf = torch.from_numpy(np.ascontiguousarray(table).view(np.int32).reshape(-1).copy())
    if device is not None:
        buf = buf.to(device)
    merge_sum_(buf, group)
    out = np.empty_like(np.ascontiguousarray(table))
    out.view(np.int32).reshape(-1)[:] = buf.cpu().numpy()
    return out


class PendingMerge:
    """An all-reduce of a (visits, score) table in flight: `result()` waits for it and returns the merged table."""

    def __init__(self, buf: torch.Tensor, work, like: np.ndarray):
        self._buf, self._work, self._like = buf, work, like

    def result(self) -> np.ndarray:
        if self._work is not None:
            self._work.wait()
        out = np.empty_like(self._like)
        out.view(np.int32).reshape(-1)[:] = self._buf.cpu().numpy()
        return out


def merge_table_i32_async(table: np.ndarray, device=None, group=None) -> PendingMerge:
    """`merge_table_i32` without the wait: the host copies the table to the device, enqueues the all-reduce and goes on
    (gather the next batch, submit it); `result()` collects the merged table one step later.  The table a batch plays
    from is a batch stale by construction, so nothing needs the merged table any sooner -- and the ranks stop running
    in lock-step through a blocking collective every step."""
    like = np.ascontiguousarray(table)
    buf = torch.from_numpy(like.view(np.int32).reshape(-1).copy())
    if device is not None:
        buf = buf.to(device, non_blocking=True)
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return PendingMerge(buf, work, like)


PRIMES = (14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991, 53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693)


def final_action_index(merged_visits: np.ndarray, r: int) -> int:
    """random_best over merged visits (search/parallel.rs:111-115, util.rs:135-163): `r` is the combined
    draw in [0, 16 n); ties go to the first maximum in visiting order."""
    n = len(merged_visits)
    i, stride = r // 16, PRIMES[r % 16]
    best, best_i = None, i
    for _ in range(n):
        if best is None or merged_visits[i] > best:
            best, best_i = merged_visits[i], i
        i = (i + stride) % n
    return best_i
