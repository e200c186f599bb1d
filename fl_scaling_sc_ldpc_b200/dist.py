"""Multi-GPU plumbing: frames shard trivially (one process per GPU, independent Philox streams); only final counters
are combined.  Works on any ``torch.distributed`` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference's parallel model is N independent OS processes with different ``INDEX`` whose output files are summed
offline with awk / pickle sums (NB:565, NB:1195, notebook cell 23); here the sum is one all-reduce.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_graph_ids(first: int, count: int, rank: int | None = None, world_size: int | None = None):
    """Graph ids [first, first+count) are dealt round-robin in blocks: rank r gets ids first + r, first + r + N, ...
    Ids are global, so the union of all ranks' realisations does not depend on N."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(first + rank, first + count, world_size))


def allreduce_counters(values, device=None) -> np.ndarray:
    """Sum an int64 counter vector over all ranks (the job's only collective)."""
    t = torch.as_tensor(np.asarray(values, dtype=np.int64))
    if device is not None:
        t = t.to(device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(value: float, device=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sequential_stop_index(fail_flags_local: np.ndarray, frame_ids_local: np.ndarray, threshold: int, device=None) -> int:
    """Deterministic early stop across ranks (``frame_err >= numero_frame_err``, BP_FULL.c:440; ``num_fuckups >=
    max_fuckups``, PD.py:698): given this rank's per-frame failure flags and global frame ids of one round, returns the
    number of frames (in global frame order) after which the threshold-th failure has occurred, or -1 if it has not.
    Every rank gets the same answer."""
    rank, n = world()
    flags = np.asarray(fail_flags_local, dtype=np.int64)
    ids = np.asarray(frame_ids_local, dtype=np.int64)
    if n > 1:
        sizes = [None] * n
        dist.all_gather_object(sizes, (ids.tolist(), flags.tolist()))
        ids = np.concatenate([np.asarray(s[0], np.int64) for s in sizes])
        flags = np.concatenate([np.asarray(s[1], np.int64) for s in sizes])
    order = np.argsort(ids, kind="stable")
    cum = np.cumsum(flags[order])
    hit = np.flatnonzero(cum >= threshold)
    return int(hit[0]) + 1 if len(hit) else -1
