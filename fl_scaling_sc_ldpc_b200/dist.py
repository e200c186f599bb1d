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


def init_from_env() -> tuple[int, int]:
    """Joins the job ``torchrun`` (or any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*) started: one
    process per GPU, NCCL when CUDA is there, gloo otherwise.  A plain ``python ...`` launch stays single-process."""
    import os
    if dist.is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return world()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    backend = os.environ.get("SCLDPC_DIST_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
    if torch.cuda.is_available():
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(backend)      # gloo: ranks may share a GPU (tests on a one-GPU box); collectives go through the host
    return world()


def shutdown():
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def shard_graph_ids(first: int, count: int, rank: int | None = None, world_size: int | None = None):
    """Graph ids [first, first+count) are dealt round-robin in blocks: rank r gets ids first + r, first + r + N, ...
    Ids are global, so the union of all ranks' realisations does not depend on N."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(first + rank, first + count, world_size))


def allreduce_counters(values, device=None) -> np.ndarray:
    """Sum an int64 counter vector over all ranks (the job's only collective)."""
    t = torch.as_tensor(np.asarray(values, dtype=np.int64))
    if world()[1] > 1:
        t = t.to(device if device is not None else _comm_device())    # NCCL reduces device tensors, gloo host tensors
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def broadcast_int(value: int) -> int:
    """rank 0's value on every rank (e.g. a time-derived default seed: every rank must draw from the same stream)"""
    if world()[1] == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_comm_device())
    dist.broadcast(t, src=0)
    return int(t.item())


def allreduce_max(value: float, device=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64)
    if world()[1] > 1:
        t = t.to(device if device is not None else _comm_device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _comm_device():
    """device the collectives run on: the current GPU under NCCL, the CPU under gloo"""
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allgather_rows(arr: np.ndarray) -> np.ndarray:
    """All ranks contribute an array of the same shape; returns them stacked as [world][...] (identical on every rank)."""
    arr = np.ascontiguousarray(arr)
    n = world()[1]
    if n == 1:
        return arr[None]
    t = torch.as_tensor(arr).to(_comm_device())
    out = [torch.empty_like(t) for _ in range(n)]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def allgather_tensor(t: torch.Tensor) -> torch.Tensor:
    """Same for a tensor that stays where it is (device tensors under NCCL): returns [world][...]."""
    n = world()[1]
    if n == 1:
        return t[None]
    home = t.device
    t = t.contiguous().to(_comm_device())
    out = [torch.empty_like(t) for _ in range(n)]
    dist.all_gather(out, t)
    return torch.stack(out).to(home)


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
