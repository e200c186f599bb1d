"""N > 1 host logic on CPU: world_size 2 over gloo (127.0.0.1) -- graph-id sharding, the counter all-reduce and the
deterministic sequential early stop."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from fl_scaling_sc_ldpc_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = D.shard_graph_ids(100, 11)
    # per-rank "results": frame f of graph g fails iff (g*7 + f) % 5 == 0 -- a function of global ids only
    F = 4
    frame_ids = np.array([g * F + f for g in ids for f in range(F)])
    fails = np.array([int((g * 7 + f) % 5 == 0) for g in ids for f in range(F)])
    tot = D.allreduce_counters([len(frame_ids), fails.sum(), frame_ids.sum()])
    cut = D.sequential_stop_index(fails, frame_ids, 3)
    tmax = D.allreduce_max(1.0 + rank)
    q.put((rank, ids, tot.tolist(), cut, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_reduction_and_early_stop():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids0, ids1 = res[0][1], res[1][1]
    assert sorted(ids0 + ids1) == list(range(100, 111)) and not set(ids0) & set(ids1)
    # single-process ground truth
    F = 4
    all_ids = np.array([g * F + f for g in range(100, 111) for f in range(F)])
    all_f = np.array([int((g * 7 + f) % 5 == 0) for g in range(100, 111) for f in range(F)])
    expect_tot = [len(all_ids), int(all_f.sum()), int(all_ids.sum())]
    expect_cut = int(np.flatnonzero(np.cumsum(all_f[np.argsort(all_ids)]) >= 3)[0]) + 1
    for r in res:
        assert r[2] == expect_tot and r[3] == expect_cut and r[4] == 2.0


def test_single_process_fallbacks():
    from fl_scaling_sc_ldpc_b200 import dist as D
    assert D.world() == (0, 1)
    assert D.shard_graph_ids(5, 4) == [5, 6, 7, 8]
    assert D.allreduce_counters([1, 2]).tolist() == [1, 2] and D.allreduce_max(3.5) == 3.5
    assert D.sequential_stop_index(np.array([0, 1, 0, 1, 1]), np.arange(5), 2) == 4
    assert D.sequential_stop_index(np.array([0, 1]), np.arange(2), 2) == -1
