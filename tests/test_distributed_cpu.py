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


# ---- the drop-in drivers' rounds: world_size batches per round, all-gathered per-frame records, sequential replay ----------------
def _fake_records(G, fpg, gid0):
    """deterministic per-frame records from global ids only: (num_lost, big, lost_exp, blocks_exp)"""
    rec = np.zeros((G * fpg, 4), np.int64)
    for g in range(G):
        for f in range(fpg):
            k = (gid0 + g) * fpg + f
            if (k * 2654435761 >> 7) % 9 == 0:
                rec[g * fpg + f] = (3 + k % 5, int(k % 3 != 0), (2 + k % 4) * int(k % 3 != 0), 1 + k % 2 if k % 3 else 0)
    return rec


def _driver_args(num_repeats, max_fuckups):
    return dict(e=0.45, l=4, r=8, L=12, M=64, is_terminated=True, is_protograph=False, is_bounded=True, is_tail_biting=False,
                num_repeats=num_repeats, max_fuckups=max_fuckups, frames_per_graph=4, graphs_per_batch=3, first_frame=24,
                progress=False, _records_fn=_fake_records)


def _driver_worker(rank, world, port, q):
    import torch.distributed as dist
    from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for nr, mf in ((1000, 7), (50, 10 ** 6), (61, 3)):
        t = pdx.simulate_sc_ldpc(**_driver_args(nr, mf))
        out.append([float(x) for i, x in enumerate(t) if i not in (8, 9)])
    q.put((rank, out))
    dist.destroy_process_group()


def test_simulate_sc_ldpc_rounds_match_single_process():
    from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx
    expect = []
    for nr, mf in ((1000, 7), (50, 10 ** 6), (61, 3)):
        t = pdx.simulate_sc_ldpc(**_driver_args(nr, mf))
        expect.append([float(x) for i, x in enumerate(t) if i not in (8, 9)])
    assert expect[0][4] <= 7 and expect[1][5] == 50                     # the cut fired / the frame budget was used up
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_driver_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, out in res:
        assert out == expect
