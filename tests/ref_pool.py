"""Process pool that decodes frames with the UNMODIFIED reference (oracle/_ref/*.so through oracle/ref_driver.py).

Test infrastructure: the workers are spawned (not forked: the parent holds a CUDA context) and import only numpy and the
oracle package.  One job = one (graph, channel realisation) pair decoded by ``decodeBP`` (BP_TRAJ.c:901, whose rows give
the executed iterations) or ``decodeBP_SW`` (BP_SW.c:628) at the size the shared object was compiled for."""
from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np


def _job(args):
    kind, dv, dc, L, defM, vn_cn, chan, prm = args
    from oracle import ref_driver as rd
    r = rd.get("sw" if kind == "sw" else "traj", dv, dc, L, defM)
    r.set_graph_fast(vn_cn)
    r.set_channel(chan)
    if kind == "sw":
        o = r.decode_bp_sw(prm["W"], prm["max_it"], prm.get("init_it", 0))
        o["erased"] = np.packbits(o["erased"])
        return o
    o = r.decode_bp(prm["max_it"], prm.get("is_term", 1))
    o["iters"] = len(o["rows"])
    o["erased"] = np.packbits(o["erased"])
    if not prm.get("want_rows", False):
        o["rows"] = None
    return o


def available(kind, dv, dc, L, defM) -> bool:
    from oracle import build_ref
    return os.path.isfile(build_ref.so_name("sw" if kind == "sw" else "traj", dv, dc, L, defM))


def run(jobs, processes=None):
    """jobs: list of (kind, dv, dc, L, defM, vn_cn int32[n][dv], chan uint8[n], params) -> list of result dicts"""
    processes = processes or min(len(jobs), os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(processes) as pool:
        return pool.map(_job, jobs, chunksize=1)
