#!/usr/bin/env python3
"""Two batches in flight on two CUDA streams (two host threads, one scldpc_bp_stream call each) against one batch after the other:
does the tail of one batch's streams hide behind the next batch?  Same work either way: STEPS steps of 4 new graphs x 16384 frames."""
import argparse, json, os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--frames", type=int, default=16384)
ap.add_argument("--pipes", type=int, nargs="*", default=[1, 2, 1, 2])
a = ap.parse_args()
ens = eng.Ensemble(4, 8, 50, 10000)
eps = [0.46, 0.47, 0.48, 0.49]
G, lanes = 4, 1024
E2 = 2.0 * ens.E

def run(pipes, first_gid):
    fbs = [eng.FrameBatch(ens, G, lanes, 16) for _ in range(pipes)]
    streams = [torch.cuda.Stream() for _ in range(pipes)]
    tot = [0] * pipes
    def worker(t):
        with torch.cuda.stream(streams[t]):
            for s in range(t, a.steps, pipes):
                gid = first_gid + s * G
                fbs[t].generate_graphs(seed=1, first_graph_id=gid)
                res, _ = eng.decode_bp_stream(fbs[t], a.frames, eps, 2, first_graph_id=gid, collect=False)
                tot[t] += int(res[0].to(torch.int64).sum().item())
            streams[t].synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    th = [threading.Thread(target=worker, args=(t,)) for t in range(pipes)]
    [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(pipes=pipes, steps=a.steps, ms=ms, edge_updates_per_s=sum(tot) * E2 / ms * 1e3)

run(1, 0)
for i, p in enumerate(a.pipes):
    print(json.dumps(run(p, 1000 * (i + 1))), flush=True)
