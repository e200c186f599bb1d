#!/usr/bin/env python3
"""Sliding-window decoder throughput (BASELINE config 4: (4,8), L=100, M=10000, W=3..10, per-window cap)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng

ap = argparse.ArgumentParser()
ap.add_argument("--W", type=int, default=10)
ap.add_argument("--cap", type=int, default=8)
ap.add_argument("--init", type=int, default=60)
ap.add_argument("--eps", type=float, default=0.45)
ap.add_argument("--graphs", type=int, default=4)
ap.add_argument("--n-words", type=int, default=16)
ap.add_argument("--L", type=int, default=100)
ap.add_argument("--M", type=int, default=10000)
ap.add_argument("--nonterm", action="store_true")
a = ap.parse_args()
ens = eng.Ensemble(4, 8, a.L, a.M)
F = 64 * a.n_words
fb = eng.FrameBatch(ens, a.graphs, F, a.n_words).generate_graphs(1).generate_erasures(a.eps, 2)
eng.decode_bp_window(fb, a.W, a.cap, a.init, True, not a.nonterm, collect=False); torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); r = eng.decode_bp_window(fb, a.W, a.cap, a.init, True, not a.nonterm); t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
frames = a.graphs * F
print(json.dumps(dict(W=a.W, cap=a.cap, init=a.init, eps=a.eps, L=a.L, M=a.M, frames=frames, ms=ms, frames_per_s=frames / ms * 1e3,
                      edge_updates_per_s=r.edge_updates / ms * 1e3, mean_iters_per_frame=float(r.iters.mean()),
                      fer=float((r.residual > 0).mean()), alg_GBs=r.edge_updates * 0.2656 / ms * 1e3 / 1e9)))
