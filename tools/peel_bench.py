#!/usr/bin/env python3
"""Peeling-trajectory throughput (BASELINE config 1: (4,8), L=50, M=1000, non-terminated, eps=0.48; and M=10000)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1000)
ap.add_argument("--frames", type=int, default=4096)
ap.add_argument("--eps", type=float, default=0.48)
ap.add_argument("--no-r1", action="store_true")
a = ap.parse_args()
l, r, L = 4, 8, 50
ens = eng.Ensemble(l, r, L, a.M)
cns, npos, total_size, steps = pdx._peel_geometry(a.eps, l, r, L, a.M, False)
fb = eng.FrameBatch(ens, a.frames, 1, 2).generate_graphs(1).generate_erasures(a.eps, 2)
torch.cuda.synchronize()
for rep in range(2):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, 3, 0, want_r1=not a.no_r1)
    t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
peeled = int(rec.sum().item())
print(json.dumps(dict(M=a.M, frames=a.frames, steps_per_frame=steps, ms=ms, frames_per_s=a.frames / ms * 1e3, peel_steps_per_s=peeled / ms * 1e3,
                      mean_plr=float(((ner - rec).float() / (L * a.M)).mean().item()), r1=not a.no_r1)))
