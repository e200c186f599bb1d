#!/usr/bin/env python3
"""Peeling throughput against the number of frames in flight (SCLDPC_PEEL_SLOTS caps the warps of the launch): is the M = 10000
decoder bound by latency (throughput grows with the frames in flight) or by a shared resource (flat)?"""
import argparse, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=10000)
ap.add_argument("--graphs", type=int, default=8)
ap.add_argument("--frames", type=int, default=256, help="frames per graph")
ap.add_argument("--slots", default="256,512,1024,2048,4096,9472")
ap.add_argument("--no-r1", action="store_true")
ap.add_argument("--one", type=int, default=0)
a = ap.parse_args()
if not a.one:
    for s in a.slots.split(","):
        env = dict(os.environ, SCLDPC_PEEL_SLOTS=s)
        cmd = [sys.executable, __file__, "--one", s, "--M", str(a.M), "--graphs", str(a.graphs), "--frames", str(max(a.frames, int(s) // a.graphs))]
        subprocess.run(cmd + (["--no-r1"] if a.no_r1 else []), env=env)
    sys.exit(0)
import torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx
l, r, L, e = 4, 8, 50, 0.48
ens = eng.Ensemble(l, r, L, a.M)
cns, npos, total_size, steps = pdx._peel_geometry(e, l, r, L, a.M, False)
fb = eng.FrameBatch(ens, a.graphs, a.frames).generate_graphs(1).generate_erasures(e, 2)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, 3, 0, want_r1=not a.no_r1)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
peeled = int(rec.sum().item())
print(json.dumps(dict(M=a.M, slots=a.one, frames=a.graphs * a.frames, ms=round(ms, 1), frames_per_s=round(a.graphs * a.frames / ms * 1e3, 1),
                      peel_steps_per_s=peeled / ms * 1e3, us_per_step_per_frame=round(min(a.one, a.graphs * a.frames) * ms * 1e3 / peeled, 2), r1=not a.no_r1)), flush=True)
