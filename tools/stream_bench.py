#!/usr/bin/env python3
"""Stream-mode micro-benchmark: one graph, lane recycling; prints per-sweep times and the lane utilisation."""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--eps", type=float, default=0.48)
ap.add_argument("--n-words", type=int, default=16)
ap.add_argument("--frames", type=int, default=3072)
ap.add_argument("--graphs", type=int, default=1)
ap.add_argument("--harvest", type=int, default=16)
a = ap.parse_args()
ens = eng.Ensemble(4, 8, 50, 10000)
lanes = 64 * a.n_words
fb = eng.FrameBatch(ens, a.graphs, lanes, a.n_words).generate_graphs(11)
lib = _lib.lib()
eng.decode_bp_stream(fb, lanes, a.eps, 12, collect=False); torch.cuda.synchronize()
_lib.check(lib.scldpc_profile_begin(4, 8000))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); r = eng.decode_bp_stream(fb, a.frames, a.eps, 12, harvest_every=a.harvest); t1.record(); torch.cuda.synchronize()
cap = 8000; ns = ctypes.c_int(0); idx = (ctypes.c_int * cap)(); cn = (ctypes.c_float * cap)(); vn = (ctypes.c_float * cap)()
_lib.check(lib.scldpc_profile_end(ctypes.byref(ns), idx, cn, vn, cap))
n = ns.value
cn = np.array(cn[:n]) * 1e3; vn = np.array(vn[:n]) * 1e3
fi = int(r.iters.astype(np.int64).sum())
wall = t0.elapsed_time(t1)
print(json.dumps(dict(eps=a.eps, frames=a.frames * a.graphs, launched=r.iters_launched, wall_ms=wall, cn_us=float(np.median(cn)), vn_us=float(np.median(vn)),
                      lane_util=fi / (r.iters_launched * lanes * a.graphs), frame_iters_per_s=fi / wall * 1e3,
                      alg_GBs=fi * 1.0625e6 / wall * 1e3 / 1e9, mean_iters=fi / (a.frames * a.graphs))))
