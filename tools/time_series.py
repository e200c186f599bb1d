#!/usr/bin/env python3
"""Per-iteration sweep times, active frames and swept work of one full-BP batch (where does a decode spend its time?)."""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--eps", type=float, default=0.49)
ap.add_argument("--n-words", type=int, default=16)
ap.add_argument("--graphs", type=int, default=1)
a = ap.parse_args()
ens = eng.Ensemble(4, 8, 50, 10000)
F = 64 * a.n_words
fb = eng.FrameBatch(ens, a.graphs, F, a.n_words).generate_graphs(11).generate_erasures(a.eps, 12)
lib = _lib.lib()
eng.decode_bp_full(fb, 0, True, collect=False); torch.cuda.synchronize()
_lib.check(lib.scldpc_profile_begin(1, 8000))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); r = eng.decode_bp_full(fb, 0, True); t1.record(); torch.cuda.synchronize()
cap = 8000; ns = ctypes.c_int(0); idx = (ctypes.c_int * cap)(); cn = (ctypes.c_float * cap)(); vn = (ctypes.c_float * cap)()
_lib.check(lib.scldpc_profile_end(ctypes.byref(ns), idx, cn, vn, cap))
n = ns.value
cn = np.array(cn[:n]) * 1e3; vn = np.array(vn[:n]) * 1e3
it = r.iters.reshape(-1)
active = np.array([(it > t).sum() for t in range(n)])
print(json.dumps(dict(eps=a.eps, launched=r.iters_launched, wall_ms=t0.elapsed_time(t1), sum_kernel_ms=float((cn.sum() + vn.sum()) / 1e3),
                      mean_iters=float(it.mean()), useful_frac=float(it.sum() / (n * it.size)))))
for lo in range(0, n, max(1, n // 25)):
    hi = min(n, lo + max(1, n // 25))
    print(f"iter {lo:5d}-{hi:5d}  active {active[lo:hi].mean():7.1f}  cn {cn[lo:hi].mean():7.1f} us  vn {vn[lo:hi].mean():7.1f} us")
