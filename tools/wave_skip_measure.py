#!/usr/bin/env python3
"""which fraction of (CN position, 128-lane chunk) pairs could the node-state sweep skip?

Needs the instrumented build (tools/build_variants.sh 16 -> libscldpc_v16.so, selected through SCLDPC_LIB): the last block of
every iteration counts the pairs whose dv-1 neighbouring CN positions saw a resolution in that iteration (the ones the next
sweep would have to visit) among the pairs of chunks that hold an active frame.  Two arming policies per eps:
  synchronous   frames_per_graph = lanes: every lane starts at iteration 0 and nobody is re-armed -- the best case a
                chunk-synchronous re-arming policy could reach (all 128 frames of a chunk in phase)
  recycled      frames_per_graph = 8 x lanes: lanes are re-armed as they finish (the bench's policy)"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import _lib

ens = eng.Ensemble(4, 8, 50, 10000)
lanes, nw = 1024, 16
lib = _lib.lib()
out = {}
for eps in (0.46, 0.48, 0.49):
    for name, B in (("synchronous", lanes), ("recycled", 8 * lanes)):
        fb = eng.FrameBatch(ens, 1, lanes, nw).generate_graphs(11)
        r = eng.decode_bp_stream(fb, B, eps, 12)
        st = (ctypes.c_longlong * 2)()
        _lib.check(lib.scldpc_bp_sweep_stats(ctypes.byref(fb.dims), 32, ctypes.c_void_p(fb._ws.data_ptr()), st))
        util = float(r.iters.astype(np.int64).sum()) / (r.iters_launched * lanes)
        out[f"eps={eps} {name}"] = {"pairs_needed": int(st[0]), "pairs_total": int(st[1]), "needed_fraction": st[0] / max(1, st[1]),
                                    "lane_utilisation": util, "iterations_launched": int(r.iters_launched),
                                    "mean_iterations_per_frame": float(r.iters.mean())}
        del fb
print(json.dumps(out, indent=1))
