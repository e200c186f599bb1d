#!/usr/bin/env python3
"""Iteration-count statistics of full BP per epsilon (how long the stragglers of a 64*W-frame word run)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fl_scaling_sc_ldpc_b200 as eng

ap = argparse.ArgumentParser()
ap.add_argument("--n-words", type=int, default=8)
ap.add_argument("--M", type=int, default=10000)
ap.add_argument("--L", type=int, default=50)
ap.add_argument("--graphs", type=int, default=2)
a = ap.parse_args()
ens = eng.Ensemble(4, 8, a.L, a.M)
F = 64 * a.n_words
for eps in (0.46, 0.47, 0.48, 0.49):
    fb = eng.FrameBatch(ens, a.graphs, F, a.n_words).generate_graphs(11).generate_erasures(eps, 12)
    r = eng.decode_bp_full(fb, 0, True)
    it = r.iters.reshape(-1); ok = r.residual.reshape(-1) == 0
    q = lambda x, p: float(np.percentile(x, p)) if len(x) else None
    print(json.dumps(dict(eps=eps, fer=float(1 - ok.mean()), mean=float(it.mean()), p50=q(it, 50), p90=q(it, 90), p99=q(it, 99), max=int(it.max()),
                          mean_ok=float(it[ok].mean()) if ok.any() else None, max_ok=int(it[ok].max()) if ok.any() else None,
                          mean_fail=float(it[~ok].mean()) if (~ok).any() else None, max_fail=int(it[~ok].max()) if (~ok).any() else None,
                          launched=r.iters_launched, useful_frac=float(it.sum() / (r.iters_launched * it.size)),
                          chunk_useful=float(it.sum() / sum(it.reshape(-1, 128).max(axis=1) * 128)))))
