#!/usr/bin/env python3
"""simulate_sc_ldpc at eps = 0.48, M = 1000 (FER ~ 0.8: almost every frame needs the stopping-set
bookkeeping) with the device-side records (scldpc_bp_stopping_sets) against the round-1 way -- one device-to-host copy and one
SciPy connected-components call per failed frame -- on the same decoded batch."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

e, l, r, L, M = 0.48, 4, 8, 50, 1000
out = {}
pdx.set_seed(5)
for fpg, G in ((128, 8), (1024, 2)):
    n = 4096
    pdx.simulate_sc_ldpc(e, l, r, L, M, True, False, True, False, num_repeats=fpg * G, max_fuckups=10 ** 9, frames_per_graph=fpg, graphs_per_batch=G, progress=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = pdx.simulate_sc_ldpc(e, l, r, L, M, True, False, True, False, num_repeats=n, max_fuckups=10 ** 9, frames_per_graph=fpg, graphs_per_batch=G, progress=False)
    dt = time.perf_counter() - t0
    out[f"simulate_sc_ldpc fpg={fpg} G={G}"] = {"frames": n, "seconds": dt, "frames_per_s": n / dt, "FER": res[0], "FER_exp": res[1], "PLR": res[2]}
# the same bookkeeping the round-1 way on one batch of 1024 frames
ens = eng.Ensemble(l, r, L, M)
fb = eng.FrameBatch(ens, 8, 128).generate_graphs(5).generate_erasures(e, 6)
resd = eng.decode_bp_full(fb, 0, True)
counted = np.ones(L, bool)
torch.cuda.synchronize()
t0 = time.perf_counter()
rec_dev = pdx.stopping_set_records(fb, resd.erased_words, counted)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
words = resd.erased_words
rec_host = np.zeros((8, 128, 4), np.int64)
for g in range(8):
    tr_g = None
    for f in np.flatnonzero(resd.residual[g] > 0):
        f = int(f)
        bits = ((words[g, :, f >> 6] >> (f & 63)) & 1).bool().cpu()
        lost = torch.nonzero(bits).reshape(-1).numpy()
        if len(lost):
            if tr_g is None:
                tr_g = fb.vn_cn[g].cpu().numpy()
            rec_host[g, f] = pdx.account_lost(lost, tr_g, M)
t_host = time.perf_counter() - t0
assert (rec_dev[:, :128] == rec_host).all()
out["bookkeeping of one 1024-frame batch"] = {"failed_frames": int((resd.residual > 0).sum()), "device_ms": 1e3 * t_dev, "round1_host_loop_ms": 1e3 * t_host,
                                            "speedup": t_host / t_dev}
print(json.dumps(out, indent=1))
