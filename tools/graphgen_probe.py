#!/usr/bin/env python3
"""Where does the time of FrameBatch.generate_graphs go inside a bench-like loop?  Host wall time of the two C-ABI calls
(scldpc_graph_generate, scldpc_graph_build_tables) per step, after a stream decode, with and without an nvidia-smi poller."""
import ctypes, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import _lib
from fl_scaling_sc_ldpc_b200.engine import _stream

ens = eng.Ensemble(4, 8, 50, 10000)
fb = eng.FrameBatch(ens, 4, 1024, 16)
L = _lib.lib()
eps = [0.46, 0.47, 0.48, 0.49]
out = {}
for poll in (None, "1000", "100"):
    proc = None
    if poll:
        proc = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,power.draw", "--format=csv,noheader", "-lms", poll],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(1.5)
    rows = []
    for s in range(6):
        gid = 1000 * (1 if poll is None else int(poll)) + 4 * s
        t0 = time.perf_counter()
        nbytes = L.scldpc_graph_generate_scratch_bytes(ctypes.byref(fb.dims), 0)
        if fb._keys is None or fb._keys.numel() * 8 < nbytes:
            fb._keys = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=fb.device)
        _lib.check(L.scldpc_graph_generate(ctypes.byref(fb.dims), ctypes.c_void_p(fb.vn_cn.data_ptr()), ctypes.c_void_p(fb._keys.data_ptr()),
                                           ctypes.c_uint64(7), ctypes.c_uint64(gid), 0, _stream()))
        t1 = time.perf_counter()
        fb._build_tables()
        t2 = time.perf_counter()
        res, launched = eng.decode_bp_stream(fb, 4096, eps, 8, first_graph_id=gid, collect=False)
        t3 = time.perf_counter()
        x = int(res[0].sum().item())
        t4 = time.perf_counter()
        rows.append([round(1e3 * (b - a), 2) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))])
    if proc:
        proc.terminate(); proc.wait()
    out[f"poll_ms={poll}"] = {"columns": ["graph_generate_ms", "build_tables_ms", "decode_stream_ms", "result_sum_ms"], "steps": rows}
print(json.dumps(out, indent=1))
