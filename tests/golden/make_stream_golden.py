#!/usr/bin/env python3
"""Generates tests/golden/stream_golden.npz from the reference's streaming / circular-buffer decoder (BP_FULL.c built
with CIRCULAR defined, oracle/build_ref.py variant "circ"): main_streaming's loop is driven step by step through
ctypes (generate_stream_pos, initialize_messages_circular, decodeBP_SW_circular) and the chain it generated is recorded
unrolled (absolute CN ids), together with the running counters after every decode step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [  # name, dv, dc, ring L, Def_M, W, eps, doped positions, decode steps, srandom seed
    ("c0", 4, 8, 16, 8, 5, 0.42, [], 70, 11),
    ("c1", 4, 8, 24, 16, 8, 0.55, [5, 7, 9], 90, 12),
    ("c2", 3, 6, 16, 12, 6, 0.40, [6], 60, 13),
    ("c3", 4, 8, 24, 16, 10, 0.47, [], 50, 14),
    # long undoped run: windows fail and the erasures they leave behind keep propagating (several pieces of ChainPieces)
    ("c4", 4, 8, 24, 16, 8, 0.50, [], 220, 15),
]


def main():
    arrays = {}
    for name, dv, dc, L, defM, W, eps, doped, steps, seed in CASES:
        r = rd.get("circ", dv, dc, L, defM)
        r.srandom(seed)
        r.reset_perm()
        o = r.stream_run(steps, W, eps, doped)
        arrays[name + "_params"] = np.array([dv, dc, L, defM, W, steps], np.int32)
        arrays[name + "_eps"] = np.float64(eps)
        arrays[name + "_doped"] = np.array(doped, np.int32)
        arrays[name + "_vn_cn"] = o["vn_cn"]
        arrays[name + "_chan"] = o["chan"]
        arrays[name + "_steps"] = o["steps"]
    path = os.path.join(OUT, "stream_golden.npz")
    np.savez_compressed(path, **arrays)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
