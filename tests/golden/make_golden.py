#!/usr/bin/env python3
"""Generates tests/golden/bp_golden_*.npz from the UNMODIFIED reference C code (oracle/_ref, compiled in place from
/root/reference by oracle/build_ref.py).  Runs in the build container only; the fixtures it writes are committed.

Each fixture holds, for one ensemble: graphs drawn by the reference's generate_code (glibc random(), seeded), channel
realisations drawn by channel_doped, and the outputs of the reference decoders on every (graph, channel) pair:
  decodeBP     (BP_TRAJ.c:901)  terminated and truncated, several iteration caps -> residual, blocks_err, erasures_exp,
                                blocks_err_exp, VNerased, trajectory rows
  decodeBP_SW  (BP_SW.c:628 square; BP_FULL.c:627 classical), several (W, cap, init)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [  # dv, dc, L, Def_M, epsilons, seed
    (4, 8, 10, 25, [0.40, 0.46, 0.50], 20260101),
    (3, 6, 10, 24, [0.38, 0.44], 20260102),
    (5, 10, 12, 20, [0.42, 0.49], 20260103),
    (4, 8, 6, 8, [0.30, 0.45, 0.60], 20260104),
]
BP_CAPS = [100000, 5, 1]
SW_CFG = [(3, 100000, 0), (4, 4, 12), (5, 2, 0), (2, 1, 1)]
N_GRAPHS, N_FRAMES = 3, 4


def main():
    for dv, dc, L, defM, epss, seed in CASES:
        rt, rs, rf = (rd.get(v, dv, dc, L, defM) for v in ("traj", "sw", "full"))
        rt.srandom(seed)
        rt.reset_perm()
        graphs, chans, eps_of = [], [], []
        bp = {}
        sw = {}
        for g in range(N_GRAPHS):
            vn_cn = rt.generate_code()
            graphs.append(vn_cn)
            rs.set_graph(vn_cn)
            rf.set_graph(vn_cn)
            for f in range(N_FRAMES):
                eps = epss[(g * N_FRAMES + f) % len(epss)]
                doped = [L // 2] if (f == N_FRAMES - 1) else []
                ch = rt.channel_doped(eps, doped)
                chans.append(ch)
                eps_of.append(eps)
                rs.set_channel(ch)
                rf.set_channel(ch)
                for is_term in (1, 0):
                    for cap in BP_CAPS:
                        o = rt.decode_bp(cap, is_term)
                        key = f"bp_t{is_term}_c{cap}"
                        d = bp.setdefault(key, dict(stats=[], erased=[], rows=[]))
                        d["stats"].append([len(o["rows"]), o["residual"], o["blocks_err"], o["erasures_exp"], o["blocks_err_exp"]])
                        d["erased"].append(o["erased"])
                        rows = np.full((64, 3), -1, np.int32)
                        k = min(64, len(o["rows"]))
                        rows[:k] = o["rows"][:k, 1:]
                        d["rows"].append(rows)
                for (W, cap, init) in SW_CFG:
                    for square, lib in ((1, rs), (0, rf)):
                        o = lib.decode_bp_sw(W, cap, init)
                        key = f"sw_s{square}_W{W}_c{cap}_i{init}"
                        d = sw.setdefault(key, dict(stats=[], erased=[]))
                        d["stats"].append([o["residual"], o["erasures_p1"], o["blocks_err"], o["erasures_exp"], o["blocks_err_exp"]])
                        d["erased"].append(o["erased"])
        arrays = dict(dims=np.array([dv, dc, L, 2 * defM, defM, N_GRAPHS, N_FRAMES], np.int32), seed=np.int64(seed),
                      vn_cn=np.array(graphs, np.int32),
                      chan=np.array(chans, np.uint8).reshape(N_GRAPHS, N_FRAMES, -1), eps=np.array(eps_of))
        for k, d in {**bp, **sw}.items():
            arrays[k + "_stats"] = np.array(d["stats"], np.int32).reshape(N_GRAPHS, N_FRAMES, -1)
            arrays[k + "_erased"] = np.packbits(np.array(d["erased"], np.uint8).reshape(N_GRAPHS, N_FRAMES, -1), axis=2)
            if "rows" in d:
                arrays[k + "_rows"] = np.array(d["rows"], np.int32).reshape(N_GRAPHS, N_FRAMES, 64, 3)
        path = os.path.join(OUT, f"bp_golden_{dv}_{dc}_L{L}_M{2 * defM}.npz")
        np.savez_compressed(path, **arrays)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
