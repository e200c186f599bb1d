#!/usr/bin/env python3
"""Throughput of the continuous streaming decoder (ChainPieces: unbounded chains decoded piece by piece with the decoder state
carried over; the reference's decodeBP_SW_circular / main_streaming, BP_FULL.c:1403-1500, 1934-2054): blocks (positions) decided
per second over all chains, graph and channel generation of every piece included."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fl_scaling_sc_ldpc_b200 import streaming

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1000, help="VNs per position (the reference's Def_M = 500 CNs per position)")
ap.add_argument("--W", type=int, default=10)
ap.add_argument("--eps", type=float, default=0.47)
ap.add_argument("--graphs", type=int, default=4)
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--piece", type=int, default=500)
ap.add_argument("--pieces", type=int, default=4)
ap.add_argument("--doped", type=int, nargs="*", default=[])
a = ap.parse_args()
ch = streaming.ChainPieces(4, 8, a.M, a.W, a.eps, a.doped, a.graphs, a.frames, 0x5C1D9C, piece=a.piece)
ch.next_piece(); torch.cuda.synchronize()
t0 = time.perf_counter(); err = 0
for _ in range(a.pieces):
    q0, plain, e0, ex = ch.next_piece()
    err += int((np.asarray(plain) > 0).sum())
torch.cuda.synchronize(); dt = time.perf_counter() - t0
blocks = a.pieces * a.piece * a.graphs * a.frames
print(json.dumps(dict(M=a.M, W=a.W, eps=a.eps, chains=a.graphs * a.frames, piece=a.piece, pieces=a.pieces, seconds=dt,
                      blocks_per_s=blocks / dt, vns_per_s=blocks * a.M / dt, block_error_rate=err / blocks, doped=a.doped)))
