#!/usr/bin/env python3
"""Generates tests/golden/var_golden.npz: the reference's own calc_nu_chunk (fl_scaling/est_scaling_params.py:90-94 ->
calc_var_chunk :131-138, reached through peeling_decoding's star import, PD.py:44) applied to two chunks of r1
trajectories -- the per-chunk (ssquares, counts) main_simulate_variance sums (PD.py:1286-1292)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import peeling_ref as pr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def inputs():
    rng = np.random.default_rng(3)
    F, S = 23, 700
    r1 = rng.integers(0, 40, size=(F, S + 11)).astype(np.int64)
    r1[:, 500:] *= (rng.random((F, S + 11 - 500)) < 0.5)
    theory = np.concatenate([rng.random(S) * 30 + 0.1, np.zeros(5)])
    return r1, theory, 1000


def reference_chunks(pd, r1, theory, M):
    out = []
    for lo, hi in ((0, 10), (10, r1.shape[0])):
        ssq, cnt = pd.calc_nu_chunk(r1[lo:hi].copy(), theory.copy(), M)
        out.append((np.asarray(ssq, np.float64), np.asarray(cnt, np.int64)))
    return out


def main():
    pd = pr.load()
    r1, theory, M = inputs()
    (s0, c0), (s1, c1) = reference_chunks(pd, r1, theory, M)
    path = os.path.join(OUT, "var_golden.npz")
    np.savez_compressed(path, r1=r1.astype(np.int32), theory=theory, M=np.int64(M), ssq0=s0, cnt0=c0, ssq1=s1, cnt1=c1)
    print(path, os.path.getsize(path), s0[:3], c0[:3])


if __name__ == "__main__":
    main()
