#!/usr/bin/env python3
"""Generates tests/golden/peel_golden.npz by running the reference's own peeling_decoding.py (imported from
/root/reference in the build container) on injected codes, erasure masks and pick sequences.

  * simulate_peeling_decoder_ldpc (PD.py:705): r1 trajectories + plrs, terminated and non-terminated, hard doping
  * simulate_sc_ldpc (PD.py:591): the 13-tuple on injected frames, terminated / non-terminated
  * test_2_6_csa_sync (PD.py:1164): the SIC known-answer vector of the reference (3 users, 6 slots)
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import peeling_ref as pr  # noqa: E402
from fl_scaling_sc_ldpc_b200.peeling_decoding import philox_picks  # noqa: E402

PEEL_SEED = 777

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_code(pd, l, r, L, M, rng):
    """a code drawn by the reference's own sc_ldpc.gen_slots (NumPy global state seeded by the caller)"""
    import sc_ldpc
    return sc_ldpc.gen_slots(l, r, L, M)


def main():
    pd = pr.load()
    rng = np.random.default_rng(424242)
    np.random.seed(424242)
    arrays = {}
    # ---- trajectories ----
    cases = [("t0", 0.46, 4, 8, 12, 40, False, []), ("t1", 0.44, 4, 8, 10, 48, True, []), ("t2", 0.40, 3, 6, 10, 36, False, []),
             ("t3", 0.47, 4, 8, 14, 32, False, [6, 7])]
    for name, e, l, r, L, M, term, dop in cases:
        num_positions = L + l - 1 if term else L
        steps = int(M * num_positions * (e + 0.1))
        frames = []
        for f in range(5):
            tr = ref_code(pd, l, r, L, M, rng)
            er = rng.random(L * M) <= e
            if dop:
                cns = int(l / r * M)
                pos_chain = (tr[:, 0] // cns)
                er = er & ~np.isin(pos_chain, dop)
            # the same Philox draws the CUDA kernel uses for global frame f (host function of libscldpc, no GPU needed)
            picks = philox_picks(PEEL_SEED, f, steps + 4)
            frames.append((tr, er, picks))
        r1, plrs = pr.ref_peel_trajectories(e, l, r, L, M, term, frames, dop)
        arrays[name + "_params"] = np.array([e, l, r, L, M, int(term)], np.float64)
        arrays[name + "_doping"] = np.array(dop, np.int32)
        arrays[name + "_tr"] = np.array([f[0] for f in frames], np.int32)
        arrays[name + "_er"] = np.packbits(np.array([f[1] for f in frames], np.uint8), axis=1)
        arrays[name + "_picks"] = np.array([f[2] for f in frames], np.uint32)
        arrays[name + "_r1"] = r1.astype(np.int32)
        arrays[name + "_plrs"] = plrs
    # ---- error rates ----
    cases = [("s0", 0.47, 4, 8, 12, 40, True, True), ("s1", 0.45, 4, 8, 10, 32, False, True), ("s2", 0.50, 4, 8, 8, 24, True, True),
             ("u0", 0.42, 4, 8, 10, 32, True, False), ("u1", 0.37, 4, 8, 6, 24, False, False),
             ("u2", 0.44, 4, 8, 8, 40, True, False)]                                            # u*: is_bounded = False
    for name, e, l, r, L, M, term, bounded in cases:
        Leff = L + (0 if term else 20) + (0 if bounded else 20)
        frames = []
        for f in range(12):
            tr = ref_code(pd, l, r, Leff, M, rng)
            er = rng.random(Leff * M) <= e
            frames.append((tr, er))
        with redirect_stdout(io.StringIO()):
            out = pr.ref_simulate_sc_ldpc(e, l, r, L, M, term, bounded, frames)
        arrays[name + "_params"] = np.array([e, l, r, L, M, int(term), int(bounded)], np.float64)
        arrays[name + "_tr"] = np.array([f[0] for f in frames], np.int32)
        arrays[name + "_er"] = np.packbits(np.array([f[1] for f in frames], np.uint8), axis=1)
        arrays[name + "_out"] = np.array([out[i] for i in (0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12)], np.float64)
    # ---- the reference's only SIC known-answer vector ----
    buf = io.StringIO()
    with redirect_stdout(buf):
        pd.test_2_6_csa_sync()
    arrays["csa_sync_stdout"] = np.frombuffer(buf.getvalue().encode(), dtype=np.uint8)
    path = os.path.join(OUT, "peel_golden.npz")
    np.savez_compressed(path, **arrays)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
