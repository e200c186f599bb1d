"""CPU tests of the peeling path: the oracle restatements and the product's host-side bookkeeping against golden
vectors produced by the reference's own peeling_decoding.py (tests/golden/make_peel_golden.py)."""
import os

import numpy as np
import pytest

import oracle
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "peel_golden.npz"))
PEEL_SEED = 777


def _params(name):
    e, l, r, L, M, term = Z[name + "_params"][:6]
    return float(e), int(l), int(r), int(L), int(M), bool(term)


@pytest.mark.parametrize("name", ["t0", "t1", "t2", "t3"])
def test_oracle_peel_trajectory_matches_reference(name):
    e, l, r, L, M, term = _params(name)
    cns = int(l / r * M)
    num_positions = L + l - 1 if term else L
    steps = int(M * num_positions * (e + 0.1))
    assert steps + 1 == Z[name + "_r1"].shape[1]
    dop = list(Z[name + "_doping"])
    total_generated = (L - len(dop)) * M
    for f in range(Z[name + "_tr"].shape[0]):
        er = np.unpackbits(Z[name + "_er"][f])[: L * M]
        picks = Z[name + "_picks"][f]
        assert (picks[:steps] == pdx.philox_picks(PEEL_SEED, f, steps)).all()     # goldens use the kernel's Philox draws
        r1, rec = oracle.peel_trajectory(Z[name + "_tr"][f], er, cns * num_positions, cns * (L + l - 1), steps, picks)
        assert (r1 == Z[name + "_r1"][f]).all()
        assert abs((er.sum() - rec) / total_generated - Z[name + "_plrs"][f]) < 1e-15


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "u0", "u1", "u2"])
def test_error_rate_bookkeeping_matches_reference(name):
    """product bookkeeping (counted_positions / account_lost / extract_stopping_sets) fed by the oracle's fixed point"""
    e, l, r, L, M, term = _params(name)
    bounded = bool(Z[name + "_params"][6])
    tail = 0 if term else 20
    head, head_sched = (0, 0) if bounded else (20, 10)
    Leff = L + tail + head
    cns = int(l / r * M)
    num_positions = Leff + l - 1 if term else Leff
    total_size = cns * num_positions
    counted = np.repeat(pdx.counted_positions(l, Leff, num_positions, head, tail), M)
    nf = nft = fail = fail_e = blocks_e = 0
    F = Z[name + "_tr"].shape[0]
    for f in range(F):
        tr = Z[name + "_tr"][f]
        er = np.unpackbits(Z[name + "_er"][f])[: Leff * M]
        lost_mask = oracle.peel_fixed_point(tr, er, cns * (Leff + l - 1), head_sched * cns, total_size).astype(bool) & counted
        lost = np.flatnonzero(lost_mask)
        if len(lost):
            n, big, le, be = pdx.account_lost(lost, tr.astype(np.int64), M)
            nf += 1; nft += int(big); fail += n; fail_e += le; blocks_e += be
    gen = (Leff - tail - head) * M * F
    blocks = (Leff - tail - head) * F
    exp = Z[name + "_out"]   # FER, FER_exp, PLR, PLR_exp, n_failed_exp, n_frames, n_vn_failed_exp, n_vn_gen, blocks_failed_exp, blocks_gen, BLER_exp
    got = [nf / F, nft / F, fail / gen, fail_e / gen, nft, F, fail_e, gen, blocks_e, blocks, blocks_e / blocks]
    assert np.allclose(got, exp, rtol=0, atol=1e-15), (got, list(exp))


def test_reference_sic_known_answer_vector_is_recorded():
    """test_2_6_csa_sync (PD.py:1164): nothing decodes in slots 0-9, all three users at t = 10, empty schedule"""
    lines = bytes(Z["csa_sync_stdout"]).decode().strip().split("\n")
    assert lines[:10] == [f"{t} set()" for t in range(10)]
    assert lines[10].startswith("10 {") and lines[10].count("User(uid=") == 3
    assert lines[-1] == "{}"


def test_philox_picks_are_stable():
    a = pdx.philox_picks(1, 2, 9)
    b = pdx.philox_picks(1, 2, 12)
    assert (a == b[:9]).all() and len(set(b.tolist())) == 12
    assert (pdx.philox_picks(1, 3, 4) != a[:4]).any()


def test_num_pd_steps_float_truncation():
    """App. A-16: int(1000*50*(0.48+0.1)) == 28999"""
    assert pdx._peel_geometry(0.48, 4, 8, 50, 1000, False)[3] == 28999
    assert pdx._peel_geometry(0.48, 4, 8, 50, 1000, True)[3] == 30739
    assert pdx._peel_geometry(0.48, 4, 8, 50, 10000, False)[3] == 290000


def test_variance_golden_is_what_the_reference_computes():
    """tests/golden/var_golden.npz against the imported reference (this container only; the GPU test compares the kernel
    with the stored arrays)"""
    from oracle import peeling_ref as pr
    if not pr.available():
        pytest.skip("reference not present")
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_var_golden", os.path.join(here, "golden", "make_var_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(os.path.join(here, "golden", "var_golden.npz"))
    r1, theory, M = mod.inputs()
    assert np.array_equal(r1.astype(np.int32), z["r1"]) and np.array_equal(theory, z["theory"])
    for k, (ssq, cnt) in enumerate(mod.reference_chunks(pr.load(), r1, theory, M)):
        assert np.array_equal(ssq, z[f"ssq{k}"]) and np.array_equal(cnt, z[f"cnt{k}"])
