"""Randomised parity sweep: random ensembles, sizes, channel parameters, caps and window shapes against the oracle."""
import numpy as np
import pytest

import fl_scaling_sc_ldpc_b200 as eng
from tests import util

pytestmark = pytest.mark.gpu
KEYS = ("iters", "residual", "blocks_err", "erasures_exp", "blocks_err_exp")


@pytest.mark.parametrize("trial", range(12))
def test_random_configuration(trial):
    rng = np.random.default_rng(9000 + trial)
    dv, dc = [(4, 8), (3, 6), (5, 10), (4, 8), (3, 9), (4, 12)][trial % 6]
    L = int(rng.integers(5, 26))
    cns_pos = int(rng.integers(6, 40))
    M = cns_pos * dc // dv
    if (M * dv) % dc:
        M = cns_pos * dc // dv + 1
        while (M * dv) % dc:
            M += 1
    thr = {(4, 8): 0.49, (3, 6): 0.48, (5, 10): 0.49, (3, 9): 0.31, (4, 12): 0.32}[(dv, dc)]
    eps = [float(thr + d) for d in rng.uniform(-0.12, 0.06, size=3)]
    G, F = int(rng.integers(1, 4)), int(rng.integers(1, 150))
    graphs, chan, _ = util.random_case(dv, dc, L, M, G, F, eps, seed=7000 + trial, doped_every=int(rng.integers(0, 6)))
    ens = eng.Ensemble(dv, dc, L, M)
    fb = eng.FrameBatch(ens, G, F).set_graphs(np.stack([g.vn_cn for g in graphs])).set_erasures(chan)
    for is_term in (True, False):
        cap = int(rng.choice([0, 1, 2, 3, 9, 30]))
        ref = util.oracle_bp(graphs, chan, cap, int(is_term), max_rows=80)
        res = eng.decode_bp_full(fb, cap, is_term, trajectory=True, max_rows=80)
        for k in KEYS:
            assert (getattr(res, k) == ref[k]).all(), (k, dv, dc, L, M, cap, is_term)
        assert (res.erased() == ref["erased"]).all()
        for g in range(G):
            for f in range(F):
                kk = min(80, ref["iters"][g, f])
                assert (res.rows[g, f, :kk] == ref["rows"][g, f, :kk]).all()
        W = int(rng.integers(1, L + 4))
        wcap = int(rng.choice([0, 1, 2, 5]))
        init = int(rng.choice([0, 1, 7]))
        square = bool(rng.integers(0, 2))
        ref = util.oracle_sw(graphs, chan, W, wcap, init if square else 0, int(square), int(is_term))
        res = eng.decode_bp_window(fb, W, wcap, init if square else 0, square, is_term)
        for k in KEYS + ("erasures_p1",):
            assert (getattr(res, k) == ref[k]).all(), (k, dv, dc, L, M, W, wcap, init, square, is_term)
        assert (res.erased() == ref["erased"]).all()
