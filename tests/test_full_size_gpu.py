"""Parity at BASELINE.json's full sizes, through the oracle on a sample and through size-independent properties:
fixed-point (stopping-set) structure, cap monotonicity, trajectory bookkeeping, window == full BP when the window
covers the chain, wave tracking == sweeping everything, peeling and BP agree on what is recoverable."""
import numpy as np
import pytest

import fl_scaling_sc_ldpc_b200 as eng
import oracle
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

pytestmark = pytest.mark.gpu
DV, DC, L, M = 4, 8, 50, 10000


@pytest.fixture(scope="module")
def big_batch():
    ens = eng.Ensemble(DV, DC, L, M)
    fb = eng.FrameBatch(ens, 1, 128, 2).generate_graphs(2026).generate_erasures(0.46, 2027)
    return ens, fb


def test_full_size_frame_matches_oracle(big_batch):
    """(4,8), L=50, M=10000, eps=0.46: one frame end to end against the CPU oracle (iterations, residual, rows)"""
    ens, fb = big_batch
    res = eng.decode_bp_full(fb, 0, True, trajectory=True, max_rows=400)
    g = oracle.Graph(fb.vn_cn[0].cpu().numpy(), L, M, ens.cns_pos, DV, DC)
    ch = fb.erasures_host()[0]
    for f in (0, 77):
        o = oracle.decode_bp(g, ch[f].astype(np.int32), 10 ** 9, 1, max_rows=400)
        assert res.iters[0, f] == o["iters"] and res.residual[0, f] == o["residual"]
        assert (res.rows[0, f, :o["iters"]] == o["rows"]).all()
        assert (res.erased()[0, f] == o["erased"]).all()


def test_full_size_fixed_point_and_caps(big_batch):
    ens, fb = big_batch
    fb2 = eng.FrameBatch(ens, 1, 128, 2)
    fb2.vn_cn.copy_(fb.vn_cn); fb2._build_tables()
    fb2.generate_erasures(0.49, 31)                       # close to the threshold: a good fraction of frames stalls
    unl = eng.decode_bp_full(fb2, 0, True)
    assert (unl.residual > 0).any() and (unl.residual == 0).any()
    ch = fb2.erasures_host()[0]
    er = unl.erased()[0]
    vn_cn = fb2.vn_cn[0].cpu().numpy()
    assert not (er & ~ch).any()                           # only channel erasures can stay erased
    for f in np.flatnonzero(unl.residual[0] > 0)[:4]:
        # stopping set: no CN sees exactly one erased VN
        cnt = np.bincount(vn_cn[er[f].astype(bool)].reshape(-1), minlength=ens.nk)
        assert (cnt != 1).all() and er[f].sum() == unl.residual[0, f]
        assert unl.blocks_err[0, f] == len(np.unique(np.flatnonzero(er[f]) // M))
    prev = None
    for cap in (40, 120, 400):
        r = eng.decode_bp_full(fb2, cap, True)
        assert (r.iters == np.minimum(cap, unl.iters)).all()
        assert (r.residual >= unl.residual).all()
        if prev is not None:
            assert (r.residual <= prev).all()
        prev = r.residual
    # trajectory bookkeeping: the dVNs column sums to n - residual, deg-1 CNs bound the recovered VNs (BP_FULL.c:1035)
    t = eng.decode_bp_full(fb2, 0, True, trajectory=True, max_rows=int(unl.iters.max()))
    for f in range(0, 128, 9):
        k = t.iters[0, f]
        assert t.rows[0, f, :k, 1].sum() == ens.n - t.residual[0, f]
        assert (t.rows[0, f, 1:k, 0] >= t.rows[0, f, 1:k, 1]).all()
        assert t.rows[0, f, k - 1, 2] == (L if t.residual[0, f] == 0 else np.flatnonzero(er[f])[0] // M)


def test_full_size_wave_tracking_equals_sweeping_everything(big_batch, monkeypatch):
    ens, fb = big_batch
    a = eng.decode_bp_full(fb, 0, True)
    monkeypatch.setenv("SCLDPC_NO_WAVE", "1")
    b = eng.decode_bp_full(fb, 0, True)
    assert (a.iters == b.iters).all() and (a.residual == b.residual).all()
    assert bool((a.erased_words == b.erased_words).all())
    monkeypatch.setenv("SCLDPC_NO_WAVE", "0")
    c = eng.decode_bp_full(fb, 0, False)                  # truncated
    monkeypatch.setenv("SCLDPC_NO_WAVE", "1")
    d = eng.decode_bp_full(fb, 0, False)
    assert (c.iters == d.iters).all() and (c.residual == d.residual).all() and (c.residual > 0).all()


def test_window_covering_the_chain_equals_full_bp():
    """L=100 (BASELINE config 4 length) at M=2000: a window as long as the chain with unlimited iterations is full BP"""
    ens = eng.Ensemble(DV, DC, 100, 2000)
    fb = eng.FrameBatch(ens, 1, 128, 2).generate_graphs(5).generate_erasures(0.47, 6)
    full = eng.decode_bp_full(fb, 0, True)
    win = eng.decode_bp_window(fb, 103, 0, 0, square=True, is_term=True)
    assert bool((full.erased_words == win.erased_words).all()) and (full.residual == win.residual).all()
    # a realistic window: W=10, 8 iterations per position, 60 for the first (sim_data/.../SW20_8it_60init naming)
    r = eng.decode_bp_window(fb, 10, 8, 60, square=True, is_term=True)
    assert (r.residual >= full.residual).all() and r.edge_updates > 0
    assert (r.iters <= 60 + 99 * 8).all()


def test_peeling_and_bp_agree_on_config_1():
    """BASELINE config 1: (4,8), L=50, M=1000, non-terminated, eps=0.48.  Random-order peeling ends in the same
    residual as truncated flooding BP, r1 starts at the number of degree-one CNs and ends at 0."""
    l, r, Lc, Mc, e = 4, 8, 50, 1000, 0.48
    ens = eng.Ensemble(l, r, Lc, Mc)
    fb = eng.FrameBatch(ens, 2, 24, 2).generate_graphs(8).generate_erasures(e, 9)
    cns, num_positions, total_size, steps = pdx._peel_geometry(e, l, r, Lc, Mc, False)
    assert steps == 28999
    r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, 10, 0)
    r1 = r1.cpu().numpy(); lost = (ner - rec).cpu().numpy()
    bp = eng.decode_bp_full(fb, 0, is_term=False)
    assert (lost == bp.residual).all()
    ch = fb.erasures_host(); vn_cn = fb.vn_cn.cpu().numpy()
    for g in range(2):
        for f in (0, 23):
            deg = np.bincount(vn_cn[g][ch[g, f].astype(bool)].reshape(-1), minlength=ens.nk)[:total_size]
            assert r1[g, f, 0] == (deg == 1).sum()
            assert r1[g, f, -1] == 0 and (np.abs(np.diff(r1[g, f])) <= l).all()
            assert (r1[g, f] > 0).sum() == rec.cpu().numpy()[g, f]       # one VN per step until the trajectory hits zero
