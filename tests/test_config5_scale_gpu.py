"""BASELINE config 5 at its largest size: (4,8) SC-LDPC, L=50, M = 10^5 VNs per position (n = 5 M VNs, E = 20 M edges), doped.
A parity sample against the oracle port for both decoders on the paths that only such sizes reach -- the peeling kernel
with its degree-one bitmap in global memory (un-forced: 2.5 M CNs do not fit shared memory) and the trajectory kernels
with int32 edge ids close to their limit -- plus the device-side moment accumulators."""
import numpy as np
import pytest

import fl_scaling_sc_ldpc_b200 as eng
import oracle
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

pytestmark = pytest.mark.gpu
DV, DC, L, M = 4, 8, 50, 100000
DOPED = [24, 25]


@pytest.fixture(scope="module")
def big():
    ens = eng.Ensemble(DV, DC, L, M)
    fb = eng.FrameBatch(ens, 1, 128, 2).generate_graphs(515)
    g = oracle.Graph(fb.vn_cn[0].cpu().numpy(), L, M, ens.cns_pos, DV, DC)
    return ens, fb, g


def test_peeling_prefix_at_M_1e5_matches_oracle(big):
    """the first 1500 peeling steps of two frames (the oracle port is O(#CN) per step, like the reference)"""
    ens, fb, g = big
    e, steps = 0.48, 1500
    fb.generate_erasures(e, 516, doping_points=DOPED)
    cns, num_positions, total_size, _ = pdx._peel_geometry(e, DV, DC, L, M, False)
    assert total_size == 2500000
    r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, 517, 0)
    r1 = r1.cpu().numpy()
    ch = fb.erasures_host()[0]
    assert not ch[:, 24 * M: 26 * M].any() and ch[:, : 24 * M].any()          # hard doping: two positions known
    for f in (0, 127):
        picks = pdx.philox_picks(517, f, steps)
        o_r1, o_rec = oracle.peel_trajectory(g.vn_cn, ch[f], total_size, g.nk, steps, picks)
        assert (r1[0, f] == o_r1).all(), f
        assert int(rec[0, f]) == o_rec == steps


@pytest.mark.parametrize("force_node", [False, True])
def test_bp_trajectory_at_M_1e5_matches_oracle_and_moments(big, force_node, monkeypatch):
    """default at this size: message kernels; with SCLDPC_F_NODE_TRAJ the node-state sweep with its latch and counters"""
    ens, fb, g = big
    if force_node:
        from fl_scaling_sc_ldpc_b200 import _lib
        monkeypatch.setattr(eng.engine, "F_TRAJECTORY", _lib.F_TRAJECTORY | _lib.F_NODE_TRAJ)
    fb.generate_erasures(0.40, 518, doping_points=DOPED)
    cap = 400
    res, erased, rows, _ = eng.decode_bp_full(fb, cap, True, trajectory=True, max_rows=cap, collect=False)
    acc = eng.engine.trajectory_moments(fb, res[0], rows).cpu().numpy()
    r = eng.engine._collect(fb, res, erased, rows)
    ch = fb.erasures_host()[0]
    o = oracle.decode_bp(g, ch[5].astype(np.int32), cap, 1, max_rows=cap)
    k = o["iters"]
    assert int(r.iters[0, 5]) == k and int(r.residual[0, 5]) == o["residual"] == 0
    assert (r.rows[0, 5, :k] == o["rows"]).all()
    # moments = column sums of the rows over the 128 frames
    live = r.iters[0][:, None] > np.arange(cap)[None, :]
    dv = r.rows[0, :, :, 1].astype(np.int64)
    assert (acc[:, 0] == live.sum(axis=0)).all() and (acc[:, 2] == (dv * live).sum(axis=0)).all()
    assert (acc[:, 3] == (dv * dv * live).sum(axis=0)).all()
    assert acc[:, 2].sum() == 128 * ens.n - int(r.residual.sum())       # dVNs add up to everything that was resolved
