"""Statistical acceptance against the reference's own published simulation results (`sim_data/error_rates/*.dat`,
SURVEY.md section 6): graphs and channels are drawn by the on-device generators (different RNG streams than the
authors'), so agreement is within Monte-Carlo error.  Numbers quoted from the reference files (file:line)."""
import numpy as np
import pytest

import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

pytestmark = pytest.mark.gpu
ENS = eng.Ensemble(4, 8, 50, 1000)          # (4,8), L=50, M=1000 (C Def_M=500)


def run(decoder, eps, n_graphs=256, fpg=128, seed=2024):
    res, blocks, frames = 0, 0, 0
    ferr = 0
    for g0 in range(0, n_graphs, 64):
        fb = eng.FrameBatch(ENS, 64, fpg).generate_graphs(seed, first_graph_id=g0).generate_erasures(eps, seed + 1, first_graph_id=g0)
        r = decoder(fb)
        res += int(r.residual.sum()); blocks += int(r.blocks_err.sum()); ferr += int((r.residual > 0).sum()); frames += r.residual.size
    return ferr / frames, res / frames / ENS.n, blocks / frames / ENS.L, frames


def close(got, published, n_events, rel=0.12):
    """within 12 % or 5 sigma of the counting error, whichever is larger (frames sharing a graph are correlated)"""
    sigma = published / np.sqrt(max(n_events, 1))
    return abs(got - published) <= max(rel * published, 5 * sigma)


@pytest.mark.parametrize("eps,fer,ber,bler", [
    (0.4625, 0.0568473, 0.00377169, 0.016011),     # SC_LDPC_4_8_L50_M500_BP_Full_175it_BEC.dat:13
    (0.4650, 0.213881, 0.0143526, 0.0607443),      # :15
])
def test_full_bp_175_iterations_matches_published(eps, fer, ber, bler):
    got_fer, got_ber, got_bler, n = run(lambda fb: eng.decode_bp_full(fb, 175, True), eps)
    assert close(got_fer, fer, fer * n), (got_fer, fer)
    assert close(got_ber, ber, fer * n), (got_ber, ber)
    assert close(got_bler, bler, fer * n), (got_bler, bler)


def test_sliding_window_matches_published():
    """SC_LDPC_4_8_L50_N1000_BP_SW20_6it_60init_square_BEC.dat:7  (eps 0.46: FER 0.061105, BER 0.00653761)"""
    got_fer, got_ber, got_bler, n = run(lambda fb: eng.decode_bp_window(fb, 20, 6, 60, square=True, is_term=True), 0.46, n_graphs=128)
    assert close(got_fer, 0.061105, 0.061105 * n), got_fer
    assert close(got_ber, 0.00653761, 0.061105 * n), got_ber
    assert close(got_bler, 0.0228884, 0.061105 * n), got_bler


def test_peeling_unlimited_matches_published():
    """terminated_fer_plr_sc_ldpc_4_8_50_1000.dat:20  (eps 0.47: 5954 / 100000 frames, PLR 0.0108755)"""
    out = pdx.simulate_sc_ldpc(0.47, 4, 8, 50, 1000, True, False, True, False, 24576, 10 ** 9, [], seed=7, frames_per_graph=128,
                               graphs_per_batch=64, progress=False)
    assert close(out[0], 0.05954, 0.05954 * out[5]), out[0]
    assert close(out[2], 0.0108755, 0.05954 * out[5]), out[2]
    assert out[1] <= out[0] and out[3] <= out[2]
