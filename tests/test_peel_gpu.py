"""GPU parity tests of the peeling path (peel_kernels.cu, peeling_decoding.py) against the reference's own outputs
(tests/golden/peel_golden.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

import fl_scaling_sc_ldpc_b200 as eng
import oracle
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "peel_golden.npz"))
PEEL_SEED = 777


def _params(name):
    e, l, r, L, M, term = Z[name + "_params"][:6]
    return float(e), int(l), int(r), int(L), int(M), bool(term)


@pytest.mark.parametrize("name", ["t0", "t1", "t2", "t3"])
def test_peel_trajectories_match_reference(name):
    """r1 of the CUDA kernel == r1 of the reference's simulate_peeling_decoder_ldpc, same code / erasures / picks"""
    e, l, r, L, M, term = _params(name)
    cns, num_positions, total_size, steps = pdx._peel_geometry(e, l, r, L, M, term)
    F = Z[name + "_tr"].shape[0]
    ens = eng.Ensemble(l, r, L, M)
    # one frame per graph (the reference draws a new code per frame): G = F graphs x 1 frame
    fb = eng.FrameBatch(ens, F, 1, 2).set_graphs(Z[name + "_tr"])
    er = np.unpackbits(Z[name + "_er"], axis=1)[:, : L * M]
    fb.set_erasures(er[:, None, :])
    r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, PEEL_SEED, 0)
    assert (r1.reshape(F, -1).cpu().numpy() == Z[name + "_r1"]).all()
    total_generated = (L - len(Z[name + "_doping"])) * M
    plrs = (ner - rec).reshape(-1).cpu().numpy() / total_generated
    assert np.allclose(plrs, Z[name + "_plrs"], rtol=0, atol=1e-15)


def test_peel_many_frames_per_graph_matches_oracle():
    """frames sharing a code, ragged frame count, terminated and non-terminated, picks from the frame's Philox stream"""
    l, r, L, M, e = 4, 8, 16, 64, 0.45
    ens = eng.Ensemble(l, r, L, M)
    G, F = 3, 37
    fb = eng.FrameBatch(ens, G, F, 2).generate_graphs(5).generate_erasures(e, 6)
    vn_cn = fb.vn_cn.cpu().numpy()
    er = fb.erasures_host()
    for term in (False, True):
        cns, num_positions, total_size, steps = pdx._peel_geometry(e, l, r, L, M, term)
        r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, 99, 1000)
        r1 = r1.cpu().numpy(); rec = rec.cpu().numpy(); ner = ner.cpu().numpy()
        for g in range(G):
            for f in range(0, F, 5):
                picks = pdx.philox_picks(99, 1000 + g * F + f, steps)
                o_r1, o_rec = oracle.peel_trajectory(vn_cn[g], er[g, f], total_size, ens.nk, steps, picks)
                assert (r1[g, f] == o_r1).all(), (term, g, f)
                assert rec[g, f] == o_rec and ner[g, f] == er[g, f].sum()


def test_peel_global_bitmap_path(monkeypatch):
    """very large M keeps the degree-one bitmap in global memory; forced here on a small case"""
    monkeypatch.setenv("SCLDPC_PEEL_SMEM_LIMIT", "64")
    test_peel_trajectories_match_reference("t0")
    test_peel_trajectories_match_reference("t1")
    test_peel_many_frames_per_graph_matches_oracle()


def test_variance_accumulation_matches_numpy():
    """calc_nu_chunk / calc_var_chunk (est_scaling_params.py:90-94,131-138) restated in NumPy vs the fused kernel"""
    rng = np.random.default_rng(3)
    F, S = 23, 700
    r1 = rng.integers(0, 40, size=(F, S + 11)).astype(np.int32)
    r1[:, 500:] *= (rng.random((F, S + 11 - 500)) < 0.5)
    theory = np.concatenate([rng.random(S) * 30 + 0.1, np.zeros(5)])
    M = 1000
    # NumPy restatement of the reference functions
    crop = r1[:, : np.max(np.where(theory > 0)) + 1][:, 0:theory.shape[0]].astype(np.int64)
    th = theory[theory > 0] / M
    r1s = crop / M
    cen = r1s - th
    cen[r1s == 0] = np.nan
    ssq_ref = np.nansum(cen ** 2, axis=0)
    cnt_ref = np.sum(~np.isnan(cen), axis=0)
    chunks = [torch.as_tensor(r1[:10]).cuda(), torch.as_tensor(r1[10:]).cuda()]
    ssq, cnt = pdx.calc_nu_chunk_device(chunks, theory, M)
    assert (cnt == cnt_ref).all()
    # two chunks are summed chunk-wise like main_simulate_variance does; compare with the same association
    def chunk(a):
        c = a[:, :S] / M - th
        c[a[:, :S] == 0] = np.nan
        return np.nansum(c ** 2, axis=0)
    assert np.array_equal(ssq, chunk(r1[:10].astype(np.int64)) + chunk(r1[10:].astype(np.int64)))
    assert np.allclose(ssq, ssq_ref, rtol=1e-13)


def test_simulate_peeling_decoder_ldpc_signature_and_shapes():
    _, r1, plrs = pdx.simulate_peeling_decoder_ldpc(0.46, 4, 8, 12, 40, False, False, 6, seed=1)
    assert r1.shape == (6, int(40 * 12 * (0.46 + 0.1)) + 1) and r1.dtype == np.dtype("int") and plrs.shape == (6,)
    assert (r1[:, 0] > 0).all() and (plrs >= 0).all() and (plrs <= 1).all()
    _, r1b, plrsb = pdx.simulate_peeling_decoder_ldpc(0.46, 4, 8, 12, 40, False, False, 6, seed=1)
    assert (r1 == r1b).all() and (plrs == plrsb).all()                      # reproducible
    _, r1c, _ = pdx.simulate_peeling_decoder_ldpc(0.46, 4, 8, 12, 40, False, False, 6, [5, 6], seed=1)
    assert r1c.shape == r1.shape
    _, r1p, plrsp = pdx.simulate_peeling_decoder_ldpc(0.46, 4, 8, 12, 40, False, True, 4, seed=1)       # protograph ensemble
    assert r1p.shape == (4, r1.shape[1]) and (plrsp >= 0).all()
    with pytest.raises(NotImplementedError):
        pdx.simulate_peeling_decoder_ldpc(0.46, 4, 8, 12, 40, False, True, 2, {3: 0.5})                 # broken upstream


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "u0", "u1", "u2"])
def test_simulate_sc_ldpc_matches_reference(name):
    """the 13-tuple of the reference's simulate_sc_ldpc on the same injected codes and erasure masks (u*: unbounded)"""
    e, l, r, L, M, term = _params(name)
    bounded = bool(Z[name + "_params"][6])
    tail = 0 if term else 20
    Leff = L + tail + (0 if bounded else 20)
    tr_all = Z[name + "_tr"]
    er_all = np.unpackbits(Z[name + "_er"], axis=1)[:, : Leff * M]
    F = tr_all.shape[0]

    def factory(ens, G, fpg, gid0):
        assert fpg == 1
        fb = eng.FrameBatch(ens, G, 1, 2).set_graphs(tr_all[gid0:gid0 + G])
        return fb.set_erasures(er_all[gid0:gid0 + G, None, :])

    out = pdx.simulate_sc_ldpc(e, l, r, L, M, term, False, bounded, False, F, 10 ** 9, [], frames_per_graph=1, graphs_per_batch=5,
                               progress=False, _batch_factory=factory)
    got = [out[i] for i in (0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12)]
    assert np.allclose(got, Z[name + "_out"], rtol=0, atol=1e-15), (got, list(Z[name + "_out"]))
    assert out[8].shape == (F,) and not out[8].any() and not out[9].any()


def test_simulate_sc_ldpc_early_stop_is_sequential():
    """max_fuckups cuts at the same frame no matter how frames are batched (PD.py:698)"""
    a = pdx.simulate_sc_ldpc(0.52, 4, 8, 10, 32, True, False, True, False, 400, 7, [], seed=3, frames_per_graph=128,
                             graphs_per_batch=1, progress=False)
    b = pdx.simulate_sc_ldpc(0.52, 4, 8, 10, 32, True, False, True, False, 400, 7, [], seed=3, frames_per_graph=128,
                             graphs_per_batch=3, progress=False)
    assert a[5] == b[5] and a[:8] == b[:8] and a[4] <= 7 and round(a[0] * a[5]) == 7


def test_generated_ensemble_is_valid_and_uniformish():
    """on-device generate_code / gen_slots: every CN position receives a permutation of its sockets"""
    from fl_scaling_sc_ldpc_b200 import sc_ldpc as scx
    l, r, L, M = 4, 8, 9, 48
    cns = M * l // r
    scx.set_seed(11)
    tr = scx.gen_slots(l, r, L, M)
    assert tr.shape == (L * M, l) and tr.dtype == np.int64
    pos = np.repeat(np.arange(L), M)
    for d in range(l):
        assert ((tr[:, d] // cns) == pos + d).all()                        # edge d goes to CN position i + d
    deg = np.bincount(tr.reshape(-1), minlength=(L + l - 1) * cns)
    assert (deg[(l - 1) * cns: L * cns] == r).all() and deg.max() <= r      # interior CNs have degree r
    tb = scx.gen_slots_tail_biting(l, r, L, M)
    assert (np.bincount(tb.reshape(-1), minlength=L * cns) == r).all() and tb.max() < L * cns
    # different graph ids give different codes; the socket of VN 0 / edge 0 is roughly uniform over the CNs
    firsts = np.array([scx.gen_slots(l, r, L, M)[0, 0] for _ in range(200)])
    assert len(set(firsts.tolist())) > 15 and firsts.max() < cns


def test_ensemble_module_drop_ins():
    from fl_scaling_sc_ldpc_b200 import sc_ldpc as scx, sc_ldpc_protograph as spx
    scx.set_seed(3)
    vi = scx.gen_vn_indices(3, 6, 5, 12)
    assert vi.shape == (5, 3, 12)
    tr = scx.vn_indices_to_transmissions(vi, 3, 5, 12)
    assert tr.shape == (60, 3) and (tr[:12, 0] // 6 == 0).all()
    assert scx.gen_vn_indices_tail_biting(3, 6, 5, 12).shape == (5, 3, 12)
    g = spx.gen_slots_from_position(4, 8, 24)                                # sc_ldpc_protograph.py:17
    assert g.shape == (24, 4)
    for portion in range(2):
        for i in range(4):
            assert sorted(g[portion * 12:(portion + 1) * 12, i].tolist()) == list(range(i * 12, (i + 1) * 12))


def test_generated_protograph_ensemble_is_valid():
    """sc_ldpc_protograph.gen_slots_from_position: per (position, portion, edge type) a permutation of the position's CNs"""
    l, r, L, M = 4, 8, 7, 48
    cns = M * l // r
    ens = eng.Ensemble(l, r, L, M)
    fb = eng.FrameBatch(ens, 3, 2).generate_graphs(17, protograph=True)
    tr = fb.vn_cn.cpu().numpy()
    assert len({tr[g].tobytes() for g in range(3)}) == 3
    for g in range(3):
        t = tr[g].reshape(L, M // cns, cns, l)                              # [position][portion][vn][edge]
        for p in range(L):
            for q in range(M // cns):
                for i in range(l):
                    col = t[p, q, :, i]
                    assert (col // cns == p + i).all() and sorted((col % cns).tolist()) == list(range(cns))
        deg = np.bincount(tr[g].reshape(-1), minlength=(L + l - 1) * cns)
        assert (deg[(l - 1) * cns: L * cns] == r).all()
    # the decoders run on it unchanged
    fb.generate_erasures(0.40, 18)
    assert eng.decode_bp_full(fb, 0, True).residual.shape == (3, 2)
    out = pdx.simulate_sc_ldpc(0.40, 4, 8, 12, 64, True, True, True, False, 256, 10 ** 9, [], seed=5, progress=False)
    assert out[5] == 256 and 0.0 <= out[0] <= 1.0


def test_generated_channel_statistics_and_doping():
    ens = eng.Ensemble(4, 8, 10, 1000)
    fb = eng.FrameBatch(ens, 2, 100).generate_graphs(1).generate_erasures([0.3, 0.48], 2, doping_points=[4])
    er = fb.erasures_host()
    assert er.shape == (2, 100, ens.n)
    assert not er[:, :, 4000:5000].any()                                   # hard-doped position is known
    keep = np.r_[0:4000, 5000:10000]
    assert abs(er[0][:, keep].mean() - 0.3) < 0.005 and abs(er[1][:, keep].mean() - 0.48) < 0.005
    fb.generate_erasures(0.5, 2, doping_points={2: 0.25})
    er = fb.erasures_host()
    assert not er[:, :, 2000:2250].any() and er[:, :, 2250:3000].any()      # soft doping: first int(alpha*M) VNs
    a = eng.FrameBatch(ens, 1, 100).generate_erasures(0.4, 9, first_graph_id=1).erasures_host()
    b = eng.FrameBatch(ens, 2, 100).generate_erasures(0.4, 9, first_graph_id=0).erasures_host()
    assert (a[0] == b[1]).all()                                            # realisations depend on global ids only


@pytest.mark.parametrize("dv,dc,L,M", [(4, 8, 12, 40), (3, 6, 9, 30)])
def test_stopping_set_records_match_host_components(dv, dc, L, M):
    """scldpc_bp_stopping_sets (local "component has at most two VNs" predicate on the device) against connected components on
    the host (extract_stopping_sets, PD.py:1077-1095) for arbitrary residual patterns: sparse ones (isolated VNs, pairs,
    small trees), dense ones, and a mask of counted positions"""
    ens = eng.Ensemble(dv, dc, L, M)
    G, F = 2, 150
    fb = eng.FrameBatch(ens, G, F).generate_graphs(3)
    rng = np.random.default_rng(11)
    pat = np.zeros((G, F, ens.n), np.uint8)
    for g in range(G):
        for f in range(F):
            pat[g, f] = rng.random(ens.n) < (0.004, 0.01, 0.03, 0.08, 0.3)[f % 5]
    fb.set_erasures(pat)
    vn_cn = fb.vn_cn.cpu().numpy().astype(np.int64)
    for counted in (np.ones(L, bool), rng.random(L) < 0.7):
        rec = pdx.stopping_set_records(fb, fb.chan, counted)
        sizes = set()
        for g in range(G):
            for f in range(F):
                lost = np.flatnonzero(pat[g, f].astype(bool) & np.repeat(counted, M))
                exp = pdx.account_lost(lost, vn_cn[g], M) if len(lost) else (0, 0, 0, 0)
                assert tuple(rec[g, f]) == tuple(exp), (g, f, rec[g, f], exp)
                sizes |= {len(s) for s in pdx.extract_stopping_sets(lost, vn_cn[g])}
        assert {1, 2, 3} <= sizes          # the interesting component sizes all occurred
        assert (rec[:, F:] == 0).all()


def test_variance_accumulation_matches_the_reference_function():
    """P5 pinned to the reference itself: tests/golden/var_golden.npz holds what the imported est_scaling_params.calc_nu_chunk
    returned for two chunks (tests/golden/make_var_golden.py); the fused kernel must reproduce both chunks bit for bit"""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "var_golden.npz"))
    r1, theory, M = z["r1"], z["theory"], int(z["M"])
    for k, (lo, hi) in enumerate(((0, 10), (10, r1.shape[0]))):
        ssq, cnt = pdx.calc_nu_chunk_device([torch.as_tensor(r1[lo:hi]).cuda()], theory, M)
        assert np.array_equal(cnt, z[f"cnt{k}"]), k
        assert np.array_equal(ssq, z[f"ssq{k}"]), k                  # float64, same summation order: exact


def test_pairwise_complete_correlation_matches_pandas():
    """the reduction of calc_theta_explicit_ss_bounds (EST.py:161-189): DataFrame(zeros -> NaN).corr() over sampled steps,
    from exact device-side moment matrices accumulated over two chunks; and the theta fit built on it"""
    import pandas as pd
    from fl_scaling_sc_ldpc_b200 import est_scaling_params as esp
    rng = np.random.default_rng(9)
    F, S, M = 400, 260, 100
    # AR(1)-like integer trajectories with a known correlation length, zeros (finished frames) towards the end
    z = np.zeros((F, S))
    z[:, 0] = rng.normal(size=F)
    for t in range(1, S):
        z[:, t] = 0.93 * z[:, t - 1] + np.sqrt(1 - 0.93 ** 2) * rng.normal(size=F)
    r1 = np.maximum(1, np.rint(60 + 9 * z)).astype(np.int32)
    for f in range(F):
        r1[f, rng.integers(150, S + 60):] = 0
    start, stop, ivl = 20, 220, 3
    acc = esp.pairwise_moments([torch.as_tensor(r1[:150]).cuda(), torch.as_tensor(r1[150:]).cuda()], start, stop, ivl)
    c = esp.corr_from_moments(acc)
    x = r1[:, start:stop:ivl].astype(float) / M
    x[x == 0] = np.nan
    ref = pd.DataFrame(x).corr().to_numpy()
    assert c.shape == ref.shape and np.allclose(c, ref, rtol=0, atol=1e-12, equal_nan=True)
    n = acc[0].cpu().numpy()
    assert n.max() == F and n.min() < F and (n == n.T).all()
    # the fit the reference runs on that matrix (EST.py:211-243), same code path on the same numbers
    theta = esp.calc_theta_explicit_ss_bounds_ppd(r1, 20, 140, M)
    assert 0.03 < theta < 0.15                                      # -ln(0.93) = 0.0726 per step
