"""CPU tests: the oracle restatement against (a) the committed golden vectors generated from the compiled reference and
(b) the compiled reference itself (oracle/_ref) when it is available."""
import numpy as np
import pytest

import oracle
from tests import util

BP_CAPS = [100000, 5, 1]
SW_CFG = [(3, 100000, 0), (4, 4, 12), (5, 2, 0), (2, 1, 1)]


@pytest.mark.parametrize("path", util.golden_files(), ids=lambda p: p.split("bp_golden_")[-1])
def test_oracle_matches_golden(path):
    z, d = util.load_golden(path)
    assert d["G"] > 0
    for g in range(d["G"]):
        gr = oracle.Graph(z["vn_cn"][g], d["L"], d["vns_pos"], d["cns_pos"], d["dv"], d["dc"])
        for f in range(d["F"]):
            ch = z["chan"][g, f].astype(np.int32)
            for is_term in (1, 0):
                for cap in BP_CAPS:
                    key = f"bp_t{is_term}_c{cap}"
                    o = oracle.decode_bp(gr, ch, cap, is_term, max_rows=64)
                    st = z[key + "_stats"][g, f]
                    assert [o["iters"], o["residual"], o["blocks_err"], o["erasures_exp"], o["blocks_err_exp"]] == list(st), key
                    assert (o["erased"] == util.unpack_erased(z[key + "_erased"][g, f], d["n"])).all(), key
                    k = min(64, o["iters"])
                    assert (o["rows"][:k] == z[key + "_rows"][g, f, :k]).all(), key
            for (W, cap, init) in SW_CFG:
                for square in (1, 0):
                    key = f"sw_s{square}_W{W}_c{cap}_i{init}"
                    o = oracle.decode_bp_sw(gr, ch, W, cap, init if square else 0, square, 1)
                    st = z[key + "_stats"][g, f]
                    assert [o["residual"], o["erasures_p1"], o["blocks_err"], o["erasures_exp"], o["blocks_err_exp"]] == list(st), key
                    assert (o["erased"] == util.unpack_erased(z[key + "_erased"][g, f], d["n"])).all(), key


def test_golden_graphs_are_the_reference_ensemble():
    """generate_code restatement reproduces the reference's graphs from the recorded srandom() seed."""
    for path in util.golden_files():
        z, d = util.load_golden(path)
        oracle.srandom(int(z["seed"]))
        perm = None
        g0, perm = oracle.generate_code(d["L"], d["vns_pos"], d["cns_pos"], d["dv"], d["dc"], perm)
        assert (g0.vn_cn == z["vn_cn"][0]).all()


def test_doping_predicate_kat():
    """test_is_position_doped_streaming (BP_FULL.c:1891): {5,7,9} => doped iff pos % 10 in {5,7,9}."""
    for pos in range(100):
        assert oracle.is_position_doped_streaming(pos, [5, 7, 9]) == (pos % 10 in (5, 7, 9))
    assert not oracle.is_position_doped_streaming(3, [])


def _ref_available():
    from oracle import ref_driver
    return ref_driver.available("traj", 4, 8, 10, 25)


@pytest.mark.skipif(not _ref_available(), reason="compiled reference (oracle/_ref) not present")
def test_oracle_matches_compiled_reference():
    from oracle import ref_driver as rd
    rng = np.random.default_rng(7)
    dv, dc, L, defM = 4, 8, 10, 25
    rt, rs, rf = (rd.get(v, dv, dc, L, defM) for v in ("traj", "sw", "full"))
    for trial in range(6):
        seed = int(rng.integers(1, 1 << 30))
        eps = float(rng.choice([0.35, 0.44, 0.48, 0.52]))
        rt.srandom(seed); rt.reset_perm()
        g_ref = rt.generate_code(); ch_ref = rt.channel_doped(eps, [3] if trial % 3 == 0 else [])
        oracle.srandom(seed)
        g, _ = oracle.generate_code(L, 2 * defM, defM, dv, dc)
        ch = oracle.channel_doped(g.n, eps, 2 * defM, [3] if trial % 3 == 0 else [])
        assert (g.vn_cn == g_ref).all() and (ch == ch_ref).all()
        for is_term in (1, 0):
            for cap in (1000, 3):
                a = rt.decode_bp(cap, is_term); b = oracle.decode_bp(g, ch, cap, is_term, max_rows=1000)
                assert a["residual"] == b["residual"] and a["blocks_err"] == b["blocks_err"]
                assert a["erasures_exp"] == b["erasures_exp"] and a["blocks_err_exp"] == b["blocks_err_exp"]
                assert (a["erased"] == b["erased"]).all() and len(a["rows"]) == b["iters"] and (a["rows"][:, 1:] == b["rows"]).all()
        rs.set_graph(g.vn_cn); rs.set_channel(ch); rf.set_graph(g.vn_cn); rf.set_channel(ch)
        for (W, cap, init) in ((3, 1000, 0), (4, 3, 9)):
            a = rs.decode_bp_sw(W, cap, init); b = oracle.decode_bp_sw(g, ch, W, cap, init, 1, 1)
            assert all(a[k] == b[k] for k in ("residual", "erasures_p1", "blocks_err", "erasures_exp", "blocks_err_exp"))
            assert (a["erased"] == b["erased"]).all()
            a = rf.decode_bp_sw(W, cap); b = oracle.decode_bp_sw(g, ch, W, cap, 0, 0, 1)
            assert all(a[k] == b[k] for k in ("residual", "erasures_p1", "blocks_err", "erasures_exp", "blocks_err_exp"))
            assert (a["erased"] == b["erased"]).all()
