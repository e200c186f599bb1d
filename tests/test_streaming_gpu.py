"""GPU: the streaming decoder built on the window kernels against the reference's own streaming outputs
(tests/golden/stream_golden.npz) and the published file format."""
import os

import numpy as np
import pytest

from fl_scaling_sc_ldpc_b200 import streaming

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stream_golden.npz"))


def test_doping_predicate_kat():
    for pos in range(60):
        assert streaming.is_position_doped_streaming(pos, [5, 7, 9]) == (pos % 10 in (5, 7, 9))
    assert not streaming.is_position_doped_streaming(4, [])


def test_row_format():
    t = dict(num_erasures=176, num_bits_generated=1056, num_blocks_err=36, num_blocks_generated=66, num_erasures_exp=170,
             num_bits_generated_exp=992, num_blocks_err_exp=32, num_blocks_generated_exp=62)
    row = streaming.result_row(0.42, t)
    assert row == "0.420000 1.666667e-01 5.454545e-01 1.713710e-01 5.161290e-01 176 1056 36 66 170 992 32 62\n"
    assert len(streaming.HEADER.split()) == 13


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c0", "c1", "c2", "c3", "c4"])
def test_streaming_decoder_matches_reference(name):
    import fl_scaling_sc_ldpc_b200 as eng
    dv, dc, L, defM, W, steps = (int(x) for x in Z[name + "_params"])
    vn_cn, chan, ref = Z[name + "_vn_cn"], Z[name + "_chan"], Z[name + "_steps"]
    npos, V = vn_cn.shape[0], vn_cn.shape[1]
    ens = eng.Ensemble(dv, dc, npos, V)
    fb = eng.FrameBatch(ens, 1, 3)
    fb.set_graphs(vn_cn.reshape(1, -1, dv))
    pat = np.stack([chan.reshape(-1), np.zeros(npos * V, np.uint8), chan.reshape(-1)])      # lanes 0 and 2 carry the frame
    fb.set_erasures(pat[None])
    plain, ex = streaming.decode_segment(ens, W, 0.0, [], 1, 3, 0, fb=fb)
    for lane in (0, 2):
        c = streaming.stream_counters(plain[0, :, lane], ex[0, :, lane], steps, dv, list(Z[name + "_doped"]), V)
        got = np.stack([c["erasures_pos"], c["num_blocks_err"], c["num_erasures_exp"], c["num_blocks_err_exp"]], axis=1)
        assert (got == ref).all(), (name, lane)
    assert not plain[0, :, 1].any()


@pytest.mark.gpu
def test_simulate_stream_and_cli(tmp_path):
    t = streaming.simulate_stream(0.50, 4, 8, 32, 6, [9], segment=60, max_blocks_err=40, max_blocks=10 ** 6, seed=3,
                                  frames_per_graph=16, graphs_per_batch=2)
    assert t["num_blocks_err_exp"] == 40 and t["num_blocks_generated"] >= t["num_blocks_err"] > 0
    assert t["num_bits_generated"] == 32 * t["num_blocks_generated"]
    assert streaming.main_streaming(["2", "6", "1", "9", "--M", "16", "--points", "2", "--eps-ini", "0.5", "--eps-delta", "0.02",
                                     "--max-blocks-err", "20", "--segment", "50", "--outdir", str(tmp_path)]) == 0
    lines = (tmp_path / "SC_LDPC_4_8_L50_M16_DOP1_BP_Stream_SW6_Random_BLER_2.dat").read_text().splitlines()
    assert lines[0] == streaming.HEADER.strip() and len(lines) == 3 and lines[1].startswith("0.500000 ")


def _golden_provider(name, lanes=3):
    """the unrolled chain the reference generated, served in pieces: local CN ids, the frame in lanes 0 and 2"""
    dv, dc, L, defM, W, steps = (int(x) for x in Z[name + "_params"])
    vn_cn, chan = Z[name + "_vn_cn"].astype(np.int64), Z[name + "_chan"]
    npos, V = vn_cn.shape[0], vn_cn.shape[1]

    def provider(A, Lloc):
        assert A + Lloc <= npos
        g = (vn_cn[A:A + Lloc] - A * defM).reshape(1, Lloc * V, dv)
        c = chan[A:A + Lloc].reshape(-1)
        pat = np.zeros((1, lanes, Lloc * V), np.uint8)
        pat[0, 0] = c
        pat[0, lanes - 1] = c
        return g, pat
    return provider, (dv, dc, L, defM, W, steps, npos, V)


@pytest.mark.gpu
@pytest.mark.parametrize("name,piece", [("c0", 16), ("c1", 17), ("c2", 14), ("c3", 10), ("c4", 23), ("c4", 60)])
def test_chain_pieces_continue_the_chain_exactly(name, piece):
    """ChainPieces (state carried from piece to piece, scldpc_bp_window_range with resume) against the reference's
    streaming decoder step by step: the decisions and the lagging expurgation must not notice the piece boundaries --
    in particular the erasures a failed window leaves behind keep propagating (c4: every window fails)"""
    provider, (dv, dc, L, defM, W, steps, npos, V) = _golden_provider(name)
    ref = Z[name + "_steps"]
    ch = streaming.ChainPieces(dv, dc, V, W, 0.0, [], 1, 3, 0, piece=piece, provider=provider)
    n_pieces = (npos - (ch.Lloc - piece)) // piece
    assert n_pieces >= 3
    plain = np.zeros((3, n_pieces * piece), np.int64)
    ex = np.zeros((3, n_pieces * piece), np.int64)
    for k in range(n_pieces):
        q0, pl, e0, e = ch.next_piece()
        assert q0 == k * piece
        plain[:, q0:q0 + pl.shape[1]] = pl[0].T
        ex[:, e0:e0 + e.shape[1]] = e[0].T
    ms = dv - 1
    n_steps = min(steps, n_pieces * piece - ms + ms)          # steps whose decided and expurgated positions are final
    n_steps = min(steps, n_pieces * piece)
    for lane in (0, 2):
        c = streaming.stream_counters(plain[lane], ex[lane], n_steps, dv, list(Z[name + "_doped"]), V)
        got = np.stack([c["erasures_pos"], c["num_blocks_err"], c["num_erasures_exp"], c["num_blocks_err_exp"]], axis=1)
        assert (got == ref[:n_steps]).all(), (name, lane, np.flatnonzero((got != ref[:n_steps]).any(axis=1))[:5])
    assert not plain[1].any()


@pytest.mark.gpu
def test_simulate_stream_replays_the_reference_counters_over_pieces():
    """simulate_stream on ONE chain (the golden c4 chain, 23 positions per piece): the row it returns when the block budget
    runs out is the reference's running counters at that step"""
    provider, (dv, dc, L, defM, W, steps, npos, V) = _golden_provider("c4", lanes=1)
    ref = Z["c4_steps"]
    budget = 150
    t = streaming.simulate_stream(0.0, dv, dc, V, W, [], segment=23, max_blocks_err=10 ** 9, max_blocks=budget, seed=0,
                                  frames_per_graph=1, graphs_per_batch=1, _provider=provider)
    k = budget + 2 * dv - 2                                   # step at which the budget-th expurgated block has been generated
    assert t["num_blocks_generated_exp"] == budget and t["num_blocks_generated"] == budget + dv
    assert (t["num_blocks_err"], t["num_erasures_exp"], t["num_blocks_err_exp"]) == tuple(int(x) for x in ref[k, 1:4])
    assert t["num_erasures"] == int(ref[:k + 1, 0].sum())
