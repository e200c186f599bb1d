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
@pytest.mark.parametrize("name", ["c0", "c1", "c2", "c3"])
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
