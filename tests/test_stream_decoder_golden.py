"""The reference's streaming / circular-buffer decoder (decodeBP_SW_circular + main_streaming, compiled out upstream) is
the classical window decoder with unlimited per-window iterations on the unrolled chain: the ring buffer is a memory
device, and check nodes beyond the window send erasures in both.  Pinned here against outputs of the reference itself
(tests/golden/stream_golden.npz, made with BP_FULL.c built with CIRCULAR defined)."""
import os

import numpy as np
import pytest

import oracle

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stream_golden.npz"))
CASES = ["c0", "c1", "c2", "c3", "c4"]


def case(name):
    dv, dc, L, defM, W, steps = (int(x) for x in Z[name + "_params"])
    vn_cn, chan = Z[name + "_vn_cn"], Z[name + "_chan"]
    return dv, dc, L, defM, W, steps, vn_cn, chan, Z[name + "_steps"]


@pytest.mark.parametrize("name", CASES)
def test_unrolled_window_decoder_reproduces_the_streaming_reference(name):
    dv, dc, L, defM, W, steps, vn_cn, chan, ref = case(name)
    npos, V = vn_cn.shape[0], vn_cn.shape[1]
    g = oracle.Graph(vn_cn.reshape(-1, dv), npos, V, defM, dv, dc)
    o = oracle.decode_bp_sw(g, chan.reshape(-1).astype(np.int32), W, 10 ** 9, 0, square=0, is_term=1)
    plain, ex = oracle.position_counts(g, o["erased"])
    got = oracle.stream_counters(plain, ex, steps, dv)
    assert (got == ref).all(), (got[:12], ref[:12])
    assert ref[-1, 1] > 0                                        # the fixtures do contain block errors


def test_doped_positions_are_known_in_the_reference_stream():
    """periodic doping (is_position_doped_streaming, BP_FULL.c:1589): {5,7,9} => positions = 5,7,9 mod 10 carry no erasure"""
    chan = Z["c1_chan"]
    for p in range(chan.shape[0]):
        doped = oracle.is_position_doped_streaming(p, [5, 7, 9])
        assert doped == (p % 10 in (5, 7, 9))
        if doped:
            assert not chan[p].any()
    assert chan[[p for p in range(chan.shape[0]) if p % 10 not in (5, 7, 9)]].any()


@pytest.mark.parametrize("name", CASES)
def test_product_stream_counters_match_the_reference(name):
    """host logic of fl_scaling_sc_ldpc_b200.streaming (counters, periodic doping, bits / blocks generated) fed with the
    oracle's per-position counts"""
    from fl_scaling_sc_ldpc_b200 import streaming
    dv, dc, L, defM, W, steps, vn_cn, chan, ref = case(name)
    npos, V = vn_cn.shape[0], vn_cn.shape[1]
    g = oracle.Graph(vn_cn.reshape(-1, dv), npos, V, defM, dv, dc)
    o = oracle.decode_bp_sw(g, chan.reshape(-1).astype(np.int32), W, 10 ** 9, 0, square=0, is_term=1)
    plain, ex = oracle.position_counts(g, o["erased"])
    doped = list(Z[name + "_doped"])
    c = streaming.stream_counters(plain, ex, steps, dv, doped, V)
    got = np.stack([c["erasures_pos"], c["num_blocks_err"], c["num_erasures_exp"], c["num_blocks_err_exp"]], axis=1)
    assert (got == ref).all()
    # main_streaming's generated-bit accounting (BP_FULL.c:2017-2026)
    nb = sum(1 for pos in range(steps) if pos - dv + 1 >= 0 and not streaming.is_position_doped_streaming(pos - dv + 1, doped))
    assert c["num_blocks_generated"][-1] == nb and c["num_bits_generated"][-1] == nb * V
    assert (c["num_erasures"] == np.cumsum(ref[:, 0])).all()
