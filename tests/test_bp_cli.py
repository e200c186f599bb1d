"""Writers / counters of the CLI drop-ins: formats from SURVEY.md App. B (CPU), end-to-end runs (GPU)."""
import os

import numpy as np
import pytest

from fl_scaling_sc_ldpc_b200 import bp_cli


def test_result_row_matches_published_file_format():
    """the row layout of sim_data/error_rates/SC_LDPC_4_8_L50_M500_BP_Full_175it_BEC.dat (risultati, BP_FULL.c:499-515)"""
    c = dict(users_err=257379, frame_err=20, block_err=692, users_err_exp=173, frame_err_exp=20, block_err_exp=20, frame_errP1=0)
    row = bp_cli.result_row(0.48, 50000, 50, 20, c)
    # the reference printed exactly this row for these counters (SURVEY.md App. C probe)
    assert row == "0.480000 2.573790e-01 1.000000e+00 6.920000e-01 1.730000e-04 1.000000e+00 2.000000e-02 50000 50 20 257379 20 692 173 20 20\n"
    assert bp_cli.HEADER.split()[:4] == ["p", "BER", "FER", "BLER"] and len(bp_cli.HEADER.split()) == 16


def test_plr_computation_rules():
    c = bp_cli.new_counters()
    bp_cli.account(c, 0, 0, 0, 0)
    bp_cli.account(c, 5, 2, 0, 0)          # erased but fully expurgated
    bp_cli.account(c, 7, 3, 7, 1, 4)
    assert c == dict(users_err=12, frame_err=2, block_err=5, users_err_exp=7, frame_err_exp=1, block_err_exp=1, frame_errP1=1)


def test_trajectory_text_format():
    rows = np.array([[132, 656, 0], [56, 52, 0], [30, 13, 10]])
    assert bp_cli.trajectory_text(rows, 3) == "0\t132\t656\t0\n1\t56\t52\t0\n2\t30\t13\t10\n\n"


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    common = ["--L", "10", "--M", "25", "--points", "2", "--max-frames", "40", "--min-frame-err", "5", "--seed", "4",
              "--frames-per-graph", "16", "--graphs-per-batch", "2", "--outdir", str(tmp_path)]
    assert bp_cli.bp_lim_iter(["3", "0", "0", "50"] + common + ["--eps-ini", "0.50", "--eps-delta", "0.05"]) == 0
    p = tmp_path / "SC_LDPC_4_8_L10_M25_BP_SW0_50it_Random_BLER_3.dat"
    lines = p.read_text().splitlines()
    assert lines[0] == bp_cli.HEADER.strip() and len(lines) == 3
    v = lines[1].split()
    assert v[0] == "0.500000" and int(v[7]) == 500 and int(v[8]) == 10 and 1 <= int(v[9]) <= 40
    assert int(v[11]) <= 5 or int(v[9]) == 40                               # stopped at the fifth frame error
    assert bp_cli.sw_lim_iter(["0", "4", "1", "3", "9", "2"] + common) == 0
    assert (tmp_path / "SC_LDPC_4_8_L10_M25_BP_SW4_3it_9init_Random_BLER_0.dat").exists()
    assert bp_cli.bp_traj(["1", "0", "0", "60", "0"] + common + ["--points", "1", "--eps-ini", "0.44"]) == 0
    t = (tmp_path / "trajectories_0.4400_truncated_SC_LDPC_4_8_L10_M25_BP_Full_60it_Random_BLER_1.dat").read_text()
    frames = t.strip("\n").split("\n\n")
    assert 1 <= len(frames) <= 40
    # BP_TRAJ.c's main_terminated writes the trajectories file only (risultati is commented out, :2176)
    assert not [f for f in os.listdir(tmp_path) if "_60it_" in f and not f.startswith("trajectories_")]
    first = [ln.split("\t") for ln in frames[0].split("\n")]
    assert [int(r[0]) for r in first] == list(range(len(first))) and all(len(r) == 4 for r in first)
    assert int(first[0][2]) > 0 and int(first[-1][3]) <= 10


@pytest.mark.gpu
def test_bp_lim_iter_stream_path_writes_the_same_file(tmp_path):
    """bp_lim_iter decodes each graph's frames as a capped stream with lane recycling when asked to (and by default from
    1024 frames per graph): same Philox realisations, so the result file is byte-identical to the synchronous path"""
    files = {}
    for mode in ("off", "on"):
        out = tmp_path / mode
        out.mkdir()
        bp_cli.run("bp_lim_iter", ["3", "0", "1", "12", "4", "--L", "12", "--M", "24", "--outdir", str(out), "--seed", "9", "--points", "2",
                                   "--eps-ini", "0.48", "--eps-delta", "0.03", "--frames-per-graph", "300", "--graphs-per-batch", "2",
                                   "--min-frame-err", "60", "--max-frames", "1500", "--stream", mode])
        (name,) = os.listdir(out)
        files[mode] = (name, open(out / name).read())
    assert files["on"] == files["off"]
    rows = files["on"][1].splitlines()
    assert len(rows) == 3 and float(rows[1].split()[2]) > 0          # header + 2 points, some frame errors at eps = 0.48


def test_compat_argv_reads_doped_positions_from_max_it_on():
    """--compat-argv: doped_positions[i] = atoi(argv[4+i]) (BP_FULL.c:2083-2091) -- MAX_IT first, then what follows; option
    values in between must not shift it"""
    from fl_scaling_sc_ldpc_b200 import bp_cli as b
    a = b._parser("bp_lim_iter").parse_args(["--M", "100", "0", "50", "2", "20", "7", "--compat-argv"])
    tail = [a.max_it] + list(a.doped)
    assert tail[: a.num_doped] == [20, 7]
