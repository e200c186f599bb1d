"""Body of tests/test_multi_rank_gpu.py: every drop-in driver once, writing its files into argv[1] (rank 0 writes)."""
import os
import pickle
import sys

import numpy as np

from fl_scaling_sc_ldpc_b200 import bp_cli, dist as D
from fl_scaling_sc_ldpc_b200 import peeling_decoding as pdx

out = sys.argv[1]
rank, world = D.init_from_env()
pdx.set_seed(77)

# 1. ber_sim.py: two eps points, one cut by max_fuckups, one by num_repeats
pdx.main_simulate_sc_ldpc(["ber_sim", os.path.join(out, "ber.txt"), "4", "8", "12", "64", "[0.47, 0.44]", "T", "U", "B", "NTB", "700", "40", "[]"])
# 2. the three C executables
common = ["--L", "12", "--M", "32", "--outdir", out, "--seed", "5", "--points", "2", "--eps-ini", "0.47", "--eps-delta", "0.03",
          "--frames-per-graph", "32", "--graphs-per-batch", "2"]
bp_cli.run("bp_lim_iter", ["1", "0", "0", "40"] + common + ["--min-frame-err", "25", "--max-frames", "600"])
bp_cli.run("sw_lim_iter", ["2", "4", "0", "6", "12"] + common + ["--min-frame-err", "25", "--max-frames", "600"])
bp_cli.run("bp_traj", ["3", "0", "0", "30", "1"] + common + ["--min-frame-err", "1000000", "--max-frames", "100"])
from fl_scaling_sc_ldpc_b200 import streaming
streaming.main_streaming(["4", "5", "0", "--L", "16", "--M", "16", "--eps-ini", "0.46", "--eps-delta", "0.02", "--points", "2", "--segment", "60",
                          "--max-blocks-err", "30", "--max-blocks", "20000", "--outdir", out])
# 3. peeling trajectories + the variance driver
_, r1, plrs = pdx.simulate_peeling_decoder_ldpc(0.45, 4, 8, 12, 64, False, False, 37, seed=11)
# frames_per_graph > 1: every rank must start on a graph boundary (frame F is graph F // fpg, lane F % fpg)
_, r1g, plrsg = pdx.simulate_peeling_decoder_ldpc(0.45, 4, 8, 12, 64, False, False, 37, seed=11, frames_per_graph=8)
if rank == 0:
    with open(os.path.join(out, "peel.pkl"), "wb") as f:
        pickle.dump((np.asarray(r1), np.asarray(plrs)), f)
    with open(os.path.join(out, "peel_fpg8.pkl"), "wb") as f:
        pickle.dump((np.asarray(r1g), np.asarray(plrsg)), f)
    with open(os.path.join(out, "theory.in"), "wb") as f:
        pickle.dump((np.asarray(r1, np.float64).mean(axis=0),), f)
if world > 1:
    import torch.distributed as dist
    dist.barrier()
pdx.main_simulate_variance(["simulate_variance", os.path.join(out, "var.pkl"), "4", "8", "12", "64", "0.45", "N", "U", "60", "20",
                            os.path.join(out, "theory.in")])
D.shutdown()
