#!/usr/bin/env python3
"""BASELINE config 5 throughput: doped (4,8) SC-LDPC, L=50, M = 10^5 -- peeling trajectories with the fused variance
accumulators and BP trajectories with the device-side moment accumulators.  Under torchrun every rank decodes its own graph
ids; the accumulators are all-reduced at the end.  Prints one JSON line (rank 0)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_scaling_sc_ldpc_b200 as eng
from fl_scaling_sc_ldpc_b200 import dist as D, peeling_decoding as pdx

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=100000)
ap.add_argument("--peel-frames", type=int, default=64)
ap.add_argument("--bp-batches", type=int, default=2)
ap.add_argument("--cap", type=int, default=500)
a = ap.parse_args()
rank, world = D.init_from_env()
dv, dc, L, M, e, doped = 4, 8, 50, a.M, 0.48, [24, 25]
ens = eng.Ensemble(dv, dc, L, M)
out = {"M": M, "n_gpus": world, "doped_positions": doped}
# ---- peeling: r1 trajectories + (ssquares, counts) against the batch mean as the theory curve
cns, num_positions, total_size, steps = pdx._peel_geometry(e, dv, dc, L, M, False)
fb = eng.FrameBatch(ens, 1, a.peel_frames)
fb.generate_graphs(9, first_graph_id=rank).generate_erasures(e, 10, first_graph_id=rank, doping_points=doped)
torch.cuda.synchronize()
t0 = time.perf_counter()
r1, rec, ner = pdx.peel_batch(ens, fb, total_size, steps, 11, rank * a.peel_frames)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
theory = r1[0].double().mean(dim=0).cpu().numpy()
ssq, cnt = pdx.calc_nu_chunk_device([r1[0]], np.maximum(theory, 1e-9) * (theory > 0), M)
fr = D.allreduce_counters([a.peel_frames])[0]
tmax = D.allreduce_max(dt, device=fb.device if world > 1 else None)
out["peeling"] = {"frames": int(fr), "steps_per_frame": steps, "frames_per_s": fr / tmax, "peel_steps_per_s": fr * steps / tmax,
                  "lost_fraction_mean": float(((ner - rec).double() / ((L - len(doped)) * M)).mean().item()), "variance_points": int((cnt > 0).sum())}
del fb, r1
torch.cuda.empty_cache()
# ---- BP trajectories (truncated code like the notebook's files) with on-device moments
fb = eng.FrameBatch(ens, 1, 1024, 16)
fb.generate_graphs(19, first_graph_id=rank)
acc = None
torch.cuda.synchronize()
t0 = time.perf_counter()
its = 0
for b in range(a.bp_batches):
    fb.generate_erasures(0.46, 20, first_graph_id=rank, doping_points=doped, first_frame=b * 1024)
    res, erased, rows, launched = eng.decode_bp_full(fb, a.cap, False, trajectory=True, max_rows=a.cap, collect=False)
    acc = eng.engine.trajectory_moments(fb, res[0], rows, acc)
    its += int(res[0].sum().item())
torch.cuda.synchronize()
dt = time.perf_counter() - t0
m = D.allreduce_counters(acc.cpu().numpy(), device=fb.device if world > 1 else None)
tot_its = D.allreduce_counters([its])[0]
tmax = D.allreduce_max(dt, device=fb.device if world > 1 else None)
frames = world * a.bp_batches * 1024
mean_dvn = m[:, 2] / np.maximum(m[:, 1], 1)
out["bp_trajectories"] = {"frames": frames, "frames_per_s": frames / tmax, "edge_updates_per_s": float(tot_its) * 2 * ens.E / tmax,
                          "mean_iterations": float(tot_its) / frames, "mean_dvn_steady_state_40_200": float(mean_dvn[40:200].mean()),
                          "mean_dvn_over_M": float(mean_dvn[40:200].mean() / M)}
if rank == 0:
    print(json.dumps(out))
D.shutdown()
