#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path: full BP over the BEC, (4,8) SC-LDPC, L=50, M=10000.

    python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU decoder on the host cores

A *step* decodes 4 graph realisations (one per eps of the 0.46..0.49 sweep of BASELINE config 2) x B frames each
(default B = 16384, 1024 bit-sliced lanes per graph), full BP with unlimited iterations until every frame has stalled
or finished.  Default mode "stream": a lane whose frame has stopped is re-armed with the graph's next channel
realisation (scldpc_bp_stream); mode "batch": B = lanes, every lane decodes one frame (scldpc_bp_full).  Throughput counts USEFUL
work only: edge-updates = sum over frames of (iterations that frame executed) * 2E, the same formula as for the CPU
(SURVEY.md section 8d); iterations a finished frame rides along for do not count.

value : batches resident in HBM, decode + per-frame counters on the device.
e2e   : the same step through the host-buffer C-ABI call scldpc_decode_host: graph tables and bit-sliced channel
        words start in pinned HOST memory, results end in host memory; H2D / D2H inside the timed region.
roofline : one flooding iteration of the stream decoder (node-state sweeps: ns_cn_kernel + ns_x_kernel), algorithmic bytes
        B_alg = (4E+n)/8 per useful frame-iteration (the message formulation's figure) over the CUDA-event time
        of sampled launches inside the timed region, against MEASURED_PEAKS.json's HBM copy bandwidth; the DRAM bytes ncu
        measured and the node-state minimum are reported beside it.  SCLDPC_STREAM_NODE=0 / --mode batch run the
        message-passing sweeps, for which the VN sweep is reported as before.
cpu_baseline : the unmodified reference decodeBP (oracle/_ref, compiled from /root/reference) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DV, DC, L, M = 4, 8, 50, 10000
EPS_SWEEP = [0.46, 0.47, 0.48, 0.49]
N_WORDS = 16
FRAMES_PER_GRAPH = 64 * N_WORDS
E_EDGES = L * M * DV
N_VNS = L * M
N_CNS = (L + DV - 1) * (M * DV // DC)
WORKLOAD = ("full BP unlimited iterations, (4,8) SC-LDPC terminated L=50 M=10000, BEC eps sweep "
            "{0.46,0.47,0.48,0.49}: 4 graphs x B frames per step")
FRAMES_PER_STREAM = 16384
HARVEST_EVERY = 0        # 0: the library adapts the period to the iterations per frame it observes
CAP_LO = 2   # reference arm: iterations of the shorter of the two capped runs
METRIC = "edge-updates/s (frames/s alongside), (4,8) SC-LDPC L=50 M=10000 BEC BP"


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own decodeBP on the host cores
# ------------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One frame on one core: generate_code + channel_doped, then decodeBP twice on the same frame, capped at CAP_LO
    and at CAP_LO+cap flooding iterations.  The difference of the two times is the cost of `cap` iterations of the
    reference's sweeps over the full-size frame; the common part (message initialisation) is subtracted out.  The
    post-decoding expurgation scan -- quadratic in the erasures a capped run leaves behind, negligible for a
    converged frame -- is skipped by passing decodeBP's L argument (used only as that scan's range) as 0.  Uses BP_TRAJ.c's decodeBP (same loops as BP_FULL.c's plus one fprintf per iteration) so that
    the number of executed iterations is known."""
    seed, eps, cap = args
    from oracle import ref_driver as rd
    r = rd.get("traj", DV, DC, L, M * DV // DC)
    r.srandom(seed)
    r.reset_perm()
    r.generate_code()
    r.channel_doped(eps)
    t0 = time.perf_counter()
    a = r.decode_bp(CAP_LO, 1, exp_positions=0)
    t1 = time.perf_counter()
    b = r.decode_bp(CAP_LO + cap, 1, exp_positions=0)
    t2 = time.perf_counter()
    return len(b["rows"]) - len(a["rows"]), (t2 - t1) - (t1 - t0)


def _port_worker(args):
    """Fallback when oracle/_ref is absent: the oracle port (same algorithm, table-driven reverse-edge lookup)."""
    seed, eps, cap = args
    import oracle
    oracle.srandom(seed)
    g, _ = oracle.generate_code(L, M, M * DV // DC, DV, DC)
    ch = oracle.channel_doped(g.n, eps, M)
    t0 = time.perf_counter()
    a = oracle.decode_bp(g, ch, CAP_LO, 1, max_rows=1)
    t1 = time.perf_counter()
    b = oracle.decode_bp(g, ch, CAP_LO + cap, 1, max_rows=1)
    t2 = time.perf_counter()
    return b["iters"] - a["iters"], (t2 - t1) - (t1 - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from oracle import ref_driver as rd
    cores = os.cpu_count() or 1
    have_ref = rd.available("traj", DV, DC, L, M * DV // DC)
    worker, kind = (_ref_worker, "reference") if have_ref else (_port_worker, "port")
    cap = args.sample_iters
    ctx = mp.get_context("fork")
    per_step = []
    with ctx.Pool(cores) as pool:
        for s in range(args.warmup + args.steps):
            jobs = [(1000003 * (s + 1) + i, EPS_SWEEP[i % len(EPS_SWEEP)], cap) for i in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(worker, jobs)
            wall = time.perf_counter() - t0
            decode_wall = max(dt for _, dt in res)
            if s >= args.warmup:
                per_step.append((sum(it for it, _ in res), decode_wall, wall))
    iters = sum(p[0] for p in per_step)
    t = sum(p[1] for p in per_step)
    eu = iters * 2.0 * E_EDGES / t
    sample = (f"{cores} processes x 1 full-size frame per step; per frame the unmodified decodeBP runs capped at {CAP_LO} and at "
              f"{CAP_LO}+{cap} flooding iterations and the time difference is charged to the {cap} extra iterations "
              f"(every iteration of the reference sweeps all nodes, so cost per iteration is constant; graph/channel "
              f"generation is not timed and the expurgation scan is skipped via decodeBP's L argument)")
    line = {
        "impl": "reference", "metric": METRIC, "value": eu, "unit": "edge-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, len(per_step)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 (0/1 messages)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "frame_iterations_per_s": iters / t,
        "cpu_baseline": {"value": eu, "unit": "edge-updates/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": eu, "unit": "edge-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".clocks.csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # timed before CUDA is initialised in this process (the workers fork)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "0",
                                "--sample-iters", str(args.sample_iters)], capture_output=True, text=True, timeout=900)
            cpu_baseline = json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": "edge-updates/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"failed: {e!r}"}

    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    import fl_scaling_sc_ldpc_b200 as eng
    from fl_scaling_sc_ldpc_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    ens = eng.Ensemble(DV, DC, L, M)
    stream_mode = args.mode == "stream"
    lanes = FRAMES_PER_GRAPH                                     # bit-sliced lanes per graph
    B = args.frames_per_graph if stream_mode else lanes         # frames decoded per graph and step
    G = len(EPS_SWEEP) * args.graphs_per_eps
    eps = [e for e in EPS_SWEEP for _ in range(args.graphs_per_eps)]
    eps_arr = np.asarray(eps, np.float64)

    # two resident batches, alternated, each an independent draw: graph ids are global, so the realisations (and
    # therefore the results) do not depend on how many GPUs share the job
    batches = []
    for bidx in range(2):
        gid0 = (rank * 2 + bidx) * G
        fb = eng.FrameBatch(ens, G, lanes, N_WORDS, device=dev)
        fb.generate_graphs(seed=args.seed, first_graph_id=gid0)
        if not stream_mode:
            fb.generate_erasures(eps, seed=args.seed + 1, first_graph_id=gid0)
        fb.gid0 = gid0
        batches.append(fb)
    torch.cuda.synchronize()

    counters = torch.zeros(4, dtype=torch.int64, device=dev)   # frame-iterations, frames, frame errors, bit errors

    iters_launched = [0]                                       # sweep pairs launched by this rank in the timed region

    def step(i, acc=True):
        fb = batches[i % 2]
        if stream_mode:
            res, launched = eng.decode_bp_stream(fb, B, eps, args.seed + 1, first_graph_id=fb.gid0, harvest_every=args.harvest_every, collect=False)
            it, resid = res[0].to(torch.int64), res[1].to(torch.int64)
        else:
            res, erased, rows, launched = eng.decode_bp_full(fb, eng.UNLIMITED, True, collect=False)
            it, resid = res[0, :, :lanes].to(torch.int64), res[1, :, :lanes].to(torch.int64)
        if acc:
            iters_launched[0] += int(launched)
            counters.add_(torch.stack([it.sum(), torch.tensor(it.numel(), device=dev), (resid > 0).sum(), resid.sum()]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, acc=False)
    barrier()

    # ---- timed region: K steps, device-resident inputs ---------------------------------------------------------
    sample_every = 13       # co-prime with the harvest periods in use, so sampled launches are representative
    _lib.check(lib.scldpc_profile_begin(sample_every, 16384))
    lib.scldpc_launch_count(1)
    clocks = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.scldpc_launch_count(1)
    clk = clocks.stop()
    cap = 16384
    ns = ctypes.c_int(0)
    it_idx = (ctypes.c_int * cap)()
    cn_ms = (ctypes.c_float * cap)()
    vn_ms = (ctypes.c_float * cap)()
    _lib.check(lib.scldpc_profile_end(ctypes.byref(ns), it_idx, cn_ms, vn_ms, cap))

    local_frame_iters = int(counters[0].item())
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = counters.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)      # the only collective of the job: final counters over NVLink
    ms = float(tmax.item())
    frame_iters, frames, ferr, berr = (int(x) for x in tot.tolist())
    value = frame_iters * 2.0 * E_EDGES / (ms * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    # algorithmic bytes per launch = useful frame-iterations of the timed region / sweeps launched x bytes per
    # frame-iteration of that sweep; average launch duration from the CUDA-event samples (every 7th iteration)
    roof = kernels = None
    if rank == 0 and ns.value > 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        n_s = ns.value
        cn_avg = float(np.mean(cn_ms[:n_s])) * 1e-3
        vn_avg = float(np.mean(vn_ms[:n_s])) * 1e-3
        n_iter_launches = iters_launched[0] or n_s * sample_every  # iterations launched by this rank in the timed region
        fi_per_launch = local_frame_iters / max(1, n_iter_launches)
        cn_bytes = fi_per_launch * (2 * E_EDGES) / 8.0
        vn_bytes = fi_per_launch * (2 * E_EDGES + N_VNS) / 8.0
        node_state = os.environ.get("SCLDPC_STREAM_NODE" if stream_mode else "SCLDPC_FULL_NODE", "1") != "0"
        k_cn, k_x = ("ns_cn_kernel<4,8>", "ns_x_kernel") if stream_mode else ("bpw_cn_node_kernel<4,8>", "bpw_vn_node_kernel")
        tj = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            try:
                tj = json.load(open(tpath))
            except Exception:
                tj = {}
        if node_state:
            # Node-state sweeps (bp_node_kernels.cu): the first timed kernel is the CN sweep (gathers E rows of x, scatters
            # the resolutions), the second the sequential state pass.  SURVEY 8(d): the denominator of record stays the
            # message formulation's B_alg = (4E+n)/8 B per frame-iteration; the DRAM bytes ncu measured and the
            # node-state minimum (2n+nk)/8 are reported next to it, so a fraction above 1 is explained, not hidden.
            it_bytes = fi_per_launch * (4 * E_EDGES + N_VNS) / 8.0
            ach = it_bytes / (cn_avg + vn_avg) / 1e9
            tn = tj.get("node_state", {}) if stream_mode else {}
            traffic = tn.get("iteration_dram_bytes_per_launch")
            roof = {"bound": "hbm", "kernel": k_cn + " + " + k_x + " (one flooding iteration)", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_note": tn.get("note"), "peak_source": peak_src, "launches_sampled": n_s,
                    "avg_launch_ms": 1e3 * (cn_avg + vn_avg), "frame_iterations_per_launch": fi_per_launch,
                    "algorithmic_bytes": "(4E+n)/8 B = 1.0625 MB per useful frame-iteration (message formulation, SURVEY 8d: the denominator of record)",
                    "node_state_min_bytes_per_launch": fi_per_launch * (2 * N_VNS + N_CNS) / 8.0,
                    "explanation": "the node-state sweeps keep 1 bit per VN and frame instead of 2 bits per edge, and their gathers are served "
                                   "by L2 (a band of dv positions); HBM traffic per iteration is a fraction of B_alg, so frac > 1 is expected. "
                                   "The kernels are bound by L2->SM sector bandwidth: see l2 below",
                    "l2": {"bytes_per_launch": tn.get("cn_l2_read_bytes_per_launch"), "note": tn.get("l2_note")}}
            pct = lambda a, q: float(np.percentile(np.asarray(a[:n_s], dtype=np.float64), q))
            kernels = {k_cn: {"avg_launch_ms": 1e3 * cn_avg, "p10_p50_p90_ms": [pct(cn_ms, 10), pct(cn_ms, 50), pct(cn_ms, 90)],
                                             "share_of_iteration": cn_avg / (cn_avg + vn_avg),
                                             "dram_bytes_per_launch": tn.get("cn_dram_bytes_per_launch")},
                       k_x: {"avg_launch_ms": 1e3 * vn_avg, "p10_p50_p90_ms": [pct(vn_ms, 10), pct(vn_ms, 50), pct(vn_ms, 90)],
                                       "share_of_iteration": vn_avg / (cn_avg + vn_avg),
                                       "dram_bytes_per_launch": tn.get("x_dram_bytes_per_launch")}}
        else:
            traffic = tj.get("vn_sweep_dram_bytes_per_launch")
            traffic_note = None
            if tj:
                traffic_note = ("ncu --set full capture of one launch with %d graph(s) decoding (%s); the average launch of "
                                "the timed region carries %.0f active frames" % (tj.get("graphs_decoding_in_captured_launch", 1),
                                                                               tj.get("algorithmic_bytes_same_launch", ""), fi_per_launch))
            vn_name = "bp_vn_stream_kernel<4>" if stream_mode else "bp_vn_wave_kernel<4,false>"
            ach = vn_bytes / vn_avg / 1e9
            roof = {"bound": "hbm", "kernel": vn_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "launches_sampled": n_s, "avg_launch_ms": 1e3 * vn_avg,
                    "frame_iterations_per_launch": fi_per_launch,
                    "algorithmic_bytes": "(2E+n)/8 B per useful frame-iteration (reads E c2v bits + n channel bits, writes E v2c bits)"}
            ach_c = cn_bytes / cn_avg / 1e9
            kernels = {"bp_cn_wave_kernel<8,false>": {"achieved": ach_c, "frac": ach_c / peak, "unit": "GB/s", "avg_launch_ms": 1e3 * cn_avg,
                                                      "algorithmic_bytes": "2E/8 B per useful frame-iteration"},
                       "both_sweeps": {"achieved": (cn_bytes + vn_bytes) / (cn_avg + vn_avg) / 1e9,
                                       "frac": (cn_bytes + vn_bytes) / (cn_avg + vn_avg) / 1e9 / peak, "unit": "GB/s"}}

    # ---- e2e: host buffers through the C ABI ---------------------------------------------------------------------
    h_vn = [fb.vn_cn.cpu().pin_memory() for fb in batches]
    dims = batches[0].dims
    if stream_mode:
        outs = [torch.zeros((G, B), dtype=torch.int32).pin_memory() for _ in range(5)]
        cfgs = []
        for fb in batches:
            cfgs.append(_lib.StreamCfg(B, args.harvest_every, _lib.F_TERMINATED, 0, 0, eps_arr.ctypes.data_as(ctypes.c_void_p).value, None, None, None,
                                       args.seed + 1, fb.gid0))
        h2d = int(h_vn[0].numel() * 4)
        d2h = int(5 * G * B * 4)
        api = "scldpc_stream_host (graph tables in pinned host memory; channel realisations drawn on the device)"

        def e2e_step(i):
            _lib.check(lib.scldpc_stream_host(ctypes.byref(dims), ctypes.c_void_p(h_vn[i % 2].data_ptr()), ctypes.byref(cfgs[i % 2]),
                                              *[ctypes.c_void_p(o.data_ptr()) for o in outs], None))
            return int(outs[0].sum().item())
    else:
        h_ch = [fb.chan.cpu().pin_memory() for fb in batches]
        outs = [torch.zeros((G, lanes), dtype=torch.int32).pin_memory() for _ in range(5)]
        flags = _lib.F_TERMINATED | _lib.F_CHAN_PACKED
        h2d = int(h_vn[0].numel() * 4 + h_ch[0].numel() * 8)
        d2h = int(6 * G * lanes * 4)
        api = "scldpc_decode_host (graph tables + bit-sliced channel words in pinned host memory)"

        def e2e_step(i):
            _lib.check(lib.scldpc_decode_host(ctypes.byref(dims), ctypes.c_void_p(h_vn[i % 2].data_ptr()),
                                              ctypes.c_void_p(h_ch[i % 2].data_ptr()), 0, 0, 0, flags,
                                              *[ctypes.c_void_p(o.data_ptr()) for o in outs], None, None, None, 0))
            return int(outs[0].sum().item())

    for i in range(min(2, args.warmup)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e_iters = 0
    for i in range(args.steps):
        e_iters += e2e_step(i)          # the call returns after its D2H copy has completed
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    et = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    ei = torch.tensor([e_iters], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(ei, op=dist.ReduceOp.SUM)
    e2e = {"value": float(ei.item()) * 2.0 * E_EDGES / float(et.item()), "unit": "edge-updates/s",
           "frames_per_s": world * args.steps * G * B / float(et.item()),
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api}

    if rank == 0:
        need = lib.scldpc_bp_stream_workspace_bytes(ctypes.byref(dims)) if stream_mode else batches[0].workspace(_lib.F_TERMINATED).numel()
        ws_mb = need / 2 ** 20
        line = {
            "metric": METRIC, "value": value, "unit": "edge-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (64 bit-sliced frames per word)", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace("B frames", f"{B} frames"), "mode": args.mode, "sweeps": ("node-state" if os.environ.get("SCLDPC_STREAM_NODE" if stream_mode else "SCLDPC_FULL_NODE", "1") != "0" else "message passing"), "frames_per_step_per_gpu": G * B,
                       "lanes_per_graph": lanes, "n_words": N_WORDS,
                       "l2": f"inputs larger than L2: {ws_mb:.0f} MB of decoder state per batch, two batches alternated",
                       "seed": args.seed},
            "frames_per_s": frames / (ms * 1e-3),
            "frame_iterations_per_s": frame_iters / (ms * 1e-3),
            "mean_iterations_per_frame": frame_iters / max(1, frames),
            "frame_error_rate": ferr / max(1, frames), "bit_error_rate": berr / max(1, frames) / N_VNS,
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roof, "kernels": kernels,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    global N_WORDS, FRAMES_PER_GRAPH, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=0x5C1D9C)
    ap.add_argument("--graphs-per-eps", type=int, default=1)
    ap.add_argument("--sample-iters", type=int, default=10, help="reference arm: flooding iterations per sampled frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--n-words", type=int, default=N_WORDS, help="64-bit lane words per node (lanes per graph / 64)")
    ap.add_argument("--mode", default="stream", choices=["stream", "batch"],
                    help="stream: lane recycling over --frames-per-graph frames per graph; batch: one frame per lane")
    ap.add_argument("--frames-per-graph", type=int, default=FRAMES_PER_STREAM)
    ap.add_argument("--harvest-every", type=int, default=HARVEST_EVERY, help="stream mode: iterations between harvests of finished frames")
    args = ap.parse_args()
    N_WORDS = args.n_words
    FRAMES_PER_GRAPH = 64 * N_WORDS
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
