"""Drop-in for the reference's streaming decoder (``main_streaming`` + ``decodeBP_SW_circular``, BP_FULL.c:1403-1500,
1934-2054; compiled out upstream behind ``#undef CIRCULAR``, BP_FULL.c:34).

The reference decodes an unbounded chain inside a ring buffer of L positions: at step ``pos`` it runs flooding BP with
unlimited iterations over the classical window (VN positions [pos-dv+1, pos+W), CN positions [pos, pos+W)), decides VN
position pos-dv+1, expurgates size-two stopping sets of position pos-2*dv+1, and generates a new position L/2 ahead.
The ring only bounds memory: check nodes beyond the window send erasures whether or not they exist yet, so the decisions
are those of the classical window decoder (``decodeBP_SW`` classical, BP_FULL.c:627) with unlimited per-window
iterations on the unrolled chain.  ``tests/test_stream_decoder_golden.py`` pins this against the reference built with
CIRCULAR defined, step by step.

Here a stream is decoded in segments of ``segment`` positions on the GPU (each bit lane is an independent channel
realisation over the segment's graph); every segment starts like the reference's stream does, with a terminated head.
Counters are accumulated position by position in main_streaming's order, so the stop rule (1000 expurgated block
errors or 10^6 expurgated blocks, BP_FULL.c:2033) cuts at the same position a sequential run would.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import engine
from ._lib import F_EXP_ALL, F_TERMINATED


def is_position_doped_streaming(pos: int, doped_positions) -> bool:
    """Periodic doping (BP_FULL.c:1589-1612): the period is the last doped position + 1."""
    doped_positions = list(doped_positions)
    if not doped_positions:
        return False
    period = doped_positions[-1] + 1
    r = pos % period
    if r < doped_positions[0]:
        return False
    return r in doped_positions


def stream_counters(plain: np.ndarray, ex: np.ndarray, n_steps: int, dv: int, doped_positions=(), vns_pos: int = 1):
    """main_streaming's counters after ``n_steps`` decode steps (BP_FULL.c:2015-2031 with decodeBP_SW_circular
    :1483-1497) from the per-position counts of one stream: ``plain[p]`` erased VNs and ``ex[p]`` expurgated erased VNs
    (pos_cnt - 2*pos_pairs) of absolute position p.  Step ``pos`` decides position pos-dv+1 and expurgates pos-2*dv+1.
    Returns a dict of int arrays of length n_steps (running values after each step)."""
    q = np.arange(n_steps) - dv + 1
    q2 = np.arange(n_steps) - 2 * dv + 1
    dec = np.where(q >= 0, plain[np.clip(q, 0, None)], 0)
    exv = np.where(q2 >= 0, np.maximum(ex[np.clip(q2, 0, None)], 0), 0)
    nd = np.array([not is_position_doped_streaming(int(p), doped_positions) for p in range(n_steps)], bool)
    gen = np.where(q >= 0, nd[np.clip(q, 0, None)], False)
    gen2 = np.where(q2 >= 0, nd[np.clip(q2, 0, None)], False)
    return dict(erasures_pos=dec, num_erasures=np.cumsum(dec), num_blocks_err=np.cumsum(dec > 0),
                num_erasures_exp=np.cumsum(exv), num_blocks_err_exp=np.cumsum(exv > 0),
                num_bits_generated=np.cumsum(gen) * vns_pos, num_blocks_generated=np.cumsum(gen),
                num_bits_generated_exp=np.cumsum(gen2) * vns_pos, num_blocks_generated_exp=np.cumsum(gen2))


def decode_segment(ens: engine.Ensemble, W: int, eps: float, doped_positions, n_graphs: int, n_frames: int, seed: int,
                   first_graph_id: int = 0, fb: engine.FrameBatch | None = None):
    """Decodes n_graphs x n_frames independent stream segments of ens.L positions.  Returns (plain, ex): int32
    [n_graphs][L][n_frames] erased and expurgated erased VNs per absolute position."""
    if fb is None:
        fb = engine.FrameBatch(ens, n_graphs, n_frames)
        fb.generate_graphs(seed, first_graph_id=first_graph_id)
        doped = [p for p in range(ens.L) if is_position_doped_streaming(p, doped_positions)]
        fb.generate_erasures(eps, seed + 1, first_graph_id=first_graph_id, doping_points=doped)
    engine.decode_bp_window(fb, W, engine.UNLIMITED, 0, square=False, is_term=True, collect=False)
    flags = F_TERMINATED | F_EXP_ALL
    cnt, pairs = engine.position_counts(fb, flags)
    return cnt, cnt - 2 * pairs


def simulate_stream(eps, dv, dc, M, W, doped_positions=(), segment=2000, max_blocks_err=1000, max_blocks=1000000, seed=0x5C1D9C,
                    frames_per_graph=128, graphs_per_batch=2):
    """One epsilon point of main_streaming (BP_FULL.c:1976-2050).  ``M`` = VNs per position.  Returns the thirteen
    numbers of a ``results_circular`` row (BP_FULL.c:547-560) as a dict."""
    ens = engine.Ensemble(dv, dc, segment + W + dv, M)        # positions beyond `segment` only feed the last windows
    tot = dict(num_erasures=0, num_bits_generated=0, num_blocks_err=0, num_blocks_generated=0, num_erasures_exp=0,
               num_bits_generated_exp=0, num_blocks_err_exp=0, num_blocks_generated_exp=0)
    from . import dist as D
    rank, world = D.world()
    gid = 0
    while True:
        # one round = one batch of segments per rank (graph ids are global); the per-position counts are all-gathered and every
        # rank replays the stop rule over the segments in order, so the row does not depend on the number of GPUs
        plain, ex = decode_segment(ens, W, eps, doped_positions, graphs_per_batch, frames_per_graph, seed, gid + rank * graphs_per_batch)
        gid += world * graphs_per_batch
        if world > 1:
            plain = D.allgather_rows(np.asarray(plain)).reshape((-1,) + plain.shape[1:])
            ex = D.allgather_rows(np.asarray(ex)).reshape((-1,) + ex.shape[1:])
        for g in range(plain.shape[0]):
            for f in range(plain.shape[2]):
                c = stream_counters(plain[g, :, f], ex[g, :, f], segment, dv, doped_positions, M)
                # main_streaming tests its stop rule after every decode step
                lim_e = max_blocks_err - tot["num_blocks_err_exp"]
                lim_b = max_blocks - tot["num_blocks_generated_exp"]
                hit = np.flatnonzero((c["num_blocks_err_exp"] >= lim_e) | (c["num_blocks_generated_exp"] >= lim_b))
                k = int(hit[0]) if len(hit) else segment - 1
                for key in tot:
                    tot[key] += int(c[key][k])
                if len(hit):
                    return tot


HEADER = "p BER BLER BER_EXP BLER_EXP bit_err bit_gen block_err block_gen bit_err_exp bit_gen_exp block_err_exp block_gen_exp\n"


def result_row(eps, t):
    """``results_circular`` row (BP_FULL.c:547-560)."""
    return "%f %e %e %e %e %d %d %d %d %d %d %d %d\n" % (
        eps, t["num_erasures"] / t["num_bits_generated"], t["num_blocks_err"] / t["num_blocks_generated"],
        t["num_erasures_exp"] / t["num_bits_generated_exp"], t["num_blocks_err_exp"] / t["num_blocks_generated_exp"],
        t["num_erasures"], t["num_bits_generated"], t["num_blocks_err"], t["num_blocks_generated"], t["num_erasures_exp"],
        t["num_bits_generated_exp"], t["num_blocks_err_exp"], t["num_blocks_generated_exp"])


def main_streaming(argv=None) -> int:
    """``sw INDEX W NUM_DOPED DOPED_POSITIONS...`` (BP_FULL.c:1934-1963) -> ``SC_LDPC_{dv}_{dc}_L{L}_M{Def_M}_DOP{n}_BP_Stream_SW{W}_
    Random_BLER_{INDEX}.dat`` with one ``results_circular`` row per epsilon.  The reference's compile-time parameters are
    options; ``--L`` is only used in the file name (the ring length does not affect the results)."""
    ap = argparse.ArgumentParser(prog="bp_stream")
    ap.add_argument("index", type=int)
    ap.add_argument("W", type=int)
    ap.add_argument("num_doped", type=int)
    ap.add_argument("doped", type=int, nargs="*")
    ap.add_argument("--dv", type=int, default=4)
    ap.add_argument("--dc", type=int, default=8)
    ap.add_argument("--L", type=int, default=50)
    ap.add_argument("--M", type=int, default=500, help="Def_M: CNs per position")
    ap.add_argument("--eps-ini", type=float, default=0.48)
    ap.add_argument("--eps-delta", type=float, default=0.00125)
    ap.add_argument("--points", type=int, default=26)
    ap.add_argument("--max-blocks-err", type=int, default=1000)        # Def_MaxNumberBlocksError
    ap.add_argument("--max-blocks", type=int, default=1000000)          # Def_MaxNumberBlocksSim
    ap.add_argument("--segment", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=0x5C1D9C)
    ap.add_argument("--outdir", default=".")
    a = ap.parse_args(argv)
    doped = list(a.doped)[: a.num_doped]
    from . import dist as D
    rank, _world = D.init_from_env()
    name = "SC_LDPC_%d_%d_L%d_M%d_DOP%d_BP_Stream_SW%d_Random_BLER_%d.dat" % (a.dv, a.dc, a.L, a.M, a.num_doped, a.W, a.index)
    for sim in range(a.points):
        eps = a.eps_ini - sim * a.eps_delta
        t = simulate_stream(eps, a.dv, a.dc, a.M * a.dc // a.dv, a.W, doped, a.segment, a.max_blocks_err, a.max_blocks, a.seed + 17 * sim)
        if rank != 0:
            continue
        sys.stdout.write("%f %e %e %e %e\n" % (eps, t["num_erasures"] / t["num_bits_generated"], t["num_blocks_err"] / t["num_blocks_generated"],
                                               t["num_erasures_exp"] / t["num_bits_generated_exp"], t["num_blocks_err_exp"] / t["num_blocks_generated_exp"]))
        with open(os.path.join(a.outdir, name), "w" if sim == 0 else "a") as f:
            if sim == 0:
                f.write(HEADER)
            f.write(result_row(eps, t))
    return 0


if __name__ == "__main__":
    sys.exit(main_streaming())
