"""Drop-in for the reference's streaming decoder (``main_streaming`` + ``decodeBP_SW_circular``, BP_FULL.c:1403-1500,
1934-2054; compiled out upstream behind ``#undef CIRCULAR``, BP_FULL.c:34).

The reference decodes an unbounded chain inside a ring buffer of L positions: at step ``pos`` it runs flooding BP with
unlimited iterations over the classical window (VN positions [pos-dv+1, pos+W), CN positions [pos, pos+W)), decides VN
position pos-dv+1, expurgates size-two stopping sets of position pos-2*dv+1, and generates a new position L/2 ahead.
The ring only bounds memory: check nodes beyond the window send erasures whether or not they exist yet, so the decisions
are those of the classical window decoder (``decodeBP_SW`` classical, BP_FULL.c:627) with unlimited per-window
iterations on the unrolled chain.  ``tests/test_stream_decoder_golden.py`` pins this against the reference built with
CIRCULAR defined, step by step.

Here every bit lane is an independent unbounded chain, decoded ``segment`` positions at a time with the decoder state
carried across pieces (``ChainPieces``): a chain is terminated only at its very beginning, like the reference's.
Counters are accumulated position by position in main_streaming's order, so the stop rule (1000 expurgated block
errors or 10^6 expurgated blocks, BP_FULL.c:2033) cuts at a definite step.
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


class ChainPieces:
    """n_graphs x n_frames independent UNBOUNDED chains, decoded ``piece`` positions at a time with the decoder state carried
    from one piece to the next -- the reference's ring buffer (decodeBP_SW_circular, BP_FULL.c:1403-1500; generation L/2
    ahead, main_streaming :2003-2012) walks one chain position by position and never restarts it, so an erasure burst that
    a window leaves behind keeps propagating; decoding the chain in independent segments would give every segment a fresh
    terminated (low-rate) head instead.

    Piece k is the local chain of absolute positions [A, A + Lloc): its first T = 2(dv-1) positions are already decided
    (they complete the check nodes and the size-two stopping-set scan of the positions decided last), the next ``piece``
    positions are decided by this piece's windows, and dv-1+W more positions only feed those windows.  Graph and channel of
    a position are functions of its absolute index (``generate_graphs(first_position=...)``), so the overlap is regenerated
    identically; the two state planes of the node-state window decoder (what the CNs see / what every VN has been told) are
    copied over for the first T+dv-1+W local positions (``scldpc_bp_window_range`` with resume)."""

    def __init__(self, dv, dc, M, W, eps, doped_positions, n_graphs, n_frames, seed, first_graph_id=0, piece=2000, provider=None):
        self.dv, self.dc, self.M, self.W, self.eps = dv, dc, M, W, eps
        self.doped = list(doped_positions)
        self.ms = dv - 1
        self.T = 2 * self.ms
        self.S = int(piece)
        self.C = self.T + self.ms + W                       # local positions whose state is carried
        self.Lloc = self.T + self.S + self.ms + W
        self.ens = engine.Ensemble(dv, dc, self.Lloc, M)
        self.fb = engine.FrameBatch(self.ens, n_graphs, n_frames)
        self.seed, self.gid = seed, first_graph_id
        self.provider = provider                            # tests: (A, Lloc) -> (vn_cn [G][Lloc*M][dv], erased [G][F][Lloc*M])
        self.A = 0                                          # absolute index of local position 0
        self.D = -1                                         # last decided absolute position
        import torch
        shape = (n_graphs, self.ens.n, self.fb.n_words)
        self.x = torch.zeros(shape, dtype=torch.int64, device=self.fb.device)
        self.xb = torch.zeros(shape, dtype=torch.int64, device=self.fb.device)
        self._prev = None
        self.k = 0

    def _load(self):
        fb, A, M = self.fb, self.A, self.M
        if self.provider is not None:
            vn_cn, erased = self.provider(A, self.Lloc)
            fb.set_graphs(vn_cn)
            fb.set_erasures(erased)
            return
        fb.generate_graphs(self.seed, first_graph_id=self.gid, first_position=A)
        doped = [j for j in range(self.Lloc) if is_position_doped_streaming(A + j, self.doped)]
        if A == 0:
            fb.generate_erasures(self.eps, self.seed + 1, first_graph_id=self.gid, doping_points=doped)
        else:
            fb.generate_erasures(self.eps, self.seed + 1, first_graph_id=self.gid, doping_points=doped, first_vn=A * M)

    def next_piece(self):
        """Decides the next ``piece`` positions.  Returns (q0, plain, e0, ex): ``plain`` int32 [G][S][F] erased VNs of the
        absolute positions q0 .. q0+S-1 (the newly decided ones) and ``ex`` int32 [G][S][F] expurgated erased VNs
        (get_deg_two_ss, BP_FULL.c:1227) of the absolute positions e0 .. e0+S-1, e0 = q0 - (dv-1) (first piece: e0 = 0 and
        S-(dv-1) positions), which is as far as the scan is final."""
        S, T, ms, M = self.S, self.T, self.ms, self.M
        first = self.k == 0
        if not first:
            self.A = self.D + 1 - T
        self._load()
        if first:
            engine.decode_bp_window_range(self.fb, self.W, engine.UNLIMITED, 0, S + ms, self.x, self.xb, resume=False)
            j0 = 0
        else:
            off = self._off * M
            self.x[:, : self.C * M] = self._px[:, off: off + self.C * M]
            self.xb[:, : self.C * M] = self._pxb[:, off: off + self.C * M]
            self.x[:, self.C * M:] = self.fb.chan[:, self.C * M:]
            self.xb[:, self.C * M:] = self.fb.chan[:, self.C * M:]
            engine.decode_bp_window_range(self.fb, self.W, engine.UNLIMITED, T + ms, S, self.x, self.xb, resume=True)
            j0 = T
        cnt, pairs = engine.position_counts(self.fb, F_TERMINATED | F_EXP_ALL)
        q0 = self.D + 1
        plain = cnt[:, j0: j0 + S]
        exl = cnt - 2 * pairs
        if first:
            e0, ex = 0, exl[:, : S - ms]
        else:
            e0, ex = q0 - ms, exl[:, j0 - ms: j0 - ms + S]
        self.D = self.A + j0 + S - 1
        # the next piece starts T positions before the first undecided one: local offset of its position 0 in this piece
        self._off = (self.D + 1 - T) - self.A
        self._px, self._pxb = self.x.clone(), self.xb.clone()
        self.k += 1
        return q0, plain, e0, ex


def simulate_stream(eps, dv, dc, M, W, doped_positions=(), segment=2000, max_blocks_err=1000, max_blocks=1000000, seed=0x5C1D9C,
                    frames_per_graph=128, graphs_per_batch=2, _provider=None):
    """One epsilon point of main_streaming (BP_FULL.c:1976-2050).  ``M`` = VNs per position.  Returns the thirteen
    numbers of a ``results_circular`` row (BP_FULL.c:547-560) as a dict.

    world x graphs_per_batch x frames_per_graph chains are decoded side by side, each CONTINUOUSLY (``ChainPieces``: only a
    chain's very first positions see a terminated head, like the reference's single chain does); every round advances all of
    them by ``segment`` positions, and the counters are accumulated step by step in main_streaming's order, chain after
    chain within a round, so the stop rule (BP_FULL.c:2033) cuts at a definite step whatever the number of GPUs."""
    tot = dict(num_erasures=0, num_bits_generated=0, num_blocks_err=0, num_blocks_generated=0, num_erasures_exp=0,
               num_bits_generated_exp=0, num_blocks_err_exp=0, num_blocks_generated_exp=0)
    from . import dist as D
    rank, world = D.world()
    ms = dv - 1
    chains = ChainPieces(dv, dc, M, W, eps, doped_positions, graphs_per_batch, frames_per_graph, seed,
                         first_graph_id=rank * graphs_per_batch, piece=segment, provider=_provider)
    ex_last = None                                        # expurgated count of the last position of the previous round's range
    while True:
        q0, plain, e0, ex = chains.next_piece()
        plain, ex = np.asarray(plain), np.asarray(ex)
        if world > 1:
            plain = D.allgather_rows(plain).reshape((-1,) + plain.shape[1:])
            ex = D.allgather_rows(ex).reshape((-1,) + ex.shape[1:])
        first = q0 == 0
        # decode steps of this round: step pos decides position pos-(dv-1) and expurgates position pos-2dv+1
        p0, p1 = (0, segment + ms) if first else (q0 + ms, q0 + ms + segment)
        pos = np.arange(p0, p1)
        qd, qe = pos - ms, pos - 2 * dv + 1
        nd = np.array([not is_position_doped_streaming(int(p), doped_positions) for p in range(max(0, p0 - 2 * dv), p1)], bool)
        ndo = max(0, p0 - 2 * dv)
        gen = np.where(qd >= 0, nd[np.clip(qd - ndo, 0, None)], False)
        gen2 = np.where(qe >= 0, nd[np.clip(qe - ndo, 0, None)], False)
        if ex_last is None:
            ex_last = np.zeros((plain.shape[0], plain.shape[2]), ex.dtype)
        new_last = ex[:, -1, :].copy()
        for g in range(plain.shape[0]):
            for f in range(plain.shape[2]):
                dec = np.where(qd >= 0, plain[g, np.clip(qd - q0, 0, None), f], 0)
                # expurgated counts: this round's scan starts at e0; the one position before it comes from the last round
                exq = np.concatenate([[ex_last[g, f]], ex[g, :, f]])
                exv = np.where(qe >= 0, np.maximum(exq[np.clip(qe - (e0 - 1), 0, len(exq) - 1)], 0), 0)
                c = dict(num_erasures=np.cumsum(dec), num_blocks_err=np.cumsum(dec > 0), num_erasures_exp=np.cumsum(exv),
                         num_blocks_err_exp=np.cumsum(exv > 0), num_bits_generated=np.cumsum(gen) * M, num_blocks_generated=np.cumsum(gen),
                         num_bits_generated_exp=np.cumsum(gen2) * M, num_blocks_generated_exp=np.cumsum(gen2))
                # main_streaming tests its stop rule after every decode step
                lim_e = max_blocks_err - tot["num_blocks_err_exp"]
                lim_b = max_blocks - tot["num_blocks_generated_exp"]
                hit = np.flatnonzero((c["num_blocks_err_exp"] >= lim_e) | (c["num_blocks_generated_exp"] >= lim_b))
                k = int(hit[0]) if len(hit) else len(pos) - 1
                for key in tot:
                    tot[key] += int(c[key][k])
                if len(hit):
                    return tot
        ex_last = new_last


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
