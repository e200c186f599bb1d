"""The node-state formulation the GPU decoders use by default, restated in numpy and held against the CPU oracle
(decodeBP / decodeBP_SW restatements, themselves pinned to the compiled reference): erased set, iteration counts and
per-window iteration counts must be identical for every schedule the library runs.

  full BP :  x_v(t) = x_v(t-1) AND NOT (some CN of v has v as its only erased neighbour in x(t-1))
  windows :  two planes -- x (what the CNs see; changes only when the VN is swept) and xb (what the VN has been told so
             far, also by CNs of windows that do not sweep it); CN sweep over [c0, c1) clears in xb, VN sweep over
             [v0, v1) copies xb to x.
"""
import numpy as np
import pytest

import oracle


def _cn_neighbours(g):
    nb = [[] for _ in range(g.nk)]
    for v in range(g.n):
        for c in g.vn_cn[v]:
            nb[c].append(v)
    return [np.asarray(a, np.int64) for a in nb]


def node_full(g, chan, is_term, max_it):
    clim = g.nk if is_term else g.L * g.cns_pos
    x = chan.astype(bool).copy()
    prec, it = g.n, 0
    while True:
        cnt = np.zeros(g.nk, np.int32)
        np.add.at(cnt, g.vn_cn.reshape(-1), np.repeat(x, g.dv).astype(np.int32))
        two = cnt >= 2
        two[clim:] = True                       # CNs beyond the truncation are never swept
        x = x & two[g.vn_cn].all(axis=1)
        it += 1
        ne = int(x.sum())
        if ne == 0 or ne == prec or it >= max_it:
            return it, x
        prec = ne


def node_window(g, nb, chan, W, max_it, init_it, square, is_term):
    n, L, vp, cp, ms = g.n, g.L, g.vns_pos, g.cns_pos, g.dv - 1
    clip = g.nk if is_term else L * cp
    x = chan.astype(bool).copy()
    xb = x.copy()
    erased = np.zeros(n, np.uint8)
    total = p1 = blocks = 0
    win_iters = []
    for pos in range(L if square else L + ms):
        sc, ec = pos * cp, min(pos * cp + W * cp, clip)
        if square:
            sv, ev = pos * vp, pos * vp + W * vp
        elif pos <= ms:
            sv, ev = 0, (W + pos) * vp
        else:
            sv, ev = (pos - ms) * vp, (pos - ms) * vp + (W + ms) * vp
        ev = min(ev, n)
        it = done = here = 0
        prec = n
        cap = init_it if (square and pos == 0) else max_it
        while True:
            for c in range(sc, ec):
                er = nb[c][x[nb[c]]] if len(nb[c]) else nb[c]
                if len(er) == 1:
                    xb[er[0]] = False
            x[sv:ev] = xb[sv:ev]
            if square or pos >= ms:
                erased[sv:sv + vp] = xb[sv:sv + vp]
                here = int(xb[sv:sv + vp].sum())
            term = int(xb[sv:ev].sum())
            done += 1
            if term == 0 or term == prec:
                break
            prec = term
            it += 1
            if it >= cap:
                break
        win_iters.append(done)
        total += here
        blocks += here > 0
        if ms <= pos <= W - 2:
            p1 += here
    return total, p1, blocks, erased, np.asarray(win_iters)


@pytest.mark.parametrize("dv,dc,L,M", [(4, 8, 12, 32), (3, 6, 10, 24), (5, 10, 8, 20)])
def test_full_bp_node_state_equals_message_passing(dv, dc, L, M):
    rng = np.random.default_rng(L * M)
    oracle.srandom(L * M)
    perm = None
    for _ in range(3):
        g, perm = oracle.generate_code(L, M, M * dv // dc, dv, dc, perm)
        for eps in (0.8 * dv / dc, 0.92 * dv / dc, 1.0 * dv / dc, 1.1 * dv / dc):
            for is_term in (1, 0):
                for cap in (10 ** 9, 7, 1):
                    for rep in range(3):
                        chan = (rng.random(g.n) < eps).astype(np.int32)
                        if rep == 2:
                            chan[2 * M:3 * M] = 0                       # a doped position
                        o = oracle.decode_bp(g, chan, cap, is_term)
                        it, x = node_full(g, chan, is_term, cap)
                        assert it == o["iters"] and (x.astype(np.uint8) == o["erased"]).all(), (eps, is_term, cap, rep)


@pytest.mark.parametrize("dv,dc,L,M", [(4, 8, 10, 16), (3, 6, 9, 12)])
def test_window_node_state_equals_message_passing(dv, dc, L, M):
    rng = np.random.default_rng(7 * L + M)
    oracle.srandom(7 * L + M)
    perm = None
    for _ in range(2):
        g, perm = oracle.generate_code(L, M, M * dv // dc, dv, dc, perm)
        nb = _cn_neighbours(g)
        for eps in (0.8 * dv / dc, 0.95 * dv / dc, 1.05 * dv / dc):
            for (W, cap, init) in ((3, 4, 10), (5, 6, 60), (4, 1, 1), (L + 3, 3, 3)):
                for square in (1, 0):
                    for is_term in (1, 0):
                        chan = (rng.random(g.n) < eps).astype(np.int32)
                        o = oracle.decode_bp_sw(g, chan, W, cap, init, square, is_term)
                        r = node_window(g, nb, chan, W, cap, init, square, is_term)
                        assert r[0] == o["residual"] and r[1] == o["erasures_p1"] and r[2] == o["blocks_err"], (eps, W, cap, square, is_term)
                        assert (r[3] == o["erased"]).all() and (r[4] == o["win_iters"]).all(), (eps, W, cap, square, is_term)
