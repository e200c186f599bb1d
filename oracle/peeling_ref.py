"""Imports the reference's own peeling_decoding.py (``/root/reference``, this container only) with its inputs injected,
to pin the peeling restatements of ``scldpc_oracle.c`` and to generate the fixtures under ``tests/golden``.

TEST INFRASTRUCTURE ONLY.  Injection points (all looked up as module globals at call time by the reference):
  * ``gen_users_sc_ldpc`` / ``gen_users_sc_ldpc_doping`` (PD.py:147,166)  -> yields Users for a given code
    (``transmissions``) and erasure mask instead of drawing them with NumPy;
  * ``pick_random_deg_1_cn`` (PD.py:1022) -> picks ``np.flatnonzero(r == 1)[u % k]`` with u from an injected sequence
    of 32-bit draws instead of ``random.choice``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("SCLDPC_REFERENCE_ROOT", "/root/reference")
PD_DIR = os.path.join(REF_ROOT, "simulators_sc_ldpc", "peeling_decoding")
_pd = None


def available() -> bool:
    return os.path.isfile(os.path.join(PD_DIR, "peeling_decoding.py"))


def load():
    """import peeling_decoding with a stub matplotlib (est_scaling_params imports it at module level, EST.py:6)."""
    global _pd
    if _pd is None:
        if "matplotlib" not in sys.modules:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
        if PD_DIR not in sys.path:
            sys.path.insert(0, PD_DIR)
        import peeling_decoding as pd  # noqa
        pd.trange = lambda n: _Quiet(range(n))
        _pd = pd
    return _pd


class _Quiet:
    def __init__(self, it):
        self.it = it

    def __iter__(self):
        return iter(self.it)

    def set_description(self, *a, **k):
        pass


def _users(pd, transmissions, erased, cns_per_pos):
    uid = 0
    for tr in transmissions[np.asarray(erased, bool)]:
        pos = int(tr[0] / cns_per_pos) * cns_per_pos
        yield pd.User(uid, pos, tr, set(), 1)
        uid += 1


def ref_peel_trajectories(e, l, r, L, M, is_terminated, frames, doping_points=[]):
    """Runs the reference ``simulate_peeling_decoder_ldpc`` on injected frames.
    frames: list of (transmissions int64 [L*M][l], erased bool [L*M], picks uint32 [>= num_pd_steps]).
    Returns (r1 [len(frames)][steps+1], plrs)."""
    pd = load()
    cns_per_pos = int(l / r * M)
    it = iter(frames)
    cur = {}

    def gen(*args):
        tr, er, picks = next(it)
        cur["picks"] = picks
        cur["step"] = 0
        return _users(pd, tr, er, cns_per_pos)

    def pick(a):
        idx = np.flatnonzero(a == 1)
        u = int(cur["picks"][cur["step"]])
        cur["step"] += 1
        if len(idx) == 0:
            return None
        return idx[u % len(idx)]

    old = (pd.gen_users_sc_ldpc, pd.gen_users_sc_ldpc_doping, pd.pick_random_deg_1_cn)
    pd.gen_users_sc_ldpc = gen
    pd.gen_users_sc_ldpc_doping = gen
    pd.pick_random_deg_1_cn = pick
    try:
        _, r1, plrs = pd.simulate_peeling_decoder_ldpc(e, l, r, L, M, is_terminated, False, len(frames), doping_points)
    finally:
        pd.gen_users_sc_ldpc, pd.gen_users_sc_ldpc_doping, pd.pick_random_deg_1_cn = old
    return r1, plrs


def ref_simulate_sc_ldpc(e, l, r, L, M, is_terminated, is_bounded, frames, max_fuckups=10 ** 9, doping_points=[], is_tail_biting=False):
    """Runs the reference ``simulate_sc_ldpc`` on injected frames (list of (transmissions, erased)); note that the
    reference enlarges L by the ignored head/tail itself (PD.py:604-607), so the injected codes must have the enlarged
    length.  Returns the 13-tuple."""
    pd = load()
    cns_per_pos = int(l / r * M)
    it = iter(frames)

    def gen(*args):
        tr, er = next(it)
        return _users(pd, tr, er, cns_per_pos)

    old = (pd.gen_users_sc_ldpc, pd.gen_users_sc_ldpc_doping)
    pd.gen_users_sc_ldpc = gen
    pd.gen_users_sc_ldpc_doping = gen
    try:
        out = pd.simulate_sc_ldpc(e, l, r, L, M, is_terminated, False, is_bounded, is_tail_biting, len(frames), max_fuckups, doping_points)
    finally:
        pd.gen_users_sc_ldpc, pd.gen_users_sc_ldpc_doping = old
    return out
