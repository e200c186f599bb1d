"""ctypes driver for the shared objects built by ``oracle/build_ref.py`` (the unmodified reference C code).

TEST INFRASTRUCTURE ONLY -- may be imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by the product package.

The reference keeps the code, the channel realisation and every message in file-scope arrays
(``BP_FULL.c:87-103``); this driver fills / reads them in place and calls the reference functions:

* ``generate_code``      (BP_FULL.c:1656)  -- glibc ``random()`` Fisher-Yates ensemble
* ``channel_doped``      (BP_FULL.c:1547)
* ``decodeBP``           (BP_FULL.c:900, BP_TRAJ.c:901 with ``FILE*`` + ``is_term``)
* ``decodeBP_SW``        (BP_FULL.c:627 classical window, BP_SW.c:628 square window)

File abbreviations as in SURVEY.md.
"""
from __future__ import annotations

import ctypes
import os
import tempfile
import threading

import numpy as np

from . import build_ref

_libc = ctypes.CDLL(None)
_libc.fopen.restype = ctypes.c_void_p
_libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
_libc.fclose.argtypes = [ctypes.c_void_p]
_libc.srandom.argtypes = [ctypes.c_uint]
_libc.random.restype = ctypes.c_long


def _run_big_stack(fn, *args, stack_mb: int = 512):
    """Run ``fn`` on a thread with a large stack (``generate_code`` keeps an (L+dv-1) x (CNsPos*dc) int VLA
    on the stack, BP_FULL.c:1664 -- 8.5 MB at M=10000, more than the default 8 MB limit)."""
    result = {}

    def target():
        try:
            result["v"] = fn(*args)
        except BaseException as e:  # pragma: no cover
            result["e"] = e

    old = threading.stack_size(stack_mb * 1024 * 1024)
    try:
        t = threading.Thread(target=target)
        t.start()
        t.join()
    finally:
        threading.stack_size(old)
    if "e" in result:
        raise result["e"]
    return result.get("v")


class RefLib:
    """One compiled reference translation unit at one (dv, dc, L, Def_M)."""

    def __init__(self, variant: str, dv: int, dc: int, L: int, defM: int):
        path = build_ref.so_name(variant, dv, dc, L, defM)
        if not os.path.isfile(path):
            if build_ref.reference_available():
                path = build_ref.build_one(variant, dv, dc, L, defM)
            else:
                raise FileNotFoundError(f"{path} not built and /root/reference not present")
        self.variant, self.dv, self.dc, self.L, self.defM = variant, dv, dc, L, defM
        self.cns_pos = defM
        # Def_VNsPos is hard-wired to Def_M*2 in the reference (BP_FULL.c:28), i.e. dc/dv == 2 is assumed.
        if dc != 2 * dv:
            raise ValueError("the reference hard-codes VNs/position = 2*Def_M; only dc == 2*dv ensembles compile correctly")
        self.vns_pos = 2 * defM
        self.n = self.vns_pos * L
        self.nk = L * defM if variant == "circ" else (L + dv - 1) * defM      # Def_nk under CIRCULAR (BP_FULL.c:36-40)
        self.lib = ctypes.CDLL(path)
        lib = self.lib
        self._VN = (ctypes.c_int * (self.n * (dv + 1))).in_dll(lib, "VNdegree")
        self._CN = (ctypes.c_int * (self.nk * (dc + 1))).in_dll(lib, "CNdegree")
        self._ch = (ctypes.c_int * self.n).in_dll(lib, "LLRsChannel")
        self._er = (ctypes.c_char * self.n).in_dll(lib, "VNerased")
        self._perm = (ctypes.c_int * (defM * dc)).in_dll(lib, "perm_code")
        self.VNdegree = np.ctypeslib.as_array(self._VN).reshape(self.n, dv + 1)
        self.CNdegree = np.ctypeslib.as_array(self._CN).reshape(self.nk, dc + 1)
        self.LLRsChannel = np.ctypeslib.as_array(self._ch)
        self.VNerased = np.frombuffer(self._er, dtype=np.uint8)
        self.perm_code = np.ctypeslib.as_array(self._perm)
        self._g = lambda name: ctypes.c_int.in_dll(lib, name)
        # dv / dc / MaxNumIt are run-time globals assigned by initialize_variables (BP_FULL.c:221-250)
        self._g("dv").value = dv
        self._g("dc").value = dc
        self.reset_perm()
        lib.generate_code.restype = ctypes.c_int
        lib.decodeBP.restype = ctypes.c_int
        lib.decodeBP_SW.restype = ctypes.c_int

    # ---- ensemble ---------------------------------------------------------------------------------
    def reset_perm(self):
        """``inizio_sim`` resets ``perm_code`` to the identity at every epsilon point (BP_FULL.c:308-311)."""
        self.perm_code[:] = np.arange(self.perm_code.size, dtype=np.int32)

    @staticmethod
    def srandom(seed: int):
        _libc.srandom(ctypes.c_uint(seed))

    def generate_code(self) -> np.ndarray:
        """Reference ``generate_code``; returns a copy of the VN->CN table ``int32[n][dv]``."""
        a = [ctypes.c_int(x) for x in (self.L, self.vns_pos, self.cns_pos, self.n, self.nk)]
        _run_big_stack(self.lib.generate_code, *a)
        return self.VNdegree[:, 1:].astype(np.int32).copy()

    def set_graph(self, vn_cn: np.ndarray):
        """Inject a graph exactly the way ``generate_code`` lays it out (BP_FULL.c:1702-1716): VN rows hold the
        degree in column 0, CN rows are appended in VN order."""
        vn_cn = np.asarray(vn_cn, dtype=np.int32).reshape(self.n, self.dv)
        self.VNdegree[:, 0] = self.dv
        self.VNdegree[:, 1:] = vn_cn
        self.CNdegree[:, 0] = 0
        CN = self.CNdegree
        for v in range(self.n):
            for i in range(self.dv):
                c = vn_cn[v, i]
                CN[c, 1 + CN[c, 0]] = v
                CN[c, 0] += 1

    def set_graph_fast(self, vn_cn: np.ndarray):
        """``set_graph`` without the Python loop (full-size graphs): CN rows list their VNs in ascending (v, i) order,
        which is the order ``generate_code`` appends them in (BP_FULL.c:1702-1716)."""
        vn_cn = np.asarray(vn_cn, dtype=np.int32).reshape(self.n, self.dv)
        self.VNdegree[:, 0] = self.dv
        self.VNdegree[:, 1:] = vn_cn
        flat = vn_cn.reshape(-1).astype(np.int64)
        vs = np.repeat(np.arange(self.n, dtype=np.int32), self.dv)
        order = np.argsort(flat, kind="stable")
        c_sorted, v_sorted = flat[order], vs[order]
        deg = np.bincount(flat, minlength=self.nk)
        if deg.size != self.nk or deg.max() > self.dc:
            raise ValueError("invalid graph")
        start = np.cumsum(deg) - deg
        rank = np.arange(flat.size) - start[c_sorted]
        self.CNdegree[:, 0] = deg
        self.CNdegree[c_sorted, 1 + rank] = v_sorted

    # ---- channel ----------------------------------------------------------------------------------
    def set_channel(self, erased: np.ndarray):
        self.LLRsChannel[:] = np.asarray(erased).astype(np.int32).reshape(self.n)

    def channel_doped(self, eps: float, doped=()):
        arr = (ctypes.c_int * max(1, len(doped)))(*doped)
        self.lib.channel_doped(ctypes.c_int(self.n), ctypes.c_double(eps), ctypes.c_int(self.vns_pos),
                               ctypes.c_int(len(doped)), arr)
        return self.LLRsChannel.astype(np.uint8).copy()

    # ---- decoders ---------------------------------------------------------------------------------
    def decode_bp(self, max_it: int, is_term: int = 1, W: int = 0, exp_positions: int | None = None) -> dict:
        """Reference ``decodeBP``.  For the "traj" variant the per-iteration rows are parsed from the text the
        reference writes (``iter deg1 dVNs first_erased_pos``).  ``exp_positions`` is the ``L`` argument, which the
        function uses only as the range of its post-decoding expurgation scan (BP_FULL.c:1075); passing 0 skips
        that scan (used when timing the iterations of capped runs, where it would be quadratic in the erasures)."""
        self._g("MaxNumIt").value = max_it
        nb, ne, nbe = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        Larg = self.L if exp_positions is None else exp_positions
        args = [ctypes.c_int(x) for x in (self.n, self.nk, Larg, W, self.vns_pos, self.cns_pos)]
        args += [ctypes.byref(nb), ctypes.byref(ne), ctypes.byref(nbe)]
        rows = None
        if self.variant == "traj":
            fd, path = tempfile.mkstemp(suffix=".traj")
            os.close(fd)
            fp = _libc.fopen(path.encode(), b"w")
            try:
                res = self.lib.decodeBP(*args, ctypes.c_void_p(fp), ctypes.c_int(is_term))
            finally:
                _libc.fclose(ctypes.c_void_p(fp))
            with open(path) as f:
                txt = f.read()
            os.unlink(path)
            rows = np.array([[int(x) for x in ln.split("\t")] for ln in txt.split("\n") if ln.strip()],
                            dtype=np.int64).reshape(-1, 4)
        else:
            if not is_term:
                raise ValueError("only BP_TRAJ.c's decodeBP has the is_term switch")
            res = self.lib.decodeBP(*args)
        return dict(residual=int(res), blocks_err=nb.value, erasures_exp=ne.value, blocks_err_exp=nbe.value,
                    erased=self.VNerased.copy(), rows=rows)

    def decode_bp_sw(self, W: int, max_it: int, init_it: int = 0) -> dict:
        """Reference ``decodeBP_SW``: classical window in the "full"/"traj" files, square window in "sw"."""
        self._g("MaxNumIt").value = max_it
        if self.variant == "sw":
            self._g("InitNumIt").value = init_it if init_it else max_it  # BP_SW.c:2099-2102
        p1, nb, ne, nbe = (ctypes.c_int(0) for _ in range(4))
        args = [ctypes.c_int(x) for x in (self.n, self.nk, self.L, W, self.vns_pos, self.cns_pos)]
        args += [ctypes.byref(p1), ctypes.byref(nb), ctypes.byref(ne), ctypes.byref(nbe)]
        res = self.lib.decodeBP_SW(*args)
        return dict(residual=int(res), erasures_p1=p1.value, blocks_err=nb.value, erasures_exp=ne.value,
                    blocks_err_exp=nbe.value, erased=self.VNerased.copy())


    # ---- streaming / circular buffer (variant "circ") -------------------------------------------------
    def stream_run(self, n_positions: int, W: int, eps: float, doped=()) -> dict:
        """Drives the reference's streaming decoder exactly like main_streaming (BP_FULL.c:1991-2046) for n_positions
        decode steps and returns, besides the per-step outputs, the UNROLLED chain it decoded: for every generated
        absolute position p the VN->CN table (absolute CN ids (p+i)*CNsPos + local) and the channel values."""
        assert self.variant == "circ"
        lib, L, V, C, dv = self.lib, self.L, self.vns_pos, self.cns_pos, self.dv
        ci = ctypes.c_int
        darr = (ctypes.c_int * max(1, len(doped)))(*doped)
        lib.initialize_arrays_circular(ci(self.n), ci(self.nk), ci(L), ci(C))
        lag = L // 2
        vn_cn, chan = [], []

        def gen(p):
            lib.generate_stream_pos(ci(p), ci(L), ctypes.c_double(eps), ci(V), ci(C), ci(len(doped)), darr)
            lib.initialize_messages_circular(ci(p), ci(L), ci(V), ci(C))
            pb = p % L
            rows = self.VNdegree[pb * V:(pb + 1) * V, 1:].astype(np.int64)
            # ring CN id -> absolute: edge i of a VN at absolute position p goes to absolute CN position p + i
            local = rows % C
            vn_cn.append(((p + np.arange(dv))[None, :] * C + local).astype(np.int32))
            chan.append(self.LLRsChannel[pb * V:(pb + 1) * V].astype(np.uint8).copy())

        for p in range(lag):
            gen(p)
        nb, ne, nbe = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        lib.decodeBP_SW_circular.restype = ctypes.c_int
        steps = []
        g = lag
        for pos in range(n_positions):
            er = lib.decodeBP_SW_circular(ci(pos), ci(self.n), ci(L), ci(W), ci(V), ci(C), ctypes.byref(nb), ctypes.byref(ne),
                                          ctypes.byref(nbe))
            steps.append((int(er), nb.value, ne.value, nbe.value))
            gen(g)
            g += 1
        return dict(vn_cn=np.stack(vn_cn), chan=np.stack(chan), steps=np.array(steps, np.int64))


_cache: dict = {}


def get(variant: str, dv: int, dc: int, L: int, defM: int) -> RefLib:
    key = (variant, dv, dc, L, defM)
    if key not in _cache:
        _cache[key] = RefLib(*key)
    return _cache[key]


def available(variant: str, dv: int, dc: int, L: int, defM: int) -> bool:
    return os.path.isfile(build_ref.so_name(variant, dv, dc, L, defM)) or build_ref.reference_available()
