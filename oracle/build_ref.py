#!/usr/bin/env python3
"""Build recipe for ``oracle/_ref`` -- the UNMODIFIED reference C simulators as shared libraries.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``fl_scaling_sc_ldpc_b200``) may import this.

The three reference translation units

    simulators_sc_ldpc/bp_decoding/SC_LDPC_Simulator_BPDecoder_BEC_full_BP_LimIter_OlmosRandomEnsemble.c      ("full")
    simulators_sc_ldpc/bp_decoding/SC_LDPC_Simulator_BPDecoder_BEC_SlidingWindow_LimIter_OlmosRandomEnsemble.c ("sw")
    simulators_sc_ldpc/bp_decoding/trajectories_SC_LDPC_Simulator_BPDecoder_BEC_full_BP_OlmosRandomEnsemble.c  ("traj")

fix the ensemble through ``#define Def_dv/Def_dc/Def_L/Def_M`` (lines 22-25 of each file), so one shared
object is built per (variant, dv, dc, L, Def_M).  The sources are compiled *where they lie* under
``/root/reference``: the four size lines are rewritten by a ``sed`` pipe that feeds ``gcc -x c -`` on stdin, so
no copy of the reference source is ever written to disk, let alone into the repository.  Only the ``.so``
files land in ``oracle/_ref/`` (git-ignored, but they travel to the GPU box with the snapshot).

``-Dmain=ref_main`` keeps the reference ``main`` callable but out of the way; every file-scope array
(``VNdegree``, ``CNdegree``, ``LLRsChannel``, ``VNerased``, ...) and every function (``generate_code``,
``channel_doped``, ``decodeBP``, ``decodeBP_SW``, ...) is a plain exported symbol of the resulting library and
is driven through ctypes by ``oracle/ref_driver.py``.  ``-Wl,-Bsymbolic`` binds the library's references to its
own globals (names such as ``l``, ``f``, ``dv`` are too generic to leave to the dynamic linker).

Note the reference's ``Def_M`` is the number of CNs per position (= M/2 for the (4,8) ensemble, where M is the
number of VNs per position used everywhere else in this repository).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("SCLDPC_REFERENCE_ROOT", "/root/reference")
REF_DIR = os.path.join(REF_ROOT, "simulators_sc_ldpc", "bp_decoding")

# "circ" is the "full" file with its `#undef CIRCULAR` line (BP_FULL.c:34) removed, i.e. the streaming / circular-buffer
# build the reference ships compiled out: Def_nk = L*CNsPos, main -> main_streaming.
SOURCES = {
    "circ": "SC_LDPC_Simulator_BPDecoder_BEC_full_BP_LimIter_OlmosRandomEnsemble.c",
    "full": "SC_LDPC_Simulator_BPDecoder_BEC_full_BP_LimIter_OlmosRandomEnsemble.c",
    "sw": "SC_LDPC_Simulator_BPDecoder_BEC_SlidingWindow_LimIter_OlmosRandomEnsemble.c",
    "traj": "trajectories_SC_LDPC_Simulator_BPDecoder_BEC_full_BP_OlmosRandomEnsemble.c",
}

# (dv, dc, L, Def_M): Def_M = CNs per position.  M (VNs/position) = Def_M * dc / dv.
#   tiny cases for parity tests, the reference's shipped size (L=50, Def_M=500), and the BASELINE size
#   (L=50, Def_M=5000; L=100, Def_M=5000 for the window decoder).
DEFAULT_SIZES = [
    (4, 8, 6, 8),
    (4, 8, 10, 25),
    (4, 8, 12, 32),
    (4, 8, 20, 64),
    (3, 6, 10, 24),
    (5, 10, 12, 20),
    (4, 8, 50, 500),
    (4, 8, 50, 5000),
    (4, 8, 100, 5000),
]
CIRC_SIZES = [(4, 8, 16, 8), (4, 8, 24, 16), (3, 6, 16, 12)]      # ring length L, Def_M for the streaming decoder


def so_name(variant: str, dv: int, dc: int, L: int, defM: int) -> str:
    return os.path.join(OUT, f"ref_{variant}_{dv}_{dc}_L{L}_M{defM}.so")


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, s)) for s in SOURCES.values())


def build_one(variant: str, dv: int, dc: int, L: int, defM: int, force: bool = False) -> str:
    """Compile one reference translation unit at one size.  Returns the path of the shared object."""
    out = so_name(variant, dv, dc, L, defM)
    src = os.path.join(REF_DIR, SOURCES[variant])
    if not force and os.path.isfile(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    if not os.path.isfile(src):
        raise FileNotFoundError(f"reference source not present: {src}")
    os.makedirs(OUT, exist_ok=True)
    sed = [
        "sed", "-E",
        "-e", rf"s/^#define Def_dv\s+[0-9]+/#define Def_dv {dv}/",
        "-e", rf"s/^#define Def_dc\s+[0-9]+/#define Def_dc {dc}/",
        "-e", rf"s/^#define Def_L\s+[0-9]+/#define Def_L {L}/",
        "-e", rf"s/^#define Def_M\s+[0-9]+/#define Def_M {defM}/",
    ]
    if variant == "circ":
        sed += ["-e", r"/^#undef CIRCULAR/d"]
    sed += [src]
    gcc = [
        os.environ.get("CC", "gcc"), "-O2", "-std=gnu11", "-w", "-shared", "-fPIC", "-Dmain=ref_main",
        "-Wl,-Bsymbolic", "-x", "c", "-", "-o", out + ".tmp", "-lm",
    ]
    p1 = subprocess.Popen(sed, stdout=subprocess.PIPE)
    p2 = subprocess.run(gcc, stdin=p1.stdout, capture_output=True, text=True)
    p1.stdout.close()
    if p1.wait() != 0 or p2.returncode != 0:
        raise RuntimeError(f"building {out} failed:\n{p2.stderr}")
    shutil.move(out + ".tmp", out)
    return out


def build_all(sizes=None, variants=("full", "sw", "traj"), verbose: bool = True) -> list[str]:
    if not reference_available():
        if verbose:
            print(f"[oracle/_ref] reference sources not found under {REF_DIR}; using prebuilt files only")
        return []
    outs = []
    for (dv, dc, L, defM) in (sizes or DEFAULT_SIZES):
        for v in variants:
            outs.append(build_one(v, dv, dc, L, defM))
            if verbose:
                print(f"[oracle/_ref] {os.path.relpath(outs[-1], HERE)}")
    if sizes is None:
        for (dv, dc, L, defM) in CIRC_SIZES:
            outs.append(build_one("circ", dv, dc, L, defM))
            if verbose:
                print(f"[oracle/_ref] {os.path.relpath(outs[-1], HERE)}")
    return outs


if __name__ == "__main__":
    build_all()
    sys.exit(0)
