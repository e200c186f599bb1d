"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/scldpc.h declares, validates its
arguments, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from fl_scaling_sc_ldpc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "scldpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(scldpc_[a-z_0-9]+)\s*\(", txt)))


def test_library_is_built_and_exports_header_symbols():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/scldpc.h but not exported"
    for s in _lib.EXPORTS:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"
    assert L.scldpc_version() >= 100


def test_argument_validation_needs_no_gpu():
    L = _lib.lib()
    bad = _lib.Dims(4, 8, 10, 50, 26, 1, 2, 100)      # vns_pos*dv != cns_pos*dc
    assert L.scldpc_bp_workspace_bytes(ctypes.byref(bad), 0) == 0
    assert b"cns_pos" in L.scldpc_last_error()
    bad = _lib.Dims(4, 8, 10, 50, 25, 1, 3, 100)      # n_words not a power of two
    assert L.scldpc_bp_workspace_bytes(ctypes.byref(bad), 0) == 0
    ok = _lib.Dims(4, 8, 10, 50, 25, 2, 2, 100)
    n, nk, E = 500, 13 * 25, 2000
    need = L.scldpc_bp_workspace_bytes(ctypes.byref(ok), _lib.F_MESSAGES)
    assert need >= 2 * ((E + 1) * 16 + nk * 8 * 16)   # the two message arrays dominate
    assert L.scldpc_bp_workspace_bytes(ctypes.byref(ok), _lib.F_TRAJECTORY) > need          # + the CNresolved latch
    # the layout is a function of the flags alone: a size query and a call with the same flags agree
    assert L.scldpc_bp_workspace_bytes(ctypes.byref(ok), 0) == L.scldpc_bp_workspace_bytes(ctypes.byref(ok), 0)
    assert L.scldpc_bp_stream_workspace_bytes(ctypes.byref(ok), 0) != L.scldpc_bp_stream_workspace_bytes(ctypes.byref(ok), _lib.F_MESSAGES)


def test_no_cpu_fallback():
    L = _lib.lib()
    if L.scldpc_device_count() > 0:
        pytest.skip("a GPU is present")
    d = _lib.Dims(4, 8, 10, 50, 25, 1, 2, 100)
    rc = L.scldpc_graph_generate(ctypes.byref(d), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_uint64(1),
                                 ctypes.c_uint64(0), 0, None)
    assert rc == -2 and b"no CUDA device" in L.scldpc_last_error()
    import numpy as np
    from fl_scaling_sc_ldpc_b200 import Ensemble, ScldpcError, decode_host
    with pytest.raises(ScldpcError):
        decode_host(Ensemble(4, 8, 10, 50), np.zeros((1, 500, 4), np.int32), np.zeros((1, 2, 500), np.uint8))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fl_scaling_sc_ldpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "scldpc_oracle" not in src, f


def test_header_is_plain_c_and_a_c_client_links_and_fails_loudly_without_a_gpu(tmp_path):
    """the drop-in boundary from C: include/scldpc.h compiles as C11, examples/decode_host.c links against the library,
    and -- on a box without a GPU -- the call returns the library's error instead of computing anything on the CPU"""
    import subprocess
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "fl_scaling_sc_ldpc_b200")
    hdr_check = tmp_path / "hdr.c"
    hdr_check.write_text('#include "scldpc.h"\nint main(void) { return scldpc_version() < 0; }\n')
    exe = tmp_path / "decode_host"
    for src, out in ((str(hdr_check), str(tmp_path / "hdr")), (os.path.join(ROOT, "examples", "decode_host.c"), str(exe))):
        r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", src, "-I", inc, "-L", libdir, "-lscldpc",
                            "-Wl,-rpath," + libdir, "-o", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    if _lib.lib().scldpc_device_count() == 0:
        assert r.returncode == 2 and "no CUDA device" in r.stderr
    else:
        assert r.returncode == 0 and r.stdout.startswith("frames 16")


def test_library_was_built_from_the_sources_on_disk():
    """scldpc_build_info carries the hash of the sources the .so was compiled from; it must be the hash of csrc/ as it is now"""
    from fl_scaling_sc_ldpc_b200 import _lib
    info = _lib.lib().scldpc_build_info().decode()
    assert "arch=sm_100a" in info and ("src=" + _lib.source_hash()) in info, info
