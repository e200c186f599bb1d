"""The drop-in drivers under a multi-process launch give the files / tuples of a single-process run.

Runs `torch.distributed.run` with two ranks.  With two or more GPUs the ranks take one GPU each over NCCL; on a one-GPU box
they share cuda:0 and talk over gloo (SCLDPC_DIST_BACKEND) -- the decoders, the record gather and the sequential replay are
the same code either way."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(nproc, args, cwd):
    import torch
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if nproc > 1 and torch.cuda.device_count() < nproc:
        env["SCLDPC_DIST_BACKEND"] = "gloo"
    if nproc == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "multi_rank_driver.py")] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_rank_driver.py")] + args
    r = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_two_ranks_reproduce_single_process(tmp_path):
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir(); two.mkdir()
    _launch(1, [str(one)], str(one))
    _launch(2, [str(two)], str(two))
    names = sorted(os.listdir(one))
    assert names == sorted(os.listdir(two)) and len(names) >= 5
    for n in names:
        a, b = open(one / n, "rb").read(), open(two / n, "rb").read()
        if n.endswith(".pkl"):
            a, b = pickle.loads(a), pickle.loads(b)
            for x, y in zip(a, b):
                assert np.array_equal(np.asarray(x), np.asarray(y)), n
        else:
            assert a == b, n


def test_c_abi_allreduce_on_a_callers_nccl_communicator():
    """scldpc_allreduce_counters with a raw ncclComm_t (a C caller's sharding path): a one-rank communicator created through
    libnccl's own C API; the int64 vector comes back unchanged (sum over one rank) and the call is stream-ordered"""
    import ctypes
    import torch
    from fl_scaling_sc_ldpc_b200 import _lib
    try:
        nccl = ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)
    except OSError:
        pytest.skip("libnccl.so.2 not loadable")
    uid = (ctypes.c_char * 128)()
    assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    comm = ctypes.c_void_p()

    class Uid(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]
    u = Uid.from_buffer_copy(bytes(uid))
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    assert nccl.ncclCommInitRank(ctypes.byref(comm), 1, u, 0) == 0
    try:
        t = torch.arange(-3, 13, dtype=torch.int64, device="cuda") * (1 << 40)
        ref = t.clone()
        L = _lib.lib()
        _lib.check(L.scldpc_allreduce_counters(comm, ctypes.c_void_p(t.data_ptr()), t.numel(),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert bool((t == ref).all())
        assert L.scldpc_allreduce_counters(None, ctypes.c_void_p(t.data_ptr()), 4, None) != 0
    finally:
        nccl.ncclCommDestroy(comm)
