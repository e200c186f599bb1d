"""Drop-in for the reference's ``simulators_sc_ldpc/peeling_decoding/sc_ldpc.py`` (SC.py): the semi-structured
(l, r, L, M) SC-LDPC ensemble.  ``gen_slots`` returns the same object as the reference -- ``transmissions``, an int64
array [L*M, l] holding the CN index of every VN edge (SC.py:48-56) -- but the L+l-1 socket permutations are drawn on
the GPU (Philox keys + bitonic sort, ``csrc/graph_kernels.cu``) instead of with ``np.random.permutation``.
"""
from __future__ import annotations

import numpy as np

from . import engine

_state = {"seed": 0x5C1D9C, "next_graph": 0}


def set_seed(seed: int):
    """Seed of the Philox stream the ensembles are drawn from (the reference uses NumPy's global state)."""
    _state["seed"] = int(seed)
    _state["next_graph"] = 0


def _gen(l, r, L, M, n_graphs, tail_biting):
    ens = engine.Ensemble(l, r, L, M)
    fb = engine.FrameBatch(ens, n_graphs, 0, 2)
    # only vn_cn is needed here; tables are built on demand by the decoders
    import ctypes

    from . import _lib
    lib = _lib.lib()
    nbytes = lib.scldpc_graph_generate_scratch_bytes(ctypes.byref(fb.dims), int(tail_biting))
    import torch
    keys = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=fb.device)
    _lib.check(lib.scldpc_graph_generate(ctypes.byref(fb.dims), ctypes.c_void_p(fb.vn_cn.data_ptr()),
                                         ctypes.c_void_p(keys.data_ptr()), ctypes.c_uint64(_state["seed"]),
                                         ctypes.c_uint64(_state["next_graph"]), int(tail_biting), engine._stream()))
    _state["next_graph"] += n_graphs
    return fb.vn_cn


def gen_slots(l, r, L, M):
    """``transmissions`` of one code of the ensemble (SC.py:53-56)."""
    return _gen(l, r, L, M, 1, False)[0].cpu().numpy().astype(np.int64)


def gen_slots_tail_biting(l, r, L, M):
    """Tail-biting variant (SC.py:59-62): CN position (i + d) mod L."""
    return _gen(l, r, L, M, 1, True)[0].cpu().numpy().astype(np.int64)


def gen_vn_indices(l, r, L, M):
    """[L][l][M] view of ``gen_slots`` (SC.py:33-38)."""
    return np.ascontiguousarray(gen_slots(l, r, L, M).reshape(L, M, l).transpose(0, 2, 1))


def gen_vn_indices_tail_biting(l, r, L, M):
    """[L][l][M] view of ``gen_slots_tail_biting`` (SC.py:41-45)."""
    return np.ascontiguousarray(gen_slots_tail_biting(l, r, L, M).reshape(L, M, l).transpose(0, 2, 1))


def vn_indices_to_transmissions(vn_indices, l, L, M):
    """SC.py:48-50: [L][l][M] -> [L*M][l]."""
    return np.ascontiguousarray(np.asarray(vn_indices).transpose(0, 2, 1).reshape(L * M, l))
