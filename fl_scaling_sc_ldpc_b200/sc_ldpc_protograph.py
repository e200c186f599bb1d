"""Drop-in for ``simulators_sc_ldpc/peeling_decoding/sc_ldpc_protograph.py``: the protograph-based ensemble.
``gen_slots_from_position(l, r, M)`` returns, like the reference (sc_ldpc_protograph.py:17-20), an int array [M, l] whose
column i holds ``i * num_cns + perm`` with an independent uniform permutation of the ``num_cns = l*M/r`` CNs per portion
and edge type -- drawn on the GPU (``scldpc_graph_generate`` with the protograph selector)."""
from __future__ import annotations

import numpy as np

from . import engine, sc_ldpc


def gen_slots_from_position(l, r, M):
    ens = engine.Ensemble(l, r, 1, M)
    fb = engine.FrameBatch(ens, 1, 0, 2)
    fb.generate_graphs(sc_ldpc._state["seed"], first_graph_id=sc_ldpc._state["next_graph"], protograph=True)
    sc_ldpc._state["next_graph"] += 1
    return fb.vn_cn[0].cpu().numpy().astype(np.int64)
