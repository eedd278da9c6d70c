"""History-coloured count matrix of weighted-ensemble lineages on the GPU.

reference: ``NonMarkovModel.fit`` (msm_we/nmm.py:117-167) applied to the discrete trajectories obtained by tracing
every walker of the last iteration back through ``seg_index['parent_id']``: ``nm_cmatrix[2 s_prev + col_prev,
2 s_now + col_now] += 1`` for every transition whose two history colours (A -> 0, B -> 1, inherited otherwise) are
defined.  The reference walks (leaves x depth) states in Python; lineages share their early segments, so here every
SEGMENT is visited once -- forward for its colour, backward for the number of traced trajectories through it -- and
the coloured transition records go through K3 (``flux_accumulate`` with ``C = 2``), which sums them in a fixed order.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import lib, check
from .engine import require_cuda


def _stream():
    return torch.cuda.current_stream().cuda_stream


def coloured_lineage_counts(labels, parents, n_states, state_a, state_b, device=None):
    """``labels[it]`` : int64 [S_it] discrete state of every segment of iteration ``it`` (0-based list over iterations);
    ``parents[it]``   : int64 [S_it] index of the segment's parent in iteration ``it - 1`` (< 0: none; ignored for it = 0);
    returns the ``[2 n_states, 2 n_states]`` float64 count matrix (numpy) over the lineages of the LAST iteration's
    segments, with the reference's lag-1 sliding-window convention (the first frame of a trajectory is never coloured)."""
    dev = require_cuda(device)
    n_it = len(labels)
    if n_it < 2:
        return np.zeros((2 * n_states, 2 * n_states))
    cls = np.zeros(n_states, dtype=np.uint8)
    cls[np.asarray(list(state_b), dtype=np.int64)] = 2
    cls[np.asarray(list(state_a), dtype=np.int64)] = 1          # A is tested first (nmm.py:140-145)
    cls_d = torch.from_numpy(cls).to(dev)
    lab = [torch.from_numpy(np.ascontiguousarray(l, dtype=np.int64)).to(dev) for l in labels]
    par = [torch.from_numpy(np.ascontiguousarray(p, dtype=np.int64)).to(dev) for p in parents]
    S = [int(l.numel()) for l in lab]
    # forward: colours (iteration 0 stays undefined: the reference's walk starts at index `lag`)
    colour = [torch.full((S[0],), -1, dtype=torch.int8, device=dev)]
    for it in range(1, n_it):
        c = torch.empty(S[it], dtype=torch.int8, device=dev)
        prev = colour[it - 1] if it >= 2 else None
        check(lib.mwe_lineage_colour(lab[it].data_ptr(), par[it].data_ptr(), S[it], None if prev is None else prev.data_ptr(),
                                     S[it - 1], cls_d.data_ptr(), n_states, c.data_ptr(), _stream()), "mwe_lineage_colour")
        colour.append(c)
    # backward: traced trajectories through every segment
    leaves = [None] * n_it
    leaves[-1] = torch.ones(S[-1], dtype=torch.int64, device=dev)
    for it in range(n_it - 1, 0, -1):
        leaves[it - 1] = torch.zeros(S[it - 1], dtype=torch.int64, device=dev)
        check(lib.mwe_lineage_leaves(par[it].data_ptr(), leaves[it].data_ptr(), S[it], leaves[it - 1].data_ptr(), S[it - 1],
                                     _stream()), "mwe_lineage_leaves")
    # records of iterations 1 .. n_it-1, stacked, then ONE K3 launch sequence
    total = int(sum(S[1:]))
    start = torch.empty(total, dtype=torch.int64, device=dev)
    end = torch.empty(total, dtype=torch.int64, device=dev)
    col0 = torch.empty(total, dtype=torch.uint8, device=dev)
    col1 = torch.empty(total, dtype=torch.uint8, device=dev)
    w = torch.empty(total, dtype=torch.float64, device=dev)
    pos = 0
    offs = [0]
    for it in range(1, n_it):
        s = S[it]
        if s:
            check(lib.mwe_lineage_records(lab[it - 1].data_ptr(), colour[it - 1].data_ptr(), S[it - 1], lab[it].data_ptr(),
                                          colour[it].data_ptr(), par[it].data_ptr(), leaves[it].data_ptr(), s,
                                          start[pos:].data_ptr(), end[pos:].data_ptr(), col0[pos:].data_ptr(),
                                          col1[pos:].data_ptr(), w[pos:].data_ptr(), _stream()), "mwe_lineage_records")
        pos += s
        offs.append(pos)
    errors = ops.DeviceErrors(dev)
    dense = ops.flux_accumulate(start, end, w, int(n_states) - 2, col0=col0, col1=col1, C=2,
                                iter_offsets=torch.tensor(offs, dtype=torch.int64, device=dev), errors=errors)
    out = dense.cpu().numpy()
    errors.check()
    return out
