"""Host-side statement of the multi-GPU row partition and halo plan.

The CUDA library builds the same plan on the device (``csrc/avs_dist.cu``); this numpy version is the
specification the CPU (gloo, world_size 2) tests run and the GPU tests compare against.

  * rank r owns a contiguous block [starts[r], starts[r+1]) of the depth-first brick row order; the device cuts
    the blocks at brick boundaries by estimated cost (``Solver.row_starts``); the CPU tests use equal counts;
  * its halo = the sorted set of off-rank columns referenced by its rows; halo slot k holds global column
    halo[k]; because the set is sorted it is grouped by owner, so one contiguous receive per neighbour;
  * a local row's column c maps to  c - row_begin  if owned, else  n_local + slot(c).
"""
from __future__ import annotations

import numpy as np


def row_range(n: int, rank: int, nranks: int):
    """Equal-count blocks (single-level / CPU tests)."""
    return n * rank // nranks, n * (rank + 1) // nranks


def halo_plan(ptr, col, row_begin: int, row_end: int, n: int, nranks: int, starts=None):
    """ptr/col: CSR of the rows [row_begin, row_end) with GLOBAL column ids.

    Returns (halo, local_col, recv_counts): halo = sorted off-rank global columns, local_col = remapped
    column array, recv_counts[q] = how many halo entries rank q owns."""
    col = np.asarray(col, np.int64)
    off = (col < row_begin) | (col >= row_end)
    halo = np.unique(col[off])
    local = col - row_begin
    local[off] = (row_end - row_begin) + np.searchsorted(halo, col[off])
    bounds = np.array(starts if starts is not None else [n * q // nranks for q in range(nranks + 1)])
    recv_counts = np.diff(np.searchsorted(halo, bounds))
    return halo, local.astype(np.int32), recv_counts
