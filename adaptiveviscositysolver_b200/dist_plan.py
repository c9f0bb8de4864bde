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


def slab_cuts(plane_liquid_cells, nranks: int, plane_cells: int):
    """z-plane cuts of the REGULAR grid for the slab-sharded stages (regular-face classification and the stage-11
    write-back, ``avs_slab_cuts`` in csrc/avs_labels.cu): ``cuts[q] .. cuts[q+1]`` are rank q's cell planes.

    The weight of a plane is its number of liquid cells (sdf < 0) plus a floor of 2 % of a full plane (every plane
    still costs one streaming pass); cut q is the first plane boundary where the running weight reaches q/P of the
    total.  Pure integer arithmetic on identical inputs, hence identical on every rank."""
    w = np.asarray(plane_liquid_cells, dtype=np.int64) + max(1, int(plane_cells) // 50)
    nz, total = w.size, int(w.sum())
    cuts = [0] * (nranks + 1)
    cuts[nranks] = nz
    run, q = 0, 1
    for z in range(nz):
        if q >= nranks:
            break
        run += int(w[z])
        while q < nranks and run * nranks >= total * q:
            cuts[q] = z + 1
            q += 1
    for r in range(q, nranks):
        cuts[r] = nz
    return cuts


def slab_range(cuts, axis: int, rank: int, face_planes: int):
    """Planes [z0, z1) of the axis-``axis`` face grid that ``rank`` owns: the z-face grid has one more plane than
    there are cell planes, the last rank takes it (``avs_slab_range``)."""
    nranks = len(cuts) - 1
    return cuts[rank], (face_planes if rank == nranks - 1 else cuts[rank + 1])


def deal_frames(nframes: int, rank: int, nranks: int):
    """C5: the frames of a prescribed-geometry sequence are independent solves, dealt round-robin (bench.py)."""
    return [f for f in range(nframes) if f % nranks == rank]
