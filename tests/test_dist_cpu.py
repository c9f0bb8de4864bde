"""N>1 path on CPU: world_size-2 gloo run of the row-partitioned Jacobi-PCG (host-side plan + exchange
pattern), checked against the single-process oracle solve.  The GPU library implements the same plan with
NCCL (csrc/avs_dist.cu); tests/test_gpu_multi.py checks that one on real GPUs."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptiveviscositysolver_b200.dist_plan import halo_plan, row_range
    from adaptiveviscositysolver_b200.scenes import sphere_drop
    from oracle import avs_oracle as orc
    import scipy.sparse as sp

    orc.set_num_threads(1)
    sc = sphere_drop(32, 11)
    tol = 1e-9
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=4, tolerance=tol))
    ptr, col, val = ref.csr()
    n = ref.n_face
    rb, re = row_range(n, rank, world)
    lptr = ptr[rb:re + 1] - ptr[rb]
    lcol, lval = col[ptr[rb]:ptr[re]], val[ptr[rb]:ptr[re]]
    halo, local_col, recv_counts = halo_plan(lptr, lcol, rb, re, n, world)
    nl, nh = re - rb, halo.size
    A = sp.csr_matrix((lval, local_col, lptr), shape=(nl, nl + nh))
    # send lists = what the peers' halos ask of me
    halos = [None] * world
    dist.all_gather_object(halos, halo)
    b, x = ref.rhs()[rb:re].copy(), ref.x0()[rb:re].copy()
    invd = 1.0 / A[:, :nl].diagonal()

    def exchange(v):
        """returns v extended with halo values pulled from the owners"""
        parts = [None] * world
        dist.all_gather_object(parts, v)          # CPU stand-in for the grouped send/recv
        full = np.concatenate(parts)
        return np.concatenate([v, full[halo]])

    def allsum(*vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return t.tolist()

    r = b - A @ exchange(x)
    (bb, rr) = allsum(b @ b, r @ r)
    thr = max(tol * tol * bb, np.finfo(np.float64).tiny)
    p = invd * r
    (rho,) = allsum(r @ p)
    it = 0
    if rr >= thr:
        while it < 2500:
            t = A @ exchange(p)
            (pt,) = allsum(p @ t)
            alpha = rho / pt
            x += alpha * p
            r -= alpha * t
            z = invd * r
            rr, rz = allsum(r @ r, r @ z)
            if rr < thr:
                break
            p = z + (rz / rho) * p
            rho = rz
            it += 1
    np.save(os.path.join(out_dir, f"x{rank}.npy"), x)
    if rank == 0:
        np.save(os.path.join(out_dir, "ref.npy"), ref.solution())
        np.save(os.path.join(out_dir, "meta.npy"), np.array([it, ref.iterations, recv_counts.sum(), n]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partitioned_cg_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x = np.concatenate([np.load(tmp_path / f"x{r}.npy") for r in range(world)])
    ref = np.load(tmp_path / "ref.npy")
    it, it_ref, nhalo, n = np.load(tmp_path / "meta.npy")
    assert abs(it - it_ref) <= 1
    assert np.abs(x - ref).max() < 1e-8
    assert 0 < nhalo < n   # (the oracle numbers level -> axis -> tile, so its blocks are not compact; the GPU uses Morton bricks)


def test_halo_plan_properties():
    sys.path.insert(0, str(ROOT))
    from adaptiveviscositysolver_b200.dist_plan import halo_plan, row_range
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    n = 200
    M = sp.random(n, n, density=0.05, random_state=1, format="csr") + sp.identity(n, format="csr")
    M = M.tocsr(); M.sort_indices()
    for P in (1, 2, 3, 8):
        covered = 0
        for r in range(P):
            rb, re = row_range(n, r, P)
            covered += re - rb
            lptr = M.indptr[rb:re + 1] - M.indptr[rb]
            lcol = M.indices[M.indptr[rb]:M.indptr[re]]
            halo, local, cnt = halo_plan(lptr, lcol, rb, re, n, P)
            assert cnt.sum() == halo.size and cnt[r] == 0
            assert np.all(np.diff(halo) > 0)
            # the remap is invertible
            glob = np.where(local < re - rb, local + rb, halo[np.maximum(local - (re - rb), 0)] if halo.size else 0)
            assert np.array_equal(glob, lcol)
        assert covered == n
