"""torchrun worker: row-partitioned multi-GPU solve vs the oracle (launched by tests/test_gpu_multi.py or by hand:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from adaptiveviscositysolver_b200.dist_plan import halo_plan
    from adaptiveviscositysolver_b200.scenes import sphere_drop
    from adaptiveviscositysolver_b200.solver import Params, Solver, nccl_unique_id
    from tests.util import perm_gpu_to_oracle

    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s = Solver(device=local, rank=rank, nranks=world, nccl_unique_id=uid[0])
    ok = True
    # A solve that converges before its first iteration (zero rhs: fluid at rest) must leave the in-kernel exchange's sequence
    # counters where the peers' flags are -- the solves below run on the same context right after it.
    rest = sphere_drop(32, 11)
    for v in rest.vel:
        v.data[...] = 0
    info0 = s.solve(rest, Params(octree_levels=4, tolerance=1e-8))
    assert info0.iterations == 0 and info0.error == 0, (info0.iterations, info0.error)
    for n, R, L in ((32, 11, 4), (64, 26, 6)):
        sc = sphere_drop(n, R, noise=0.01)
        tol = 1e-10
        out = [v.data.copy() for v in sc.vel]
        # chunked relaunches of the persistent CG kernel (check_every iterations per cooperative launch) must not change a bit
        s.solve(sc, Params(octree_levels=L, tolerance=tol, check_every=5))
        x_chunked = s.solution()
        info = s.solve(sc, Params(octree_levels=L, tolerance=tol), out)
        assert np.array_equal(x_chunked, s.solution()), "chunked and single-launch CG differ"
        rb, re = s.local_range()
        starts = s.row_starts(world)
        assert (rb, re) == (starts[rank], starts[rank + 1]) and starts[0] == 0 and starts[-1] == info.octree_dofs
        assert all(b >= a for a, b in zip(starts, starts[1:]))
        x_local = s.solution()
        parts = [None] * world
        dist.all_gather_object(parts, x_local)
        outs = [None] * world
        dist.all_gather_object(outs, [o.copy() for o in out])
        # device halo plan == host statement of the plan
        ptr, col, val, rhs, x0 = s.system()
        halo, _, cnt = halo_plan(ptr, col, rb, re, info.octree_dofs, world, starts)
        assert halo.size == info.halo_columns, (halo.size, info.halo_columns)
        halos = [None] * world
        dist.all_gather_object(halos, int(halo.size))
        if rank == 0:
            from oracle import avs_oracle as orc
            ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=L, tolerance=tol))
            x = np.concatenate(parts)
            perm = perm_gpu_to_oracle(s.keys(), ref.face_keys())
            err = np.abs(x - ref.solution()[perm]).max()
            same_out = all(all(np.array_equal(a, b) for a, b in zip(outs[0], o)) for o in outs[1:])
            frac = max(halos) / (info.octree_dofs / world)
            print(f"[dist_worker] world={world} n={n} N={info.octree_dofs} iters={info.iterations} (oracle {ref.iterations}) "
                  f"max|x-x_oracle|={err:.3e} halo/rows={frac:.3f} outputs_identical={same_out} dist_mode={info.dist_mode}", flush=True)
            ok = ok and err < 1e-6 and abs(info.iterations - ref.iterations) <= 2 and same_out and frac < 0.5
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
