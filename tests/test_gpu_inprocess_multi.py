"""The single-process multi-GPU entry (avs_create_multi / avs_solve_multi, include/avs.h): one host thread drives the
row-partitioned solve on several ranks -- the shape of the DOP, whose solveGasSubclass runs on one cook thread
(HDK_AdaptiveViscosity.cpp:126-128).  No NCCL, no CUDA IPC.  With one GPU visible the ranks SHARE it (devices = [0, 0]), so the
whole multi-rank path -- partition, halo lists, in-kernel halo push and scalar all-reduce of the persistent CG kernel, slab-wise
write-back -- is exercised on the driver's one-GPU box too; with two or more GPUs every rank gets its own."""
import numpy as np
import pytest

from adaptiveviscositysolver_b200.scenes import sphere_drop
from oracle import avs_oracle as orc
from tests.util import perm_gpu_to_oracle

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [r % max(have, 1) for r in range(n)]


@pytest.mark.parametrize("nranks", [2, 3])
def test_single_process_multi_rank_solve(nranks):
    from adaptiveviscositysolver_b200.solver import MultiSolver, Params, Solver
    m = MultiSolver(_devices(nranks))
    single = Solver(device=0)
    try:
        # a solve that converges before its first iteration, then real ones on the same contexts (sequence counters of the
        # in-kernel exchange must stay in step with the peers' flags)
        rest = sphere_drop(32, 11)
        for v in rest.vel:
            v.data[...] = 0
        info0 = m.solve(rest, Params(octree_levels=4, tolerance=1e-8))
        assert info0.iterations == 0
        for n, R, L in ((32, 11, 4), (64, 26, 6)):
            sc = sphere_drop(n, R, noise=0.01)
            p = Params(octree_levels=L, tolerance=1e-10)
            out = [v.data.copy() for v in sc.vel]
            info = m.solve(sc, p, out)
            ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=L, tolerance=1e-10))
            x = m.solution()
            assert x.size == ref.n_face == info.octree_dofs
            perm = perm_gpu_to_oracle(m.keys(), ref.face_keys())
            assert info.error < 1e-10 and abs(info.iterations - ref.iterations) <= 2
            assert np.abs(x - ref.solution()[perm]).max() < 1e-6
            # row blocks tile [0, N) and every rank did work
            ranges = [m.rank(r).local_range() for r in range(nranks)]
            assert ranges[0][0] == 0 and ranges[-1][1] == info.octree_dofs
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(nranks - 1)) and all(e > b for b, e in ranges)
            # the regular-grid output assembled from the ranks' z-slabs == oracle write-back == single-GPU result
            out1 = [v.data.copy() for v in sc.vel]
            info1 = single.solve(sc, p, out1)
            assert abs(info1.iterations - info.iterations) <= 1
            for a in range(3):
                assert np.abs(out[a].astype(np.float64) - ref.out_velocity(a)).max() < 1e-6
                assert np.abs(out[a].astype(np.float64) - out1[a]).max() < 1e-7
            # chunked relaunches of the persistent kernel on the multi-rank path: bit-identical
            m.solve(sc, Params(octree_levels=L, tolerance=1e-10, check_every=5))
            assert np.array_equal(m.solution(), x)
    finally:
        m.close()
        single.close()


def test_multi_rank_solve_refuses_a_scene_outside_the_reference_contract():
    """Liquid at the grid boundary puts the column -3 into the reference's matrix (tests/test_reference_fuzz.py).  Only the rank that
    owns such a row sees it; the ranks sum their assembly error flags so that ALL of them leave with AVS_ERR_UNSUPPORTED (none waits
    for a peer that left), and the group solves a valid scene afterwards exactly like a single context."""
    import importlib.util
    from pathlib import Path
    from adaptiveviscositysolver_b200.solver import AvsError, MultiSolver, Params, Solver
    spec = importlib.util.spec_from_file_location("fuzz_reference_pin", Path(__file__).resolve().parent.parent / "scripts" / "fuzz_reference_pin.py")
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    sc, op, _ = fz.fuzz_case(fz.OUT_OF_CONTRACT_SEED)
    bad = Params(octree_levels=op.octree_levels, tolerance=op.tolerance, dt=op.dt, use_enhanced_gradients=op.use_enhanced_gradients,
                 do_apply_solid_weights=op.do_apply_solid_weights, fine_bandwidth=op.fine_bandwidth,
                 number_super_samples=op.number_super_samples, extrapolation=op.extrapolation)
    m = MultiSolver(_devices(2))
    single = Solver(device=0)
    try:
        with pytest.raises(AvsError) as e:
            m.solve(sc, bad, [v.data.copy() for v in sc.vel])
        assert e.value.status == -10 and "boundary of the grid" in str(e.value)      # AVS_ERR_UNSUPPORTED
        good, p = sphere_drop(32, 11, noise=0.01), Params(octree_levels=4, tolerance=1e-10)
        out, out1 = [v.data.copy() for v in good.vel], [v.data.copy() for v in good.vel]
        info, info1 = m.solve(good, p, out), single.solve(good, p, out1)
        assert info.octree_dofs == info1.octree_dofs and abs(info.iterations - info1.iterations) <= 1 and info.error < 1e-10
        for a in range(3):
            assert np.abs(out[a].astype(np.float64) - out1[a]).max() < 1e-7
    finally:
        m.close()
        single.close()


def test_multi_api_rejects_device_pointers_and_bad_arguments():
    import torch
    from adaptiveviscositysolver_b200 import _lib
    from adaptiveviscositysolver_b200.solver import MultiSolver, Params
    from adaptiveviscositysolver_b200.scenes import SampledField, Scene
    with pytest.raises(_lib.AvsError):
        MultiSolver([])
    m = MultiSolver(_devices(2))
    try:
        sc = sphere_drop(32, 10)
        dev = torch.device("cuda", 0)
        mv = lambda f: f if f.data is None else SampledField(torch.from_numpy(f.data).to(dev), f.org, f.dx, f.constant)  # noqa: E731
        dsc = Scene(sc.res, sc.origin, sc.dx, mv(sc.surface), [mv(v) for v in sc.vel], [mv(v) for v in sc.face_weights], mv(sc.viscosity),
                    mv(sc.density), mv(sc.collision), [mv(v) for v in sc.collision_vel])
        with pytest.raises(_lib.AvsError) as e:
            m.solve(dsc, Params(octree_levels=3))
        assert e.value.status == -10          # AVS_ERR_UNSUPPORTED
        assert m.solve(sc, Params(octree_levels=3, tolerance=1e-6)).iterations > 0    # still usable afterwards
    finally:
        m.close()
