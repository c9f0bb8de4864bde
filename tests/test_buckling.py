"""C5 -- the prescribed-geometry buckling-sheet sequence (BASELINE.json configs[4]; geometry of
Scenes/viscousBuckling.hip): variable viscosity, a solid ground plane, a thin folded sheet.

CPU leg: the scene generator's own invariants and the oracle's known answers on it.
GPU leg: the CUDA path through the C-ABI against the oracle, frame by frame.
"""
import numpy as np
import pytest

from adaptiveviscositysolver_b200.scenes import buckling_sequence, buckling_sheet
from oracle import avs_oracle as orc
from tests.util import csr_permuted, perm_gpu_to_oracle

DT = 1.0 / 120.0


def _volume(sc):
    return sum(float(w.data.sum(dtype=np.float64)) for w in sc.face_weights) / 3.0 * sc.dx ** 3


def test_sequence_geometry():
    """10 frames, contact inside the sequence, liquid volume conserved by the fold (the hip's 0.1 x 0.1 x 0.01 box)."""
    frames = list(buckling_sequence(10, dx=0.002))
    assert len(frames) == 10
    consumed = [f.meta["consumed"] for f in frames]
    assert consumed[0] == 0.0 and consumed[-1] > 0.02 and all(b >= a for a, b in zip(consumed, consumed[1:]))
    for f in frames:
        assert f.res == frames[0].res and f.origin == frames[0].origin       # one fixed domain for the sequence
        assert abs(_volume(f) - 1e-4) < 0.02e-4
        assert f.viscosity.data is not None and f.collision.data is not None
        # ground plane: collision > 0 exactly below y = 0
        ys = f.collision.org[1] + f.dx * np.arange(f.res[1])
        assert np.array_equal(f.collision.data[0, :, 0] > 0, ys < 0)


def test_free_fall_frame_is_a_fixed_point():
    """K3 (translation invariance) with variable viscosity: before contact the velocity is one constant vector,
    so rhs = M u and Eigen's CG returns after 0 iterations."""
    sc = buckling_sheet(0, dx=0.002)
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=4, tolerance=1e-8, dt=DT))
    assert ref.iterations == 0 and ref.n_face > 10000
    assert np.array_equal(ref.solution(), ref.x0())


def test_folded_frame_system_properties():
    """K1/K2 on a folded frame touching the ground: symmetric, positive diagonal, SOLIDBOUNDARY faces present,
    viscosity really varies across the rows."""
    sc = buckling_sheet(6, dx=0.002)
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=4, tolerance=1e-8, dt=DT))
    A = ref.scipy_matrix()
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    assert A.diagonal().min() > 0
    assert ref.iterations > 100 and ref.error < 1e-8
    assert any((ref.regular_index(a) == orc.SOLIDBOUNDARY).any() for a in range(3))
    # x^T A x > 0 on random vectors (SPD)
    rng = np.random.default_rng(5)
    for _ in range(3):
        v = rng.normal(size=ref.n_face)
        assert v @ (A @ v) > 0


def test_ground_plane_on_a_grid_plane_is_sensitive_to_fp32_voxel_size():
    """The DOP boundary (integration/hdk) carries the voxel size as fp32.  At dx = 0.002 -- not an fp32 number -- the
    ground plane y = 0 of this scene lies exactly on a grid plane, the fp32 dx moves those faces 8e-10 off it, and the centre
    sub-sample of their solid weights changes side: same DOFs, a different (equally valid) system.  At dx = 2^-9 fp32 and fp64
    describe the same grid; tests/test_gpu_hdk_shim.py therefore runs the buckling frame at 2^-9."""
    import copy
    p = orc.OracleParams(octree_levels=4, tolerance=1e-10, dt=DT)

    def as_fp32_grid(sc):
        out = copy.deepcopy(sc)
        dxf = float(np.float32(sc.dx))
        out.dx = dxf
        centre = [o + 0.5 * dxf for o in sc.origin]     # the stand-in's UT_Vector3 keeps the origin exact; dx is fp32
        for f in (out.surface, out.viscosity, out.collision):
            if f.data is not None:
                f.dx, f.org = dxf, tuple(centre)
        for a in range(3):
            org = list(centre)
            org[a] = sc.origin[a]
            for f in (out.vel[a], out.face_weights[a]):
                f.dx, f.org = dxf, tuple(org)
        return out

    sc = buckling_sheet(6, dx=0.002)
    exact, f32 = orc.OracleRun(sc, p), orc.OracleRun(as_fp32_grid(sc), p)
    assert exact.n_face == f32.n_face and exact.regular_dofs == f32.regular_dofs
    assert abs(exact.iterations - f32.iterations) > exact.iterations // 10       # 1332 vs 925
    sc = buckling_sheet(6, dx=2.0 ** -9)
    exact, f32 = orc.OracleRun(sc, p), orc.OracleRun(as_fp32_grid(sc), p)
    assert exact.iterations == f32.iterations and np.array_equal(exact.solution(), f32.solution())


@pytest.fixture(scope="module")
def solver():
    from adaptiveviscositysolver_b200.solver import Solver
    s = Solver(device=0)
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("frame,dx,levels", [(0, 0.002, 4), (3, 0.002, 4), (9, 0.002, 4), (5, 0.00125, 4)])
def test_gpu_buckling_parity(solver, frame, dx, levels):
    from adaptiveviscositysolver_b200.solver import Params
    sc = buckling_sheet(frame, dx=dx)
    gp = Params(octree_levels=levels, tolerance=1e-10, dt=DT)
    op = orc.OracleParams(octree_levels=levels, tolerance=1e-10, dt=DT)
    out = [v.data.copy() for v in sc.vel]
    info = solver.solve(sc, gp, out)
    ref = orc.OracleRun(sc, op)
    assert (info.levels, info.octree_dofs, info.edge_dofs, info.center_dofs, info.regular_dofs, info.nnz) == \
        (ref.levels, ref.n_face, ref.n_edge, ref.n_center, ref.regular_dofs, ref.nnz)
    assert np.array_equal(solver.center_weights(), ref.center_weights())
    for l in range(ref.levels):
        assert np.array_equal(solver.labels(l), ref.labels(l))
    for a in range(3):
        rg, ro = solver.regular_labels(a), ref.regular_index(a)
        assert np.array_equal(np.minimum(rg, 0), np.minimum(ro, 0))
    perm = perm_gpu_to_oracle(solver.keys(), ref.face_keys())
    ptr, col, val, rhs, x0 = solver.system()
    n = ref.n_face
    A, Ao = csr_permuted(ptr, col, val, perm, n), ref.scipy_matrix()
    A.sort_indices(); Ao.sort_indices()
    assert np.array_equal(A.indptr, Ao.indptr) and np.array_equal(A.indices, Ao.indices)
    scale = abs(Ao).max()
    assert np.allclose(A.data, Ao.data, rtol=1e-12, atol=1e-12 * scale)
    b = np.empty(n); b[perm] = rhs
    assert np.allclose(b, ref.rhs(), rtol=1e-12, atol=1e-12 * abs(ref.rhs()).max())
    # solve: the system is stiff (dt mu / dx^2 ~ 1e6 rho), both CGs run to 1e-10
    assert abs(info.iterations - ref.iterations) <= max(2, ref.iterations // 50)
    x, xo = solver.solution(), ref.solution()[perm]
    assert np.abs(x - xo).max() < 1e-6
    assert info.interpolated_faces == ref.interpolated_faces
    for a in range(3):
        oo = ref.out_velocity(a)
        assert np.abs(out[a].astype(np.float64) - oo.astype(np.float64)).max() < 1e-6
        reg = ref.regular_index(a)
        untouched = (reg == orc.UNASSIGNED) | (reg == orc.OUTSIDE)
        assert np.array_equal(out[a][untouched], sc.vel[a].data[untouched])
