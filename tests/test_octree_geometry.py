"""Octree geometry dump (doPrintOctree / onlyPrintOctree, AV.cpp:283-294; HDK_OctreeGrid::outputOctreeGeometry,
OG.cpp:245-308): one point (P, pscale, octreeLevel) per ACTIVE cell -- SURVEY.md section 8f rank 4."""
import numpy as np
import pytest

from adaptiveviscositysolver_b200.scenes import sphere_drop
from oracle import avs_oracle as orc


def _sorted(pos, pscale, level):
    order = np.lexsort((pos[:, 0], pos[:, 1], pos[:, 2], level))
    return pos[order], pscale[order], level[order]


def test_oracle_points_are_the_active_cells():
    sc = sphere_drop(64, 26)
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=6), stop_after_stage=3)
    pos, pscale, level = ref.octree_points()
    assert pos.dtype == np.float32 and pscale.dtype == np.float32 and level.dtype == np.int32
    assert level.max() == ref.levels - 1
    for l in range(ref.levels):
        lab = ref.labels(l)
        m = level == l
        assert m.sum() == (lab == orc.ACTIVE).sum() > 0
        h = np.float32(sc.dx * 2 ** l)
        assert np.all(pscale[m] == h)
        # positions are cell centres: (P - origin) / h - 1/2 is the integer cell index, and that cell is ACTIVE
        idx = np.rint(pos[m].astype(np.float64) / h - 0.5).astype(int)
        assert np.abs(pos[m] / h - 0.5 - idx).max() < 1e-4
        assert np.all(lab[idx[:, 2], idx[:, 1], idx[:, 0]] == orc.ACTIVE)
    # the ACTIVE cells of all levels tile the refined region exactly once (activeCountUnitTest, OG.cpp:984-1080):
    # their volumes add up to the volume of the level-0 cells that are not INACTIVE
    vol = (pscale.astype(np.float64) ** 3).sum()
    assert vol == pytest.approx((ref.labels(0) != orc.INACTIVE).sum() * sc.dx ** 3, rel=1e-12)


@pytest.fixture(scope="module")
def solver():
    from adaptiveviscositysolver_b200.solver import Solver
    s = Solver(device=0)
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw,levels", [(dict(n=64, radius_cells=26), 6), (dict(n=32, radius_cells=10), 1),
                                       (dict(n=64, radius_cells=14, res=(48, 64, 40), center=(0.375, 0.5, 0.3125)), 5)])
def test_gpu_octree_points_match_oracle(solver, kw, levels):
    from adaptiveviscositysolver_b200.solver import Params
    sc = sphere_drop(**kw)
    info = solver.build_octree(sc, Params(octree_levels=levels))          # onlyPrintOctree path: stages 1-3 only
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=levels), stop_after_stage=3)
    assert info.levels == ref.levels
    g, o = _sorted(*solver.octree_points()), _sorted(*ref.octree_points())
    for a, b in zip(g, o):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    # the same dump is available after a full solve (doPrintOctree without onlyPrintOctree)
    solver.solve(sc, Params(octree_levels=levels))
    g2 = _sorted(*solver.octree_points())
    for a, b in zip(g2, o):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_operator_only_print_octree():
    from adaptiveviscositysolver_b200.solver import HDK_AdaptiveViscosity, SIM_Object
    sc = sphere_drop(32, 10)
    before = [v.data.copy() for v in sc.vel]
    op = HDK_AdaptiveViscosity(octreeLevels=3, doPrintOctree=True, onlyPrintOctree=True)
    obj = SIM_Object.from_scene(sc)
    assert op.solveGasSubclass(None, obj, 0.0, 1 / 24) is True
    pos, pscale, level = obj.geometry["octreeGeometry"]
    assert len(pos) == len(pscale) == len(level) > 0
    for a in range(3):
        assert np.array_equal(sc.vel[a].data, before[a])                 # early return: velocity untouched (AV.cpp:292-293)
    op2 = HDK_AdaptiveViscosity(octreeLevels=3, doPrintOctree=True)
    assert op2.solveGasSubclass(None, obj, 0.0, 1 / 24) is True
    assert len(obj.geometry["octreeGeometry"][0]) == len(pos) and op2.info.iterations > 0
