"""The Houdini-side shim of this repository (integration/hdk/HDK_AdaptiveViscosityB200.{h,cpp}), COMPILED and RUN: built against
the HDK stand-ins of oracle/mock_hdk (the same headers that compile the reference itself) and linked to libavs_b200.so, then driven
like Houdini drives a DOP -- named fields on a SIM_Object, options behind the GET_DATA_FUNC getters, solveGasSubclass on one
thread -- and compared with the REFERENCE's solveGasSubclass (oracle/_ref/libavs_ref.so) on the same fields."""
import numpy as np
import pytest

from adaptiveviscositysolver_b200 import scenes
from oracle import avs_oracle as orc
from oracle import avs_ref as ref

pytestmark = pytest.mark.gpu


def _reference(sc, p):
    if ref.available():
        return ref.RefRun(sc, p)
    return orc.OracleRun(sc, p)      # snapshot without the prebuilt reference library: the restated oracle (pinned to it on the CPU side)


@pytest.mark.parametrize("gpus", [1, 2])
@pytest.mark.parametrize("case", ["sphere64_l5", "solid_ground", "buckling"])
def test_shim_equals_reference_dop(case, gpus):
    if case == "sphere64_l5":
        sc, p = scenes.sphere_drop(64, 26, noise=0.01), orc.OracleParams(octree_levels=5, tolerance=1e-10)
    elif case == "solid_ground":
        sc = scenes.sphere_drop(32, 9, center=(0.5, 0.34, 0.5), ground_height=0.125, ground_velocity=(0.1, 0.0, -0.2))
        p = orc.OracleParams(octree_levels=3, tolerance=1e-10)
    else:
        # dx = 2^-9: the DOP boundary carries the voxel size and the field origins as fp32 (UT_Vector3), and this scene has its
        # ground plane exactly ON a grid plane.  With dx = 0.002 (not an fp32 number) the shim's fp32 dx moves the y = 0 faces by
        # 8e-10 off the plane, the centre sub-sample of their solid weights flips side, and the system is a (legitimately)
        # different one: 925 instead of 1332 iterations -- reproduced with the oracle fed the fp32 dx (tests/test_buckling.py).
        sc, p = scenes.buckling_sheet(frame=6, dx=2.0 ** -9), orc.OracleParams(octree_levels=4, tolerance=1e-10, dt=1.0 / 120.0)
    S = ref.ShimRun(sc, p, gpus=gpus)          # gpus = 2 on a one-GPU box: both ranks share the GPU (ordinals wrap)
    assert S.returned_true and not S.errors, S.errors
    R = _reference(sc, p)
    info = S.info()
    assert int(info["octree DOFS"]) == R.n_face and int(info["regular DOFs"]) == R.regular_dofs
    assert abs(int(info["iterations"]) - R.iterations) <= max(2, R.iterations // 50)
    for a in range(3):
        so, ro = S.out_velocity(a), R.out_velocity(a)
        assert np.abs(so.astype(np.float64) - ro.astype(np.float64)).max() < 1e-6
        reg = R.regular_index(a)
        untouched = (reg == orc.UNASSIGNED) | (reg == orc.OUTSIDE)
        assert np.array_equal(so[untouched], sc.vel[a].data[untouched])
    # doPrintOctree: same geometry dump
    got = sorted(map(tuple, np.column_stack(S.octree_points()).tolist()))
    want = sorted(map(tuple, np.column_stack(R.octree_points()).tolist()))
    assert got == want


def test_shim_only_print_octree_returns_early():
    sc = scenes.sphere_drop(32, 10)
    S = ref.ShimRun(sc, orc.OracleParams(octree_levels=4), octree_only=True)
    assert S.returned_true and not S.errors
    assert S.extra_info == ""                  # no solve happened (AV.cpp:292-293)
    for a in range(3):
        assert np.array_equal(S.out_velocity(a), sc.vel[a].data)
    assert S.octree_points()[0].shape[0] > 0
