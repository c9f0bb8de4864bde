"""The CG's execution modes on one GPU: persistent cooperative kernel (default), chunked relaunches of it (`check_every`,
the cancel-polling granularity), fp32 (USESINGLEPRECISION), and -- in a subprocess, because the mode is read once per
process -- the per-launch loop (AVS_CG_MODE=launch) with and without the TMA-staged SpMV (AVS_SPMV_TMA=1)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from adaptiveviscositysolver_b200.scenes import sphere_drop
from oracle import avs_oracle as orc
from tests.util import perm_gpu_to_oracle

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    from adaptiveviscositysolver_b200.solver import Solver
    s = Solver(device=0)
    yield s
    s.close()


def test_persistent_kernel_is_one_launch_and_reports_its_phases(solver):
    from adaptiveviscositysolver_b200.solver import Params
    sc = sphere_drop(64, 26, noise=0.01)
    info = solver.solve(sc, Params(octree_levels=5, tolerance=1e-8))
    assert info.iterations > 50 and info.error < 1e-8
    assert info.cg_kernel_launches == 1 and info.cg_kernel_ms > 0
    assert info.spmv_launches == info.iterations + 1          # the converging iteration runs its SpMV and x,r phase, then breaks
    assert 0 < info.spmv_ms <= info.cg_kernel_ms * 1.05
    assert info.kernel_launches < 400                         # no per-iteration launches


def test_chunked_relaunch_is_bit_identical(solver):
    """check_every = iterations per cooperative launch: the relaunched kernel continues from the device-resident scalars."""
    from adaptiveviscositysolver_b200.solver import Params
    sc = sphere_drop(64, 26, noise=0.01)
    one = solver.solve(sc, Params(octree_levels=5, tolerance=1e-9))
    x_one = solver.solution()
    many = solver.solve(sc, Params(octree_levels=5, tolerance=1e-9, check_every=7))
    assert many.iterations == one.iterations and many.error == one.error
    assert many.cg_kernel_launches == one.iterations // 7 + 1 and one.cg_kernel_launches == 1
    assert np.array_equal(solver.solution(), x_one)


def test_single_precision_multi_cta(solver):
    """USESINGLEPRECISION (HDK_Utilities.h:25-37) through the persistent kernel on a system that spans many CTAs."""
    from adaptiveviscositysolver_b200.solver import Params
    sc = sphere_drop(64, 26)
    info = solver.solve(sc, Params(octree_levels=5, tolerance=1e-3, single_precision=True))
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=5, tolerance=1e-9))
    perm = perm_gpu_to_oracle(solver.keys(), ref.face_keys())
    assert info.error < 1e-3 and 10 < info.iterations < 2500
    assert np.abs(solver.solution() - ref.solution()[perm]).max() < 5e-2


_WORKER = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
from adaptiveviscositysolver_b200.scenes import sphere_drop
from adaptiveviscositysolver_b200.solver import Params, Solver
from oracle import avs_oracle as orc
from tests.util import perm_gpu_to_oracle
EXPECT_LAUNCH_MODE = %r
sc = sphere_drop(64, 26, noise=0.01)
s = Solver(device=0)
out = [v.data.copy() for v in sc.vel]
info = s.solve(sc, Params(octree_levels=5, tolerance=1e-10), out)
ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=5, tolerance=1e-10))
perm = perm_gpu_to_oracle(s.keys(), ref.face_keys())
err = float(np.abs(s.solution() - ref.solution()[perm]).max())
oerr = max(float(np.abs(out[a] - ref.out_velocity(a)).max()) for a in range(3))
assert (info.cg_kernel_launches == 0) == (EXPECT_LAUNCH_MODE), info.cg_kernel_launches   # the requested loop really ran
assert abs(info.iterations - ref.iterations) <= 2 and err < 1e-6 and oerr < 1e-6, (info.iterations, ref.iterations, err, oerr)
print("ok", info.iterations, err)
"""


@pytest.mark.parametrize("tma", [False, True])
def test_per_launch_cg_mode(tma):
    env = dict(os.environ)
    env["AVS_CG_MODE"] = "launch"
    if tma:
        env["AVS_SPMV_TMA"] = "1"
    else:
        env.pop("AVS_SPMV_TMA", None)
    r = subprocess.run([sys.executable, "-c", _WORKER % (str(ROOT), True)], capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().startswith("ok")


@pytest.mark.parametrize("variant", ["v2", "v2_pf", "v2_ring", "v2_ring4", "v2_call", "v1"])
def test_persistent_kernel_variants(variant):
    """Both persistent CG kernels on one GPU (the default there is round 1's k_cg_persistent; k_cg_persistent2 is the multi-GPU
    default and is forced here with AVS_PCG_KERNEL=v2), and the slice loops selectable with AVS_SPMV_MODE inside k_cg_persistent2
    and in the stand-alone SpMV: L2 prefetch, matrix stream staged through a per-warp cp.async ring (two depths), slice loop as a
    separate function."""
    env = dict(os.environ)
    env.pop("AVS_CG_MODE", None)
    env.pop("AVS_SPMV_TMA", None)
    env.pop("AVS_SPMV_MODE", None)
    env["AVS_PCG_KERNEL"] = "v1" if variant == "v1" else "v2"
    if "_" in variant:
        env["AVS_SPMV_MODE"] = variant.split("_", 1)[1]
    r = subprocess.run([sys.executable, "-c", _WORKER % (str(ROOT), False)], capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().startswith("ok")
