"""Random scenes: the restated oracle against the REFERENCE'S OWN CODE (oracle/_ref/libavs_ref.so, see tests/test_reference_pin.py).

scripts/fuzz_reference_pin.py draws the scenes and options; it was run over seeds 0-299 when this file was written (every seed whose
scene respects the reference's contract agrees: labels, weights, numbering, matrix, right-hand side and restricted velocity bit for
bit; a liquid that reaches the grid boundary is outside that contract -- the reference's own debug checks fail on it).  This file
replays a fixed subset in the CPU suite; tests/test_gpu_reference.py holds the CUDA path to the reference on seeds of the same generator.
"""
import importlib.util
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import avs_oracle as orc
from oracle import avs_ref as ref

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(ref.build() is None, reason="oracle/_ref/libavs_ref.so not built and /root/reference not present to build it")

_spec = importlib.util.spec_from_file_location("fuzz_reference_pin", ROOT / "scripts" / "fuzz_reference_pin.py")
fz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fz)

# solids (plane / sphere, own collision grid, solid weights), variable viscosity / density, non-cubic grids, shifted origins,
# 1-4 built levels, every option away from its default somewhere
# 4017 / 4022 / 4046: variant 4, the solid's velocity as sampled fields on a grid of their own (tests/test_zz_gpu_late_additions.py)
SEEDS = [0, 4, 9, 15, 25, 29, 34, 43, 48, 57, 61, 70, 73, 77, 89, 100, 4017, 4022, 4046]


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_equals_reference_on_a_random_scene(seed):
    sc, p, desc = fz.fuzz_case(seed)
    levels, n, iterations = fz.compare(sc, p)
    assert n > 0 and iterations >= 0, desc


FINGERPRINTS = {4: "78f7dd0c1e74c147", 23: "0ee5fad76c7d641f", 43: "a45f814aea087ffc", 135: "0dcb907ea1449a40", 192: "d1d035df3c89a34c",
                1011: "b9a76be6749841c1", 2005: "14134869a31ba6d4", 3000: "df1b00fb6c5dbaf9", 4022: "aa1c5d86a7598ab3"}


def test_seeds_still_mean_the_same_scenes():
    """The GPU tests (tests/test_gpu_reference.py, test_gpu_inprocess_multi.py) and the recorded sweeps name scenes by seed: an edit of
    the generator must not change what a seed draws."""
    import hashlib
    for seed, want in FINGERPRINTS.items():
        sc, p, _ = fz.fuzz_case(seed)
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(sc.surface.data).tobytes())
        h.update(np.ascontiguousarray(sc.vel[1].data).tobytes())
        h.update(repr((sc.res, sc.dx, sc.origin, p)).encode())
        if sc.collision_vel[0].data is not None:       # variant 4: the sampled solid velocity is part of what the seed means
            h.update(np.ascontiguousarray(sc.collision_vel[0].data).tobytes())
        assert h.hexdigest()[:16] == want, seed


def test_generator_is_deterministic():
    a, pa, da = fz.fuzz_case(43)
    b, pb, db = fz.fuzz_case(43)
    assert da == db and pa == pb and np.array_equal(a.surface.data, b.surface.data) and np.array_equal(a.vel[1].data, b.vel[1].data)


def test_liquid_at_the_grid_boundary_is_outside_the_reference_contract():
    """What the CUDA library refuses with AVS_ERR_UNSUPPORTED (tests/test_gpu_reference.py): on such a scene the reference's debug
    build trips its own checks, and its release build -- like the restated oracle, entry for entry -- emits matrix columns < 0
    (getEdgeStressFaces appends a parent face it only asserts to be a DOF, AV.cpp:1886-1894)."""
    seed = fz.OUT_OF_CONTRACT_SEED
    if ref.REFERENCE_SOURCES.exists():
        r = subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref-debug"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        d = subprocess.run([sys.executable, str(ROOT / "scripts" / "fuzz_reference_pin.py"), "--debug", str(seed)], capture_output=True, text=True, timeout=600)
        assert d.returncode != 0 and "Assertion" in d.stderr, d.stdout[-500:] + d.stderr[-500:]
    c = subprocess.run([sys.executable, str(ROOT / "scripts" / "fuzz_reference_pin.py"), "--compare", str(seed)], capture_output=True, text=True, timeout=600)
    assert c.returncode == 0 and "OUT-OF-RANGE COLUMNS" in c.stdout, c.stdout[-500:] + c.stderr[-500:]
    sc, p, _ = fz.fuzz_case(seed)
    assert orc.OracleRun(sc, p, stop_after_stage=9).csr()[1].min() < 0     # assembly only: no CG on a matrix with column -3


@pytest.mark.parametrize("seed", [4, 7, 10, 79, 163])
def test_single_precision_oracle_equals_reference_single_precision_build_on_a_random_scene(seed):
    """USESINGLEPRECISION (BASELINE configs[2]) on multi-level random scenes: scripts/fuzz_reference_f32.py (swept over seeds 0-299,
    profiles/r2_fuzz.md) -- numbering and sparsity identical, matrix within one float rounding, iteration counts and solutions where
    both float runs converge."""
    if ref.build_f32() is None:
        pytest.skip("/root/reference not present (the fp32 reference build is made on demand)")
    spec = importlib.util.spec_from_file_location("fuzz_reference_f32", ROOT / "scripts" / "fuzz_reference_f32.py")
    f32 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(f32)
    sc, p, desc = fz.fuzz_case(seed)
    p.single_precision = True
    p.tolerance = max(p.tolerance, 1e-5)
    r = f32.compare_f32(sc, p)
    assert r is not None and r["levels"] >= 2 and r["converged"] and r["dsol"] < 2e-3, (desc, r)
