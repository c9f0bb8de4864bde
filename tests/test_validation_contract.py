"""The eleven validation errors at the top of solveGasSubclass (HDK_AdaptiveViscosity.cpp:152-229) -- the part of the reference's
host side that stays in front of the C-ABI.  For each way of breaking the simulation object, the REFERENCE (its own sources,
oracle/_ref/libavs_ref.so), this repository's C++ DOP shim (integration/hdk/HDK_AdaptiveViscosityB200.cpp, compiled against the same
HDK stand-ins) and its Python mirror (adaptiveviscositysolver_b200.solver.HDK_AdaptiveViscosity) must report the same message, in
the same precedence, and return false.  All three validate before they touch a GPU, so this runs on the CPU; with valid fields and
no GPU the shim and the mirror must fail loudly (there is no CPU path)."""
import copy

import numpy as np
import pytest

from adaptiveviscositysolver_b200 import scenes
from adaptiveviscositysolver_b200.scenes import SampledField
from adaptiveviscositysolver_b200.solver import HDK_AdaptiveViscosity, SIM_Object
from oracle import avs_oracle as orc
from oracle import avs_ref as ref

pytestmark = pytest.mark.skipif(ref.build() is None, reason="oracle/_ref/libavs_ref.so not built and /root/reference not present to build it")

# (what is done to the object, the field, the reference's message AV.cpp:line)
CASES = [
    ("remove", "velocity", "Liquid velocity field missing"),                              # :154
    ("unstagger", "velocity", "Liquid velocity field must be a staggered grid"),          # :159
    ("remove", "faceWeights", "Face weights field missing"),                              # :165
    ("misalign", "faceWeights", "Face weights must align with velocity samples"),         # :171
    ("remove", "collision", "Solid surface field missing"),                               # :177
    ("remove", "collisionvel", "Solid velocity field missing"),                           # :185
    ("remove", "surface", "Liquid surface field is missing"),                             # :191
    ("remove", "viscosity", "Viscosity field is missing"),                                # :207
    ("misalign", "viscosity", "Viscosity field must align with the surface volume"),      # :212
    ("remove", "density", "Density field is missing"),                                    # :222
    ("misalign", "density", "Density field must align with the surface volume"),          # :227
]
MIRROR_KEYS = {"velocity": "vel", "faceWeights": "surfaceweights", "collision": "collision", "collisionvel": "collisionvel",
               "surface": "surface", "viscosity": "viscosity", "density": "massdensity"}


def _scene():
    # dense viscosity and density so that they can be misaligned (a constant field is aligned with everything)
    return scenes.sphere_drop(16, 5, variable_viscosity=True, variable_density=True)


def _mirror_object(sc, op, field):
    obj = SIM_Object.from_scene(copy.deepcopy(sc))
    key = MIRROR_KEYS[field]
    if op == "remove":
        del obj.fields[key]
    elif op == "misalign":     # the same tampering as the harness: the field's grid shifted by one voxel along x
        f = obj.fields[key]
        shift = lambda c: SampledField(c.data, (c.org[0] + c.dx, c.org[1], c.org[2]), c.dx, c.constant)
        obj.fields[key] = [shift(c) for c in f] if isinstance(f, list) else shift(f)
    else:                      # "unstagger": component 0 sampled at the cell centres
        v = list(obj.fields[key])
        nz, ny, nx = v[0].data.shape
        v[0] = SampledField(np.ascontiguousarray(v[0].data[:, :, : nx - 1]), sc.surface.org, sc.dx)
        obj.fields[key] = v
    return obj


@pytest.mark.parametrize("op,field,message", CASES)
def test_reference_shim_and_mirror_report_the_same_validation_error(op, field, message):
    sc, p = _scene(), orc.OracleParams(octree_levels=2)
    R = ref.RefRun(sc, p, tamper=[(op, field)])
    assert not R.returned_true and R.errors == [message]
    S = ref.ShimRun(sc, p, tamper=[(op, field)])
    assert not S.returned_true and S.errors == [message]
    m = HDK_AdaptiveViscosity(octreeLevels=2)
    assert m.solveGasSubclass(None, _mirror_object(sc, op, field), 0.0, 1.0 / 24.0) is False and m.errors == [message]


def test_precedence_is_the_references():
    """Two problems at once: the one the reference tests first is the one reported."""
    sc, p = _scene(), orc.OracleParams(octree_levels=2)
    both = [("remove", "density"), ("misalign", "faceWeights")]
    R, S = ref.RefRun(sc, p, tamper=both), ref.ShimRun(sc, p, tamper=both)
    assert R.errors == S.errors == ["Face weights must align with velocity samples"]
    obj = _mirror_object(sc, "misalign", "faceWeights")
    del obj.fields["massdensity"]
    m = HDK_AdaptiveViscosity(octreeLevels=2)
    assert m.solveGasSubclass(None, obj, 0.0, 1.0 / 24.0) is False and m.errors == R.errors


def test_valid_fields_without_a_gpu_fail_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sc, p = _scene(), orc.OracleParams(octree_levels=2)
    S = ref.ShimRun(sc, p)
    assert not S.returned_true and len(S.errors) == 1 and "no CUDA device" in S.errors[0]
    m = HDK_AdaptiveViscosity(octreeLevels=2)
    assert m.solveGasSubclass(None, SIM_Object.from_scene(sc), 0.0, 1.0 / 24.0) is False
    assert len(m.errors) == 1 and "no CUDA device" in m.errors[0]
