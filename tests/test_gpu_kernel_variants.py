"""Round-2 kernel rewrites against the kernels they replace, BIT FOR BIT on one GPU:

* the split assembly (k_assemble_simple for the level-0 rows without T-junctions / solids, k_assemble with the hashed row
  accumulator for the rest) against the single-pass assembly with the linear-search accumulator (AVS_ASM=generic,
  AVS_ASM_ROW=linear): same rows, same entry order, same roundings -- hence the same CG iterates and output velocities;
* the tile-culled level-0 labelling and node pyramid (memset + classification only on the 16^3 tiles that can hold anything but
  the default label) against the dense sweeps (AVS_LABELS=dense): same labels, same numbering, same interpolated velocities.

Scenes cover variable viscosity / density, solids (boundary terms), doApplySolidWeights, enhanced gradients off, a ragged
non-power-of-two grid, a uniform grid and BASELINE configs[1].  A variant is chosen once per process (environment), so each
configuration runs in its own subprocess and prints the sha256 of everything the pipeline produced; the digests must be identical."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

_WORKER = r"""
import hashlib, json, sys
import numpy as np
sys.path.insert(0, %r)
from adaptiveviscositysolver_b200 import scenes
from adaptiveviscositysolver_b200.solver import Params, Solver

def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()

CASES = {
    "sphere64_l5_noise": (scenes.sphere_drop(64, 26, noise=0.01), dict(octree_levels=5, tolerance=1e-9)),
    "padded_varmu": (scenes.sphere_drop(64, 14, res=(48, 64, 40), center=(0.375, 0.5, 0.3125), variable_viscosity=True, variable_density=True),
                     dict(octree_levels=5, tolerance=1e-9)),
    "solid_ground_moving": (scenes.sphere_drop(32, 9, center=(0.5, 0.34, 0.5), ground_height=0.125, ground_velocity=(0.1, 0.0, -0.2)),
                            dict(octree_levels=3, tolerance=1e-9)),
    "solid_weights": (scenes.sphere_drop(32, 9, center=(0.5, 0.34, 0.5), ground_height=0.12, ground_velocity=(0.0, 0.05, 0.0)),
                      dict(octree_levels=3, tolerance=1e-9, do_apply_solid_weights=True)),
    "no_enhanced_gradients": (scenes.sphere_drop(32, 11), dict(octree_levels=4, tolerance=1e-9, use_enhanced_gradients=False)),
    "buckling_ragged_66x84x35": (scenes.buckling_sheet(frame=6, dx=0.002), dict(octree_levels=4, tolerance=1e-9, dt=1.0 / 120.0)),
    "uniform32": (scenes.sphere_drop(32, 10), dict(octree_levels=1, tolerance=1e-9)),
    "c2_literal_128": (scenes.sphere_drop(128, 56), dict(octree_levels=5, tolerance=1e-8)),
}
s = Solver(device=0)
out = {}
for name, (sc, kw) in CASES.items():
    vel = [v.data.copy() for v in sc.vel]
    info = s.solve(sc, Params(**kw), vel)
    ptr, col, val, rhs, x0 = s.system()
    d = {"labels": digest(*[s.labels(l) for l in range(info.levels)], s.center_weights(), *[s.edge_weights(a) for a in range(3)],
                          *[s.regular_labels(a) for a in range(3)],
                          *[s.face_labels(l, a) for l in range(info.levels) for a in range(3)],
                          *[s.edge_labels(l, a) for l in range(info.levels) for a in range(3)],
                          *[s.center_labels(l) for l in range(info.levels)]),
         "keys": digest(s.keys()), "matrix": digest(ptr, col, val), "rhs": digest(rhs), "x0": digest(x0),
         "solution": digest(s.solution()), "velocity": digest(*vel),
         "counts": [int(info.octree_dofs), int(info.edge_dofs), int(info.center_dofs), int(info.regular_dofs), int(info.nnz), int(info.iterations),
                    int(info.interpolated_faces)]}
    out[name] = d
print("DIGESTS " + json.dumps(out))
"""


def _run(extra_env):
    env = dict(os.environ)
    for k in ("AVS_ASM", "AVS_LABELS", "AVS_ASM_ROW", "AVS_CG_MODE", "AVS_PCG_KERNEL", "AVS_SPMV_MODE"):
        env.pop(k, None)
    env.update(extra_env)
    r = subprocess.run([sys.executable, "-c", _WORKER % str(ROOT)], capture_output=True, text=True, timeout=900, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DIGESTS ")][-1]
    return json.loads(line[len("DIGESTS "):])


@pytest.fixture(scope="module")
def default_digests():
    return _run({})


@pytest.mark.parametrize("variant", [{"AVS_ASM": "generic", "AVS_ASM_ROW": "linear"}, {"AVS_LABELS": "dense"}],
                         ids=["single_pass_assembly", "dense_level0_labelling"])
def test_rewritten_kernels_are_bit_identical_to_the_ones_they_replace(default_digests, variant):
    other = _run(variant)
    assert other.keys() == default_digests.keys()
    for name in default_digests:
        assert other[name] == default_digests[name], (name, {k: (other[name][k], default_digests[name][k]) for k in other[name]
                                                             if other[name][k] != default_digests[name][k]})


def test_pinned_face_weights_are_read_in_place_and_give_the_same_system():
    """Face weights are read once per level-0 row (k_gather_face_weights).  Pageable host arrays are uploaded whole; pinned ones
    are not copied at all -- the kernel reads them through the mapped host pointer.  Same matrix, rhs and solution, bit for bit;
    the scene has a solid ground, so the face weights are not all 1."""
    import numpy as np
    import torch

    sys.path.insert(0, str(ROOT))
    import bench
    from adaptiveviscositysolver_b200 import scenes
    from adaptiveviscositysolver_b200.solver import Params, Solver

    sc = scenes.sphere_drop(64, 20, center=(0.5, 0.4, 0.5), ground_height=0.125, noise=0.01)
    rng = np.random.default_rng(3)
    for f in sc.face_weights:                      # partially open faces, as Houdini's collision weights have them near solids
        f.data[...] = np.where(rng.random(f.data.shape) < 0.2, rng.random(f.data.shape), 1.0).astype(np.float32)
    p = Params(octree_levels=4, tolerance=1e-9)
    s = Solver(device=0)
    out_a = [v.data.copy() for v in sc.vel]
    info_a = s.solve(sc, p, out_a)                 # pageable numpy arrays: bulk upload
    sys_a, x_a = s.system(), s.solution()
    psc = bench.to_pinned_scene(sc, torch)
    out_b = [torch.from_numpy(v.data.copy()).pin_memory() for v in sc.vel]
    info_b = s.solve(psc, p, out_b)                # pinned: mapped host pointers
    sys_b, x_b = s.system(), s.solution()
    assert info_a.iterations == info_b.iterations and info_a.iterations > 10
    for a, b in zip(sys_a, sys_b):
        assert np.array_equal(a, b)
    assert np.array_equal(x_a, x_b)
    for a in range(3):
        assert np.array_equal(out_a[a], out_b[a].numpy())
    s.close()
