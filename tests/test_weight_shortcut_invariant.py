"""CPU statement of the invariant behind the light pass of the integration weights (csrc/avs_labels.cu: k_sign_x / k_sign_axis and
their 4-wide forms, k_weights_classify4): 98 % of the samples never reach the super-sampler -- a sample whose clamped cell has a 3x3x3
voxel neighbourhood of one sign gets the weight 1 (all negative) or 0 (all non-negative) directly.  That is only correct if the
REFERENCE's super-sampled weight (computeSDFWeightsSampled, AV.cpp:712-791) is exactly 1 / 0 on every such sample, for centre and
edge samples alike -- checked here on the weights of the compiled reference (oracle/_ref), on smooth, distorted and noisy SDFs.
The GPU counterpart: every GPU parity test compares the weight grids with the reference's bit for bit."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from adaptiveviscositysolver_b200 import scenes
from oracle import avs_oracle as orc
from oracle import avs_ref as ref

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(ref.build() is None, reason="oracle/_ref/libavs_ref.so not built and /root/reference not present to build it")

_spec = importlib.util.spec_from_file_location("fuzz_reference_pin", ROOT / "scripts" / "fuzz_reference_pin.py")
fz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fz)


def sign_classes(sdf):
    """class 0: every voxel of the clamped 3x3x3 neighbourhood < 0, 1: every voxel >= 0, 2: mixed -- separable, x then y then z
    (k_sign_x, k_sign_axis; combine3)"""
    c = np.where(sdf < 0, 0, 1).astype(np.uint8)
    for ax in (2, 1, 0):            # numpy axes of x, y, z
        lo = np.concatenate([np.take(c, [0], axis=ax), np.take(c, range(c.shape[ax] - 1), axis=ax)], axis=ax)
        hi = np.concatenate([np.take(c, range(1, c.shape[ax]), axis=ax), np.take(c, [c.shape[ax] - 1], axis=ax)], axis=ax)
        c = np.where((lo == c) & (c == hi), c, 2).astype(np.uint8)
    return c


def _check(sc, p):
    assert not p.do_apply_solid_weights           # the division by the collision weights comes after the pass checked here
    p.max_iterations = 1                          # only stage 1 matters here
    R = ref.RefRun(sc, p)
    cls = sign_classes(sc.surface.data)
    nz, ny, nx = cls.shape
    decided = 0
    for w in [R.center_weights()] + [R.edge_weights(a) for a in range(3)]:
        z, y, x = np.meshgrid(*[np.minimum(np.arange(n), m - 1) for n, m in zip(w.shape, (nz, ny, nx))], indexing="ij", sparse=True)
        c = cls[z, y, x]                         # class of the sample's clamped cell (k_weights_classify4)
        assert np.all(w[c == 0] == 1.0), "a sample the light pass sets to 1 is not 1 in the reference"
        assert np.all(w[c == 1] == 0.0), "a sample the light pass sets to 0 is not 0 in the reference"
        decided += int((c != 2).sum())
    return decided / sum(w.size for w in [R.center_weights()] + [R.edge_weights(a) for a in range(3)])


def test_light_pass_premise_on_sphere_drops():
    for sc, p in ((scenes.sphere_drop(32, 11, noise=0.01), orc.OracleParams(octree_levels=4)),
                  (scenes.sphere_drop(64, 14, res=(48, 64, 40), center=(0.375, 0.5, 0.3125)), orc.OracleParams(octree_levels=5, number_super_samples=2)),
                  (scenes.buckling_sheet(frame=6, dx=0.002), orc.OracleParams(octree_levels=4))):
        assert _check(sc, p) > 0.5               # the pass decides most samples


@pytest.mark.parametrize("seed", [2, 3, 5, 6, 14, 43, 192, 2001, 2005, 2010, 2017, 2030])
def test_light_pass_premise_on_random_scenes(seed):
    """incl. distorted, noisy SDFs (seeds 2000+): the premise is convexity of the trilinear interpolant, not the SDF's Lipschitz bound"""
    sc, p, _ = fz.fuzz_case(seed)
    p.do_apply_solid_weights = False
    _check(sc, p)
