"""The torch (device-side) scene generator against the numpy one: same fields, so the large configurations (1024^3) can be
generated on the GPU without ever existing in host memory.  Runs on the CPU device here."""
import numpy as np
import pytest
import torch

from adaptiveviscositysolver_b200.scenes import sphere_drop
from adaptiveviscositysolver_b200.scenes_torch import sphere_drop_device, to_host_scene


@pytest.mark.parametrize("n,R,slab", [(32, 10, 32), (48, 17.5, 7), (40, 12, 64)])
def test_device_generator_matches_numpy(n, R, slab):
    ref = sphere_drop(n, R)
    dev = to_host_scene(sphere_drop_device(n, R, torch.device("cpu"), slab=slab))
    assert dev.res == ref.res and dev.dx == ref.dx and dev.origin == ref.origin
    assert np.array_equal(dev.surface.data, ref.surface.data)                     # sqrt / subtract only: bit-identical
    assert dev.surface.org == ref.surface.org
    for a in range(3):
        assert dev.vel[a].org == ref.vel[a].org and dev.face_weights[a].org == ref.face_weights[a].org
        assert np.array_equal(dev.face_weights[a].data, ref.face_weights[a].data)  # supersampled fractions: bit-identical
        assert dev.vel[a].data.shape == ref.vel[a].data.shape
        assert np.abs(dev.vel[a].data - ref.vel[a].data).max() <= 2.4e-7           # sin / cos: float32 rounding of a last-bit difference
    assert dev.viscosity.data is None and dev.viscosity.constant == ref.viscosity.constant
    assert dev.density.constant == ref.density.constant and dev.collision.constant == ref.collision.constant


def test_device_scene_drives_the_oracle_like_the_numpy_scene():
    """Same DOF set, and a solution within the float32 input perturbation of the velocity."""
    from oracle import avs_oracle as orc
    a = orc.OracleRun(sphere_drop(32, 10), orc.OracleParams(octree_levels=3, tolerance=1e-10))
    b = orc.OracleRun(to_host_scene(sphere_drop_device(32, 10, torch.device("cpu"))), orc.OracleParams(octree_levels=3, tolerance=1e-10))
    assert a.n_face == b.n_face and a.nnz == b.nnz and np.array_equal(a.face_keys(), b.face_keys())
    assert np.abs(a.solution() - b.solution()).max() < 1e-6
