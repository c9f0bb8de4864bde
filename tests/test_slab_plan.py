"""Host-side statement of the z-slab sharding of the regular-grid stages and of the C5 frame dealing
(adaptiveviscositysolver_b200/dist_plan.py mirrors avs_slab_cuts / avs_slab_range of csrc/avs_labels.cu)."""
import numpy as np
import pytest

from adaptiveviscositysolver_b200.dist_plan import deal_frames, slab_cuts, slab_range
from adaptiveviscositysolver_b200.scenes import sphere_drop


@pytest.mark.parametrize("P", [1, 2, 3, 4, 8])
def test_slabs_partition_every_face_grid(P):
    sc = sphere_drop(64, 26)
    liquid = (sc.surface.data < 0).reshape(sc.res[2], -1).sum(axis=1)
    cuts = slab_cuts(liquid, P, sc.res[0] * sc.res[1])
    assert cuts[0] == 0 and cuts[-1] == sc.res[2] and all(b >= a for a, b in zip(cuts, cuts[1:]))
    for axis in range(3):
        planes = sc.res[2] + (1 if axis == 2 else 0)
        covered = np.zeros(planes, int)
        for q in range(P):
            z0, z1 = slab_range(cuts, axis, q, planes)
            covered[z0:z1] += 1
        assert np.all(covered == 1), "every plane of every face grid belongs to exactly one rank"
    # balance: no rank carries more than its share plus one plane of weight
    w = liquid + max(1, sc.res[0] * sc.res[1] // 50)
    loads = [int(w[cuts[q]:cuts[q + 1]].sum()) for q in range(P)]
    assert max(loads) <= w.sum() / P + w.max()


def test_slabs_follow_the_liquid():
    """A drop in the lower part of the domain: the cuts crowd where the liquid is, empty planes stay cheap but not free."""
    sc = sphere_drop(64, 12, center=(0.5, 0.5, 0.25))
    liquid = (sc.surface.data < 0).reshape(sc.res[2], -1).sum(axis=1)
    cuts = slab_cuts(liquid, 4, sc.res[0] * sc.res[1])
    assert cuts[1] < 16 and cuts[2] <= 20 and cuts[3] < 48
    empty = slab_cuts(np.zeros(64, int), 4, 64 * 64)
    assert empty == [0, 16, 32, 48, 64]                      # no liquid at all: equal slabs
    assert slab_cuts(liquid, 1, 64 * 64) == [0, 64]


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_frames_are_dealt_once(P):
    got = sorted(f for r in range(P) for f in deal_frames(10, r, P))
    assert got == list(range(10))
    assert max(len(deal_frames(10, r, P)) for r in range(P)) == -(-10 // P)
