"""CPU statement of the invariant behind the tile-culled labelling of csrc/avs_labels.cu (k_tile_flags / k_tile_list /
k_classify_*_tiles): the label grids are memset to UNASSIGNED and only the 16^3 tiles selected by two flags of the cell tile
(any ACTIVE cell / any cell that is not UP) and the reference's occupancy maps are classified.  That is only correct if EVERY
entry the reference labels with anything but UNASSIGNED lies in a selected tile -- checked here on the oracle's labels
(the oracle is pinned to the reference's code, tests/test_reference_pin.py) for faces, edges and centres of every level.
The GPU counterpart is tests/test_gpu_kernel_variants.py (tile-culled vs dense sweeps, bit for bit)."""
import numpy as np
import pytest

from adaptiveviscositysolver_b200 import scenes
from oracle import avs_oracle as orc

TILE = 16


def _tiles_of(mask):
    """set of tile coordinates (tz, ty, tx) that contain a True entry of the (z, y, x) mask"""
    z, y, x = np.nonzero(mask)
    return set(zip((z // TILE).tolist(), (y // TILE).tolist(), (x // TILE).tolist()))


def _check(sc, params):
    O = orc.OracleRun(sc, params, stop_after_stage=5)      # stages 1-5: weights, octree, labels
    L = O.levels
    dx = float(np.float32(sc.dx))
    for l in range(L):
        lab = O.labels(l)                                   # (z, y, x) uint8
        active, non_up = _tiles_of(lab == orc.ACTIVE), _tiles_of(lab != orc.UP)
        ncell_tiles = tuple((n + TILE - 1) // TILE for n in lab.shape)
        for a in range(3):
            ax = 2 - a                                      # numpy axis of grid axis a
            # ---- faces
            f = O.face_index(l, a)
            zz, yy, xx = np.nonzero(f != orc.UNASSIGNED)
            if l == 0:
                # findOccupiedRegularVelocityTiles (AV.cpp:886-943): faces of every cell with sdf < 2 dx (both faces along `a`)
                near = sc.surface.data.astype(np.float64) < 2.0 * dx
                occ = np.zeros(f.shape, bool)
                sl_lo = [slice(0, s) for s in near.shape]
                occ[tuple(sl_lo)] |= near
                sl_hi = [slice(0, s) for s in near.shape]
                sl_hi[ax] = slice(1, near.shape[ax] + 1)
                occ[tuple(sl_hi)] |= near
                occupied = _tiles_of(occ)
            for z, y, x in zip(zz.tolist(), yy.tolist(), xx.tolist()):
                t = (z // TILE, y // TILE, x // TILE)
                lower = list(t); lower[ax] -= 1; lower = tuple(lower)
                if l == 0:
                    in_grid = all(t[k] < ncell_tiles[k] for k in range(3))
                    selected = t in occupied and (not in_grid or t[ax] == 0 or t in non_up or lower in non_up)
                else:
                    selected = t in active or lower in active
                assert selected, f"level {l} axis {a}: face {(x, y, z)} = {f[z, y, x]} lies in an unselected tile"
            # ---- edges: occupied tiles = tiles of the 4 a-edges of every ACTIVE cell (AV.cpp:1002-1057)
            e = O.edge_index(l, a)
            act = lab == orc.ACTIVE
            eocc = np.zeros(e.shape, bool)
            t1, t2 = [k for k in range(3) if k != ax]       # the two numpy axes across the edge direction
            for d1 in (0, 1):
                for d2 in (0, 1):
                    sl = [slice(0, s) for s in act.shape]
                    sl[t1] = slice(d1, act.shape[t1] + d1)
                    sl[t2] = slice(d2, act.shape[t2] + d2)
                    eocc[tuple(sl)] |= act
            etiles = _tiles_of(eocc)
            zz, yy, xx = np.nonzero(e != orc.UNASSIGNED)
            for z, y, x in zip(zz.tolist(), yy.tolist(), xx.tolist()):
                assert (z // TILE, y // TILE, x // TILE) in etiles, f"level {l} axis {a}: edge {(x, y, z)} outside the occupied edge tiles"
        # ---- centres
        c = O.center_index(l)
        zz, yy, xx = np.nonzero(c != orc.UNASSIGNED)
        for z, y, x in zip(zz.tolist(), yy.tolist(), xx.tolist()):
            assert (z // TILE, y // TILE, x // TILE) in active
    return L


@pytest.mark.parametrize("name", ["sphere64_l5", "padded_varmu", "solid_ground", "buckling"])
def test_every_labelled_sample_lies_in_a_selected_tile(name):
    if name == "sphere64_l5":
        sc, p = scenes.sphere_drop(64, 26, noise=0.01), orc.OracleParams(octree_levels=5)
    elif name == "padded_varmu":
        sc = scenes.sphere_drop(64, 14, res=(48, 64, 40), center=(0.375, 0.5, 0.3125), variable_viscosity=True, variable_density=True)
        p = orc.OracleParams(octree_levels=5)
    elif name == "solid_ground":
        sc = scenes.sphere_drop(32, 9, center=(0.5, 0.34, 0.5), ground_height=0.125, ground_velocity=(0.1, 0.0, -0.2))
        p = orc.OracleParams(octree_levels=3)
    else:
        sc, p = scenes.buckling_sheet(frame=6, dx=0.002), orc.OracleParams(octree_levels=4, dt=1.0 / 120.0)
    assert _check(sc, p) >= 1
