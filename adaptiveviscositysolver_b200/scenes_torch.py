"""Sphere-drop scene generated with torch ops on a chosen device (caller side, like ``scenes.py``).

The numpy generator of ``scenes.sphere_drop`` needs ~30 GB of host arrays (and minutes of CPU time) for the 1024^3 /
40 M-DOF configuration (BASELINE.json configs[3]) -- per rank.  This module evaluates the same analytic fields directly
on the GPU, slab by slab, so the large configurations never exist in host memory ("Memory at large
N").  Same conventions, same formulas, same float64 -> float32 rounding as ``scenes.py``: the surface SDF and the
supersampled face weights are bit-identical to the numpy version (sqrt / compare only); the velocity goes through
sin / cos, whose last bit may differ between libm and CUDA, i.e. it agrees to float32 rounding (tests/test_scenes_torch.py).

PyTorch is plumbing here (device memory + elementwise ops); the solver never sees anything but the field pointers.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from .scenes import SampledField, Scene, center_org, face_org, face_res


def _axis_coords(torch, res, org, dx, device):
    return [org[a] + dx * torch.arange(res[a], dtype=torch.float64, device=device) for a in range(3)]


def _sphere(torch, X, Y, Z, c, R):
    return torch.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - R


def _fill_by_slabs(torch, out, res, org, dx, device, fn, slab: int):
    """out[z0:z1] = fn(X, Y, Z) for z-slabs of at most ``slab`` planes (bounds the float64 temporaries)."""
    xs, ys, zs = _axis_coords(torch, res, org, dx, device)
    X = xs.view(1, 1, -1)
    Y = ys.view(1, -1, 1)
    for z0 in range(0, res[2], slab):
        z1 = min(z0 + slab, res[2])
        Z = zs[z0:z1].view(-1, 1, 1)
        out[z0:z1] = fn(X, Y, Z, z0, z1)


def sphere_drop_device(n: int, radius_cells: float, device, *, center: Sequence[float] = (0.5, 0.5, 0.5), rho: float = 1000.0,
                       mu: float = 200.0, supersamples: int = 3, U: float = 1.0, omega_z: float = 2.0, slab: int = 32) -> Scene:
    """``scenes.sphere_drop(n, radius_cells)`` (analytic velocity, no noise, no solid, constant mu / rho) with every dense
    field as a float32 torch tensor of shape (nz, ny, nx) on ``device``."""
    import torch

    dx = 1.0 / n
    res = (n, n, n)
    origin = (0.0, 0.0, 0.0)
    c = tuple(float(v) for v in center)
    R = radius_cells * dx
    tp = 2.0 * math.pi

    corg = center_org(origin, dx)
    surface = torch.empty((n, n, n), dtype=torch.float32, device=device)
    _fill_by_slabs(torch, surface, res, corg, dx, device, lambda X, Y, Z, z0, z1: _sphere(torch, X, Y, Z, c, R).to(torch.float32), slab)

    offs = [((k + 0.5) / supersamples - 0.5) * dx for k in range(supersamples)]
    vel, fw = [], []
    for a in range(3):
        forg = face_org(origin, dx, a)
        fres = face_res(res, a)
        shape = (fres[2], fres[1], fres[0])
        v = torch.empty(shape, dtype=torch.float32, device=device)
        w = torch.empty(shape, dtype=torch.float32, device=device)

        def velocity(X, Y, Z, z0, z1, a=a):
            if a == 0:
                comp = U * torch.sin(tp * X) * torch.cos(tp * Y) * torch.cos(tp * Z) - omega_z * (Y - c[1])
            elif a == 1:
                comp = -U * torch.cos(tp * X) * torch.sin(tp * Y) * torch.cos(tp * Z) + omega_z * (X - c[0])
            else:
                comp = 0.0 * (X + Y + Z)
            return comp.expand(z1 - z0, shape[1], shape[2]).to(torch.float32)

        def weights(X, Y, Z, z0, z1):
            # fraction of the n^3 sub-samples with sdf < 0; only samples within sqrt(3)/2 dx of the zero set are supersampled
            # (exact: the sphere SDF is 1-Lipschitz), everything else is 0 or 1 -- the rule of scenes._supersampled_fraction
            phi = _sphere(torch, X, Y, Z, c, R)
            out = (phi < 0).to(torch.float32)
            band = (phi.abs() < 0.87 * dx).nonzero(as_tuple=True)
            if band[0].numel():
                px = X.reshape(-1)[band[2]]
                py = Y.reshape(-1)[band[1]]
                pz = Z.reshape(-1)[band[0]]
                cnt = torch.zeros(px.shape, dtype=torch.int32, device=device)
                for oz in offs:
                    for oy in offs:
                        for ox in offs:
                            cnt += (_sphere(torch, px + ox, py + oy, pz + oz, c, R) < 0).to(torch.int32)
                out[band] = (cnt.to(torch.float64) / float(supersamples ** 3)).to(torch.float32)
            return out

        _fill_by_slabs(torch, v, fres, forg, dx, device, velocity, slab)
        _fill_by_slabs(torch, w, fres, forg, dx, device, weights, slab)
        vel.append(SampledField(v, forg, dx))
        fw.append(SampledField(w, forg, dx))

    return Scene(res, origin, dx, SampledField(surface, corg, dx), vel, fw, SampledField.const(mu), SampledField.const(rho),
                 SampledField.const(-1.0), [SampledField.const(0.0) for _ in range(3)],
                 meta={"kind": "sphere_drop", "n": n, "radius_cells": radius_cells, "mu": mu, "rho": rho, "generator": "torch"})


def to_host_scene(scene: Scene) -> Scene:
    """Copies every dense tensor field to a numpy array (tests / the e2e leg of small configurations)."""
    def mv(f: SampledField) -> SampledField:
        if f.data is None or isinstance(f.data, np.ndarray):
            return f
        return SampledField(f.data.detach().cpu().numpy(), f.org, f.dx, f.constant)

    return Scene(scene.res, scene.origin, scene.dx, mv(scene.surface), [mv(v) for v in scene.vel], [mv(v) for v in scene.face_weights],
                 mv(scene.viscosity), mv(scene.density), mv(scene.collision), [mv(v) for v in scene.collision_vel], meta=dict(scene.meta))
