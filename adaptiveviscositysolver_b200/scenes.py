"""Caller-side field containers and synthetic scenes.

In the reference these fields are owned by Houdini (``SIM_ScalarField`` / ``SIM_VectorField``
fetched at HDK_AdaptiveViscosity.cpp:138-231).  Here a ``Scene`` is the flat, host-side image of
exactly those seven inputs, in the layout the C-ABI takes (include/avs.h): float32, x-fastest,
numpy shape ``(nz, ny, nx)``.  The generators below produce the analytic test scenes of
SURVEY.md section 8(d); they live on the *caller* side of the boundary -- the solver library never
sees anything but the field arrays.

Sample conventions: a field component records ``org`` = world position of its
sample (0,0,0) and its spacing ``dx``; centre samples sit at (i+1/2) dx, FACE_a samples at integer
i on axis a and +1/2 on the others.
"""
from __future__ import annotations

from dataclasses import dataclass, field as _dc_field
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class SampledField:
    """One scalar component: dense float32 ``data[z, y, x]`` or a constant (``data is None``)."""
    data: Optional[np.ndarray]
    org: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    dx: float = 1.0
    constant: float = 0.0

    @staticmethod
    def const(value: float) -> "SampledField":
        return SampledField(None, (0.0, 0.0, 0.0), 1.0, float(value))

    @property
    def res(self) -> Tuple[int, int, int]:
        if self.data is None:
            return (1, 1, 1)
        nz, ny, nx = self.data.shape
        return (nx, ny, nz)


@dataclass
class Scene:
    res: Tuple[int, int, int]                 # liquid surface resolution (nx, ny, nz)
    origin: Tuple[float, float, float]
    dx: float
    surface: SampledField                     # "surface": liquid SDF, negative inside (AV.cpp:138)
    vel: List[SampledField]                   # "vel": face sampled, read + written (AV.cpp:139)
    face_weights: List[SampledField]          # "surfaceweights" (AV.cpp:144)
    viscosity: SampledField                   # "viscosity" (AV.cpp:203)
    density: SampledField                     # "massdensity" (AV.cpp:218)
    collision: SampledField                   # "collision": positive inside the solid (AV.cpp:141)
    collision_vel: List[SampledField]         # "collisionvel" (AV.cpp:142)
    meta: dict = _dc_field(default_factory=dict)


def center_org(origin, dx):
    return tuple(o + 0.5 * dx for o in origin)


def face_org(origin, dx, axis):
    return tuple(o + (0.0 if a == axis else 0.5 * dx) for a, o in enumerate(origin))


def face_res(res, axis):
    r = list(res)
    r[axis] += 1
    return tuple(r)


def _coords(res, org, dx):
    """1-D world coordinates of the samples along x, y, z (float64)."""
    return [org[a] + dx * np.arange(res[a], dtype=np.float64) for a in range(3)]


def _sphere_sdf(x, y, z, c, radius):
    return np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - radius


def _supersampled_fraction(sdf_fn, res, org, dx, n):
    """Fraction of the n^3 sub-samples (offsets ((k+1/2)/n - 1/2) dx) with sdf < 0, per sample.

    Only samples whose centre value is within sqrt(3)/2 dx of the zero set are supersampled; the
    rest are 0 or 1 exactly (valid because the analytic SDFs used here are 1-Lipschitz).
    """
    xs, ys, zs = _coords(res, org, dx)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij", sparse=True)
    phi = sdf_fn(X, Y, Z)
    w = (phi < 0).astype(np.float32)
    band = np.abs(phi) < 0.87 * dx
    kz, ky, kx = np.nonzero(band)
    if kz.size:
        px, py, pz = xs[kx], ys[ky], zs[kz]
        cnt = np.zeros(kz.size, np.int32)
        offs = ((np.arange(n) + 0.5) / n - 0.5) * dx
        for oz in offs:
            for oy in offs:
                for ox in offs:
                    cnt += (sdf_fn(px + ox, py + oy, pz + oz) < 0)
        w[kz, ky, kx] = (cnt.astype(np.float64) / float(n ** 3)).astype(np.float32)
    return w


def analytic_velocity(x, y, z, c, U=1.0, omega_z=2.0):
    """u = U (sin2pi x cos2pi y cos2pi z, -cos2pi x sin2pi y cos2pi z, 0) + Omega x (x - c)  (SURVEY 8d)."""
    tp = 2.0 * np.pi
    ux = U * np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z) - omega_z * (y - c[1])
    uy = -U * np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z) + omega_z * (x - c[0])
    uz = 0.0 * (x + y + z)
    return ux, uy, uz


def sphere_drop(n: int, radius_cells: float, *, res: Optional[Sequence[int]] = None,
                center: Sequence[float] = (0.5, 0.5, 0.5), rho: float = 1000.0, mu: float = 200.0,
                supersamples: int = 3, velocity: str = "analytic", noise: float = 0.0,
                constant_velocity: Sequence[float] = (0.3, -0.2, 0.1), variable_viscosity: bool = False,
                variable_density: bool = False, ground_height: Optional[float] = None,
                ground_velocity: Sequence[float] = (0.0, 0.0, 0.0), seed: int = 1234) -> Scene:
    """Liquid sphere of radius ``radius_cells * dx`` in the unit cube, dx = 1/n.

    ``res`` defaults to (n, n, n); a non-cubic / non-power-of-two ``res`` exercises the reference's
    padding (HDK_OctreeGrid.cpp:18-24).  ``ground_height`` adds a solid half-space y < ground_height
    (collision SDF positive inside the solid) for the solid-boundary rows.
    """
    dx = 1.0 / n
    res = tuple(int(v) for v in (res if res is not None else (n, n, n)))
    origin = (0.0, 0.0, 0.0)
    c = tuple(float(v) for v in center)
    R = radius_cells * dx
    sdf = lambda x, y, z: _sphere_sdf(x, y, z, c, R)

    corg = center_org(origin, dx)
    xs, ys, zs = _coords(res, corg, dx)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij", sparse=True)
    surface = SampledField(sdf(X, Y, Z).astype(np.float32), corg, dx)

    rng = np.random.default_rng(seed)
    vel, fw = [], []
    for a in range(3):
        forg = face_org(origin, dx, a)
        fres = face_res(res, a)
        fx, fy, fz = _coords(fres, forg, dx)
        FZ, FY, FX = np.meshgrid(fz, fy, fx, indexing="ij", sparse=True)
        if velocity == "analytic":
            comp = analytic_velocity(FX, FY, FZ, c)[a]
            comp = np.broadcast_to(comp, (fres[2], fres[1], fres[0])).astype(np.float32)
        elif velocity == "constant":
            comp = np.full((fres[2], fres[1], fres[0]), constant_velocity[a], np.float32)
        elif velocity == "zero":
            comp = np.zeros((fres[2], fres[1], fres[0]), np.float32)
        else:
            raise ValueError(velocity)
        if noise > 0:
            comp = (comp + rng.normal(0.0, noise, comp.shape)).astype(np.float32)
        vel.append(SampledField(np.ascontiguousarray(comp), forg, dx))
        fw.append(SampledField(_supersampled_fraction(sdf, fres, forg, dx, supersamples), forg, dx))

    if variable_viscosity:
        visc = SampledField(np.broadcast_to((mu * (1.0 + 4.0 * Y)), (res[2], res[1], res[0])).astype(np.float32).copy(), corg, dx)
    else:
        visc = SampledField.const(mu)
    if variable_density:
        dens = SampledField(np.broadcast_to((rho * (1.0 + 0.5 * X)), (res[2], res[1], res[0])).astype(np.float32).copy(), corg, dx)
    else:
        dens = SampledField.const(rho)

    if ground_height is None:
        collision = SampledField.const(-1.0)
        cvel = [SampledField.const(0.0) for _ in range(3)]
    else:
        g = np.broadcast_to(ground_height - Y, (res[2], res[1], res[0])).astype(np.float32).copy()
        collision = SampledField(g, corg, dx)
        cvel = [SampledField.const(float(v)) for v in ground_velocity]

    return Scene(res, origin, dx, surface, vel, fw, visc, dens, collision, cvel,
                 meta={"kind": "sphere_drop", "n": n, "radius_cells": radius_cells, "mu": mu, "rho": rho})
