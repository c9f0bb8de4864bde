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


# ---- generic analytic scene ------------------------------------------------------------------------
def analytic_scene(res: Sequence[int], origin: Sequence[float], dx: float, sdf_fn, vel_fn, *, mu=200.0, rho=1000.0,
                   collision_fn=None, collision_velocity: Sequence[float] = (0.0, 0.0, 0.0), supersamples: int = 3,
                   noise: float = 0.0, seed: int = 1234, meta: Optional[dict] = None) -> Scene:
    """The seven fields of solveGasSubclass (AV.cpp:138-231) sampled from analytic functions.

    ``sdf_fn(x, y, z)`` must be 1-Lipschitz (negative inside the liquid); ``vel_fn(x, y, z) -> (u, v, w)``;
    ``mu`` / ``rho`` are floats (constant field fast path, AV.cpp:2090, 2501) or callables sampled at cell
    centres; ``collision_fn(x, y, z)`` is positive inside the solid (None = no solid).
    """
    res = tuple(int(v) for v in res)
    origin = tuple(float(v) for v in origin)
    corg = center_org(origin, dx)
    xs, ys, zs = _coords(res, corg, dx)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij", sparse=True)
    shape = (res[2], res[1], res[0])

    def center_field(fn):
        return SampledField(np.ascontiguousarray(np.broadcast_to(fn(X, Y, Z), shape).astype(np.float32)), corg, dx)

    surface = center_field(sdf_fn)
    rng = np.random.default_rng(seed)
    vel, fw = [], []
    for a in range(3):
        forg = face_org(origin, dx, a)
        fres = face_res(res, a)
        fx, fy, fz = _coords(fres, forg, dx)
        FZ, FY, FX = np.meshgrid(fz, fy, fx, indexing="ij", sparse=True)
        comp = np.broadcast_to(vel_fn(FX, FY, FZ)[a] + 0.0 * (FX + FY + FZ), (fres[2], fres[1], fres[0])).astype(np.float32)
        if noise > 0:
            comp = (comp + rng.normal(0.0, noise, comp.shape)).astype(np.float32)
        vel.append(SampledField(np.ascontiguousarray(comp), forg, dx))
        fw.append(SampledField(_supersampled_fraction(sdf_fn, fres, forg, dx, supersamples), forg, dx))
    visc = center_field(mu) if callable(mu) else SampledField.const(mu)
    dens = center_field(rho) if callable(rho) else SampledField.const(rho)
    if collision_fn is None:
        collision = SampledField.const(-1.0)
    else:
        collision = center_field(collision_fn)
    cvel = [SampledField.const(float(v)) for v in collision_velocity]
    return Scene(res, origin, dx, surface, vel, fw, visc, dens, collision, cvel, meta=dict(meta or {}))


# ---- C5: buckling sheet (geometry of Scenes/viscousBuckling.hip) -------------------------------------
BUCKLING_FPS = 120.0          # viscousBuckling.hip:3
BUCKLING_GRAVITY = 9.80665


def buckling_sheet(frame: int, *, dx: float = 0.001, first_frame: int = 22, variable_viscosity: bool = True,
                   mu0: float = 200.0, rho: float = 1000.0, supersamples: int = 3, pad_cells: int = 8) -> Scene:
    """Frame ``frame`` (0-based) of the prescribed-geometry buckling-sheet sequence (BASELINE.json configs[4]).

    Geometry from ``Scenes/viscousBuckling.hip``: a 0.1 x 0.1 x 0.01 box centred at (0, 0.25, 0)
    (hip:68358-68375) -- a vertical sheet, thin along z -- poured on the ground plane y = 0 (hip:34106-34120),
    grid spacing dx = 2 x particle separation = 0.001 (hip:33771-33774), 120 fps (hip:3), viscosity 200
    (hip:33993).  Cooking the scene needs Houdini, so the motion is prescribed:

      * the sheet falls freely; its lower edge reaches the ground at t* = sqrt(2 * 0.2 / g) (frame 24.2).  The
        sequence starts at ``first_frame`` so that contact happens inside the 10 frames;
      * after contact the upper part keeps descending at 0.3 x the impact speed and the length fed into the
        ground (``consumed``) is taken up by a sinusoidal fold of the lower end -- a z-displacement of the
        mid-surface, wavelength 0.02, whose amplitude follows from arc-length conservation
        (the folded mid-surface is fold height + consumed long; the slab keeps its normal thickness): the "buckling";
      * velocity = rigid descent above the fold, decelerating linearly to zero at the ground inside it, plus the
        time derivative of the fold displacement; it is sampled on ALL faces (extrapolated FLIP velocity);
      * "variable viscosity" is the synthetic overlay mu(x) = mu0 * (1 + 4 y / 0.25) (exercises AV.cpp:2146-2151,
        2275-2282); density constant; the ground is a solid half space (collision = -y > 0 inside), at rest.

    The SDF is min(max(box planes, fold slab / Lipschitz constant), exact box of the straight upper part):
    1-Lipschitz with the exact zero set, so the supersampling shortcut of ``_supersampled_fraction`` stays exact;
    around the fold the bands of the refinement mask are measured in the compressed distance (wider than nominal).
    """
    g = BUCKLING_GRAVITY
    W, H, T = 0.1, 0.1, 0.01                       # box size (x, y, z)
    y_bottom0 = 0.25 - 0.5 * H
    lam, amp_max = 0.02, 0.014
    k = 2.0 * np.pi / lam
    t_hit = float(np.sqrt(2.0 * y_bottom0 / g))
    v_hit = g * t_hit

    def fold_shape(amp, fold_h, y):
        """mid-surface displacement and its slope: amp * ramp(y) * sin(k y), ramp 1 at the ground, 0 at fold_h"""
        s = np.clip((fold_h - y) / fold_h, 0.0, 1.0)
        ds = np.where((y > 0) & (y < fold_h), -1.0 / fold_h, 0.0)
        return amp * s * np.sin(k * y), amp * (ds * np.sin(k * y) + s * k * np.cos(k * y))

    def state(t):
        """(y_lo, y_hi, descent speed of the upper part, consumed length, fold height, fold amplitude)"""
        if t <= t_hit:
            drop = 0.5 * g * t * t
            return y_bottom0 - drop, y_bottom0 + H - drop, g * t, 0.0, 0.0, 0.0
        consumed = min(0.3 * v_hit * (t - t_hit), 0.8 * H)
        fold_h = min(0.05, 0.01 + consumed)
        ys = np.linspace(0.0, fold_h, 4001)
        lo, hi = 0.0, amp_max                       # arc length of the folded mid-surface = fold_h + consumed
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            arc = np.trapezoid(np.sqrt(1.0 + fold_shape(mid, fold_h, ys)[1] ** 2), ys)
            lo, hi = (mid, hi) if arc < fold_h + consumed else (lo, mid)
        return 0.0, H - consumed, 0.3 * v_hit, consumed, fold_h, 0.5 * (lo + hi)

    t = (first_frame + frame) / BUCKLING_FPS
    y_lo, y_hi, v_top, consumed, fold_h, amp = state(t)
    eps = 1e-4
    damp = (state(t + eps)[5] - state(t - eps)[5]) / (2 * eps)     # d(amp)/dt
    # normal thickness T: z-thickness T * q(y), q = sqrt(1 + slope^2); Lipschitz constant of |z - zmid| - T q / 2 from a
    # fine 1-D sample of |zmid'| + T |q'| / 2 (5 % margin)
    if fold_h > 0:
        ys = np.linspace(0.0, fold_h, 20001)
        slope = fold_shape(amp, fold_h, ys)[1]
        q = np.sqrt(1.0 + slope ** 2)
        gmax = float(np.max(np.abs(slope) + 0.5 * T * np.abs(np.gradient(q, ys))))
        lip = float(np.sqrt(1.0 + (1.05 * gmax) ** 2))
    else:
        lip = 1.0

    def sdf(x, y, z):
        side = np.abs(x) - 0.5 * W
        if fold_h <= 0:
            return np.maximum(np.maximum(side, np.abs(z) - 0.5 * T + 0.0 * y), np.maximum(y_lo - y, y - y_hi))
        zm, sl = fold_shape(amp, fold_h, y)
        slab = (np.abs(z - zm) - 0.5 * T * np.sqrt(1.0 + sl ** 2)) / lip
        whole = np.maximum(np.maximum(side, slab), np.maximum(y_lo - y, y - y_hi))
        # the straight part above the fold as a body of its own: exact distances there, so its interior coarsens
        upper = np.maximum(np.maximum(side, np.abs(z) - 0.5 * T + 0.0 * y), np.maximum(fold_h - y, y - y_hi))
        return np.minimum(whole, upper)

    def vel(x, y, z):
        if fold_h > 0:
            s = np.clip((fold_h - y) / fold_h, 0.0, 1.0)
            return 0.0 * x, -v_top * (1.0 - s) + 0.0 * x, damp * s * np.sin(k * y) + 0.0 * x
        return 0.0 * x, -v_top + 0.0 * (x + y), 0.0 * x

    y_top0 = state(first_frame / BUCKLING_FPS)[1]
    nx = int(round(W / dx)) + 2 * pad_cells
    ny = int(np.ceil(y_top0 / dx)) + 2 * pad_cells
    nz = int(round((T + 2 * amp_max) / dx)) + 2 * pad_cells
    origin = (-0.5 * nx * dx, -pad_cells * dx, -0.5 * nz * dx)
    mu = (lambda x, y, z: mu0 * (1.0 + 4.0 * y / 0.25) + 0.0 * (x + z)) if variable_viscosity else mu0
    return analytic_scene((nx, ny, nz), origin, dx, sdf, vel, mu=mu, rho=rho, collision_fn=lambda x, y, z: -y + 0.0 * (x + z),
                          supersamples=supersamples,
                          meta={"kind": "buckling_sheet", "frame": frame, "time": t, "dx": dx, "consumed": consumed,
                                "fold_amplitude": amp, "dt": 1.0 / BUCKLING_FPS})


def buckling_sequence(frames: int = 10, **kw):
    """The 10 prescribed-geometry frames of C5 (generator: one Scene at a time, they are large at dx = 0.001)."""
    for f in range(frames):
        yield buckling_sheet(f, **kw)
