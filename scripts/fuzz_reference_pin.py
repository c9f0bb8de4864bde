"""Random-scene fuzzing of the restated oracle (oracle/avs_oracle.cpp) against the reference's own code (oracle/_ref/libavs_ref.so =
/root/reference/Source/*.cpp compiled unchanged against the stand-ins of oracle/mock_hdk).  TEST INFRASTRUCTURE: CPU only.

    python scripts/fuzz_reference_pin.py [first_seed] [count]

Every seed draws a scene (union of 1-3 liquid blobs -- spheres and boxes -- on a random, usually non-cubic grid with a random origin
and voxel size; optionally a solid: tilted plane or sphere, moving; optionally variable viscosity / density; optional velocity noise)
and a set of DOP options (octree levels 1-6, enhanced gradients, solid weights, band width, super-samples, extrapolation, dt,
tolerance) and holds the oracle to the reference exactly as tests/test_reference_pin.py does: labels, weights, numbering, sparsity,
matrix values, rhs and restricted velocity bit for bit, iteration counts exactly, solution and output velocity to 1e-9.
`fuzz_case(seed)` is what tests/test_reference_fuzz.py replays for a fixed list of seeds.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from adaptiveviscositysolver_b200 import scenes  # noqa: E402
from oracle import avs_oracle as orc  # noqa: E402


OUT_OF_CONTRACT_SEED = 23     # liquid reaching the grid boundary: the reference's debug checks fail, its release build emits column -3


def _sphere(c, r):
    return lambda x, y, z: np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r


def _box(c, h):
    def f(x, y, z):
        qx, qy, qz = np.abs(x - c[0]) - h[0], np.abs(y - c[1]) - h[1], np.abs(z - c[2]) - h[2]
        outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
        return outside + np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0.0)
    return f


def fuzz_case(seed: int):
    """(scene, OracleParams, description) of one seed -- deterministic."""
    rng = np.random.default_rng(100000 + seed)
    variant = seed // 1000       # 0: grids of 16-48 cells; 1: 56-96 cells (deeper octrees); 2: as 0 with a distorted, noisy SDF;
    #                              3: 96-144 cells, one large blob first (4-5 levels built; tens of seconds per seed);
    #                              4: as 0, always with a solid whose velocity is a sampled field on a grid of its own
    sizes = {1: [56, 64, 72, 80, 96], 3: [96, 112, 128, 144]}.get(variant, [16, 20, 24, 28, 32, 36, 40, 48])
    res = tuple(int(v) for v in rng.choice(sizes, size=3))
    if rng.random() < 0.3:
        res = (res[0],) * 3
    dx = float(rng.choice([1.0 / 32, 0.013, 0.1, 0.0625, 0.37]))
    origin = tuple(float(v) for v in (rng.uniform(-1, 1, 3) * dx * rng.choice([0.0, 3.0, 17.5])))
    ext = np.array(res) * dx
    lo = np.array(origin)

    # liquid blobs; unless `touch`, they keep `margin` cells away from the grid boundary (a liquid that reaches the boundary of the
    # grid trips the reference's own debug checks -- edgeStressUnitTest / centerStresUnitTest / octreeLabels.unitTest -- and can put
    # the column -3 into its matrix: see run_seed)
    touch = rng.random() < 0.1
    margin = 0.0 if touch else 4.5 * dx
    blobs = []
    if variant == 3:             # a body thick enough for coarse cells: centred sphere filling most of the smallest extent
        blobs.append(_sphere(lo + 0.5 * ext, float(rng.uniform(0.36, 0.44) * ext.min())))
    for _ in range(int(rng.integers(1, 4))):
        c = lo + ext * rng.uniform(0.25, 0.75, 3)
        room = float(np.min(np.minimum(c - lo, lo + ext - c))) - margin
        if rng.random() < 0.6:
            blobs.append(_sphere(c, min(float(rng.uniform(0.12, 0.42) * ext.min()), max(room, 1.5 * dx))))
        else:
            blobs.append(_box(c, np.minimum(rng.uniform(0.08, 0.35, 3) * ext, np.maximum(np.minimum(c - lo, lo + ext - c) - margin, 1.5 * dx))))
    if touch:                    # liquid reaching the grid boundary
        blobs.append(_sphere(lo + ext * np.array([0.5, 0.0, 0.5]), float(0.3 * ext.min())))

    def sdf_exact(x, y, z):
        v = blobs[0](x, y, z)
        for b in blobs[1:]:
            v = np.minimum(v, b(x, y, z))
        return v

    sdf = sdf_exact
    if variant == 2:             # what a simulation hands over is not an exact distance: scaled by a smooth factor in [0.6, 1.4]
        kd = 2.0 * np.pi / (float(ext.max()) * float(rng.uniform(0.2, 1.0)))
        pd = rng.uniform(0, 6.28, 3)
        sdf = lambda x, y, z: sdf_exact(x, y, z) * (1.0 + 0.4 * np.sin(kd * x + pd[0]) * np.cos(kd * y + pd[1]) * np.sin(kd * z + pd[2]))

    U = float(rng.uniform(0.1, 2.0))
    k = 2.0 * np.pi / float(ext.max())
    ph = rng.uniform(0, 6.28, 3)

    def vel(x, y, z):
        return (U * np.sin(k * x + ph[0]) * np.cos(k * y) + 0.0 * z, -U * np.cos(k * x) * np.sin(k * y + ph[1]) + 0.3 * U * z / ext[2],
                0.5 * U * np.sin(k * z + ph[2]) + 0.0 * (x + y))

    solid = rng.random() * (0.55 if variant == 4 else 1.0)     # variant 4: always a solid (its velocity becomes a sampled field below)
    collision_fn, cvel = None, (0.0, 0.0, 0.0)
    if solid < 0.35:             # tilted ground plane
        n = np.array([rng.uniform(-0.3, 0.3), 1.0, rng.uniform(-0.3, 0.3)])
        n /= np.linalg.norm(n)
        h = float(lo[1] + ext[1] * rng.uniform(0.15, 0.45))
        collision_fn = lambda x, y, z: -(n[0] * (x - lo[0] - 0.5 * ext[0]) + n[1] * (y - h) + n[2] * (z - lo[2] - 0.5 * ext[2]))
        cvel = tuple(float(v) for v in rng.uniform(-0.5, 0.5, 3))
    elif solid < 0.55:           # solid sphere poking into the liquid
        c = lo + ext * rng.uniform(0.3, 0.7, 3)
        r = float(rng.uniform(0.1, 0.25) * ext.min())
        collision_fn = lambda x, y, z: r - np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2)
        cvel = tuple(float(v) for v in rng.uniform(-0.5, 0.5, 3))

    mu0 = float(rng.choice([0.5, 20.0, 200.0, 5000.0]))
    mu = mu0
    if rng.random() < 0.4:
        mu = lambda x, y, z: mu0 * (1.0 + 3.0 * (y - lo[1]) / ext[1] + 0.5 * np.sin(k * x)) + 0.0 * z
    rho0 = float(rng.choice([1.0, 1000.0]))
    rho = rho0
    if rng.random() < 0.3:
        rho = lambda x, y, z: rho0 * (1.0 + 0.5 * (x - lo[0]) / ext[0]) + 0.0 * (y + z)

    ss = int(rng.choice([1, 2, 3, 4]))
    sc = scenes.analytic_scene(res, origin, dx, sdf, vel, mu=mu, rho=rho, collision_fn=collision_fn, collision_velocity=cvel,
                               supersamples=ss, noise=float(rng.choice([0.0, 0.0, 0.01])), seed=seed)
    if variant == 2:             # + noise of a few percent of a voxel on the stored samples (zero set moves, band edges get ragged)
        sc.surface.data += (rng.normal(0.0, 0.03 * dx, sc.surface.data.shape)).astype(np.float32)
    if collision_fn is not None and rng.random() < 0.4:
        # the reference samples the collision field by world position (AV.cpp:141, 853, 1157): give it a grid of its own
        cdx = dx * float(rng.choice([0.75, 1.5, 2.0]))
        cres = tuple(int(np.ceil(e / cdx)) + 2 for e in ext)
        corg = tuple(float(o - 0.7 * cdx) for o in origin)
        xs, ys, zs = [corg[a] + cdx * np.arange(cres[a]) for a in range(3)]
        Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij", sparse=True)
        sc.collision = scenes.SampledField(np.ascontiguousarray(np.broadcast_to(collision_fn(X, Y, Z), (cres[2], cres[1], cres[0])).astype(np.float32)), corg, cdx)
    if variant == 4:
        # the solid's velocity as three dense fields on a grid of their own (the reference samples "collisionvel" by world position:
        # boundary terms of the stencils AV.cpp:1896-1905, 1952-1961, and the solid faces of the write-back AV.cpp:2860-2890):
        # rigid rotation about a random axis through the domain centre plus a translation
        w = rng.uniform(-2.0, 2.0, 3)
        t0 = rng.uniform(-0.5, 0.5, 3)
        cc = lo + 0.5 * ext
        vdx = dx * float(rng.choice([1.0, 1.7, 2.5]))
        vres = tuple(int(np.ceil(e / vdx)) + 3 for e in ext)
        vorg = tuple(float(o - 1.2 * vdx) for o in origin)
        xs, ys, zs = [vorg[a] + vdx * np.arange(vres[a]) for a in range(3)]
        Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij", sparse=True)
        rx, ry, rz = X - cc[0], Y - cc[1], Z - cc[2]
        comps = (t0[0] + w[1] * rz - w[2] * ry, t0[1] + w[2] * rx - w[0] * rz, t0[2] + w[0] * ry - w[1] * rx)
        sc.collision_vel = [scenes.SampledField(np.ascontiguousarray(np.broadcast_to(c, (vres[2], vres[1], vres[0])).astype(np.float32)), vorg, vdx)
                            for c in comps]
    p = orc.OracleParams(octree_levels=int(rng.integers(4, 8)) if variant == 3 else int(rng.integers(1, 7)), tolerance=float(rng.choice([1e-3, 1e-6, 1e-10])),
                         dt=float(rng.choice([1.0 / 24, 1.0 / 120, 0.5])), use_enhanced_gradients=bool(rng.random() < 0.75),
                         do_apply_solid_weights=bool(collision_fn is not None and rng.random() < 0.4),
                         fine_bandwidth=int(rng.choice([1, 2, 2, 3, 4])), number_super_samples=ss,
                         extrapolation=float(rng.choice([0.5, 0.0, 1.0])), max_iterations=int(rng.choice([2500, 2500, 40])))
    if variant == 3:
        p.max_iterations = 40        # half a million rows: the reference's serial CG is not the subject of this variant
    desc = (f"seed {seed}{['', ' (deep)', ' (distorted sdf)', ' (large)', ' (sampled solid velocity)'][min(variant, 4)]}: res {res} dx {dx:g} origin {tuple(round(o, 4) for o in origin)} blobs {len(blobs)}{' TOUCHING-THE-BOUNDARY' if touch else ''} solid "
            f"{'plane' if solid < 0.35 else 'sphere' if solid < 0.55 else 'none'} own-collision-grid {sc.collision.data is not None and sc.collision.dx != dx} "
            f"mu {'var' if callable(mu) else mu0} rho {'var' if callable(rho) else rho0} | levels {p.octree_levels} tol {p.tolerance:g} dt {p.dt:g} "
            f"enh {p.use_enhanced_gradients} solidw {p.do_apply_solid_weights} band {p.fine_bandwidth} ss {ss} extrap {p.extrapolation} maxit {p.max_iterations}")
    return sc, p, desc


def compare(sc, p):
    """Runs both; raises AssertionError on any difference.  Returns (levels, octree DOFs, iterations); a negative third entry is
    minus the number of out-of-range columns of a scene outside the reference's contract."""
    from oracle import avs_ref as ref
    sys.path.insert(0, str(ROOT / "tests"))
    import test_reference_pin as trp
    # doPrintOctree / onlyPrintOctree (AV.cpp:283-294, OG.cpp:245-308): the same points, pscale and octreeLevel
    Rg, Og = ref.RefRun(sc, p, octree_only=True), orc.OracleRun(sc, p, stop_after_stage=3)
    assert sorted(map(tuple, np.column_stack(Rg.octree_points()).tolist())) == sorted(map(tuple, np.column_stack(Og.octree_points()).tolist())), "octree geometry"
    R, O = ref.RefRun(sc, p), orc.OracleRun(sc, p)
    if R.n_face == 0 or O.n_face == 0:
        assert R.n_face == O.n_face
        return R.levels, 0, 0
    # A scene can violate the reference's own assumptions: getEdgeStressFaces appends the parent face of an UNASSIGNED face without
    # looking at its label (`assert(parentVelocityIndex >= 0)`, AV.cpp:1886-1894, compiled out in release builds), so a triplet can
    # carry the column -3 (OUTSIDE).  Real Eigen's setFromTriplets would write out of bounds there; the stand-in keeps the entry.
    # Such scenes are compared on the raw CSR arrays (no permutation) and reported.
    rp, rc, rv = R.csr()
    if rc.size and rc.min() < 0:
        op, oc, ov = O.csr()
        assert np.array_equal(R.face_keys(), O.face_keys()), "numbering"
        assert np.array_equal(rp, op) and np.array_equal(rc, oc), "raw CSR structure (scene with out-of-range columns)"
        assert np.array_equal(rv, ov), "raw CSR values (scene with out-of-range columns)"
        return R.levels, R.n_face, -int((rc < 0).sum())
    perm, Ar, Ao = trp.assert_same_run(R, O, sc)
    assert np.array_equal(Ar.data, Ao.data), "matrix values"
    assert np.array_equal(R.rhs(), O.rhs()[perm]), "rhs"
    assert np.array_equal(R.x0(), O.x0()[perm]), "restricted velocity"
    # CG: the two loops are the same recurrence but sum their dot products in different orders, so after hundreds of iterations on
    # an ill-conditioned system (high viscosity on rho = 1 with a long time step: seeds 105, 162, 272, 297, 424 of the first 800) they
    # stop up to 8 % apart.  What must hold exactly is checked above (matrix, rhs, initial guess); here: iteration counts exact
    # below 100 iterations and within 10 % above, both solutions meet the tolerance on the TRUE
    # residual (unless the iteration limit ended the loop), and the regular-grid outputs differ by no more than the solutions do
    # (+ one float32 rounding of the stored value).
    slack = 0 if R.iterations < 100 else R.iterations // 10
    assert abs(R.iterations - O.iterations) <= slack, (R.iterations, O.iterations)
    scale = max(1.0, float(np.abs(O.solution()).max()))
    b = O.rhs()
    bn = float(np.linalg.norm(b))
    for name, run, x in (("reference", R, R.solution()), ("oracle", O, O.solution()[perm])):
        if run.iterations < p.max_iterations and bn > 0:
            # 1e-7: drift of the recursive residual (seed 828 -- viscosity 5000, dt 0.5, tolerance 1e-10: reference and oracle both stop
            # after 193 iterations with the SAME recursive error 8.98e-11 and the SAME true residual 4.7e-8)
            assert np.linalg.norm(b[perm] - Ar @ x) <= (10.0 * p.tolerance + 1e-7) * bn, f"true residual of the {name}'s solution"
    dsol = float(np.abs(R.solution() - O.solution()[perm]).max())
    if R.iterations == O.iterations and R.iterations < 100:      # (seed 749: 90 iterations to 1e-3 on rho = 1 end 3e-6 apart)
        assert dsol < max(1e-9, 0.01 * p.tolerance) * scale, "solution"
    for a in range(3):
        ro, oo = R.out_velocity(a).astype(np.float64), O.out_velocity(a).astype(np.float64)
        assert np.all(np.abs(ro - oo) <= 4.0 * dsol + 1e-9 * scale + 2.0 ** -22 * np.abs(oo)), f"output velocity, axis {a}"
        assert np.array_equal(ro != sc.vel[a].data, oo != sc.vel[a].data) or (ro != sc.vel[a].data).sum() > 0, f"written faces, axis {a}"
    return R.levels, R.n_face, R.iterations


def _worker(mode: str, seed: int) -> int:
    sc, p, _ = fuzz_case(seed)
    if mode == "--debug":      # the reference with its asserts and debug unit tests compiled in: does the scene respect its contract?
        from oracle import avs_ref as ref
        ref._LIB_PATH = ref._HERE / "_ref" / "libavs_ref_debug.so"
        R = ref.RefRun(sc, p)
        print(f"debug-ok {R.levels} {R.n_face} {R.iterations}")
        return 0
    lv, n, it = compare(sc, p)
    print(f"compare-ok levels built {lv} N {n} " + (f"iterations {it}" if it >= 0 else f"OUT-OF-RANGE COLUMNS {-it}"))
    return 0


def run_seed(seed: int, timeout: int = 900):
    """('ok' | 'out-of-contract' | 'FAIL', detail).  Both runs happen in child processes: a scene outside the reference's contract
    can make the reference (or SciPy, fed a negative column) write out of bounds."""
    import subprocess
    me = str(Path(__file__).resolve())
    d = subprocess.run([sys.executable, me, "--debug", str(seed)], capture_output=True, text=True, timeout=timeout)
    c = subprocess.run([sys.executable, me, "--compare", str(seed)], capture_output=True, text=True, timeout=timeout)
    last = lambda r: ([ln for ln in (r.stdout + r.stderr).strip().splitlines() if ln.strip()] or ["(no output)"])[-1][:300]
    if d.returncode != 0:
        return "out-of-contract", f"reference assert: {last(d)} | release reference vs oracle: {last(c) if c.returncode == 0 else 'rc %d %s' % (c.returncode, last(c))}"
    if c.returncode != 0:
        return "FAIL", f"rc {c.returncode}: {last(c)}"
    return "ok", last(c)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] in ("--debug", "--compare"):
        sys.exit(_worker(sys.argv[1], int(sys.argv[2])))
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref", "ref-debug"], check=True, capture_output=True)
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    tally = {"ok": [], "out-of-contract": [], "FAIL": []}
    for s in range(first, first + count):
        verdict, detail = run_seed(s)
        tally[verdict].append(s)
        print(f"{verdict:16s}{fuzz_case(s)[2]}\n                -> {detail}", flush=True)
    print(f"{len(tally['ok'])} of {count} seeds agree, {len(tally['out-of-contract'])} outside the reference's contract {tally['out-of-contract']}, "
          f"failing seeds: {tally['FAIL']}")
    sys.exit(1 if tally["FAIL"] else 0)
