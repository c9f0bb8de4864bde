#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the CPU oracle (oracle/avs_oracle.cpp).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference ships no fixtures and cannot be built here (no Houdini HDK, no Eigen -- DESIGN.md section 6), so
these vectors pin the *oracle* (the restated reference algorithm), not the reference binary: they freeze the
oracle's outputs at the commit that passed the structural and analytic known-answer tests
(tests/test_oracle_*.py), so that a later change to the oracle or to the CUDA path that alters any label,
matrix entry or velocity is caught -- by the CPU suite for the oracle, by the `-m gpu` suite for the CUDA path,
neither of which needs the other to run.

Every case is a seeded analytic scene of adaptiveviscositysolver_b200.scenes.  What is
stored per case (rows sorted by the geometric key (level, axis, k, j, i) because DOF numbering is not part of the
parity contract):

  input_sha256      hashes of the seven input fields, so a silent change of the scene generator is caught first
  counts            levels, octree / edge / centre / regular DOFs, nnz, interpolated faces, CG iterations, error
  label_sha256      per level: cell labels; per level and axis: face / edge classes; centre classes; regular
                    face classes; integration weights (all bit-exact quantities)
  keys              (N, 5) int32
  rhs, x0, x        float64 vectors in key order (x0 = restricted u^n, x = CG solution at tolerance 1e-10)
  diag, Av          matrix diagonal and A @ v for a seeded v (pins every matrix entry through linearity)
  row_nnz           entries per row (pins the sparsity pattern's shape)
  out_sha256/out_*  regular-grid output velocity: hash of the oracle's float32 arrays plus the changed faces
                    (flat index + value) of each axis, capped -- enough to localise a mismatch
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from adaptiveviscositysolver_b200.scenes import buckling_sheet, sphere_drop  # noqa: E402

# name -> scene kwargs, octree levels, extra params
CASES = {
    # BASELINE.json configs[0]: 32^3 uniform (depth 1) sphere drop
    "c1_uniform32": dict(scene=dict(n=32, radius_cells=10), levels=1),
    # "a 64^3 depth-3 case", with seeded velocity noise
    "sphere64_l3_noise": dict(scene=dict(n=64, radius_cells=24, noise=0.01), levels=3),
    # deep octree on a padded non-power-of-two grid, variable viscosity and density
    "padded_48x64x40_l5_varmu": dict(scene=dict(n=64, radius_cells=14, res=(48, 64, 40), center=(0.375, 0.5, 0.3125),
                                                variable_viscosity=True, variable_density=True), levels=5),
    # solid ground plane with a moving solid (SOLIDBOUNDARY faces, boundary stencil terms)
    "solid_ground32_l3": dict(scene=dict(n=32, radius_cells=9, center=(0.5, 0.34, 0.5), ground_height=0.125,
                                         ground_velocity=(0.1, 0.0, -0.2)), levels=3),
    # BASELINE.json configs[4] at reduced resolution: folded buckling-sheet frame, variable viscosity, ground plane, dt = 1/120
    "buckling_f6_dx2mm": dict(maker="buckling_sheet", scene=dict(frame=6, dx=0.002), levels=4, dt=1.0 / 120.0),
}
TOL = 1e-10
MAX_OUT = 200_000


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def key_order(keys: np.ndarray) -> np.ndarray:
    k = keys.astype(np.int64)
    code = (((k[:, 0] * 3 + k[:, 1]) * 4096 + k[:, 4]) * 4096 + k[:, 3]) * 4096 + k[:, 2]
    return np.argsort(code, kind="stable")


def input_hashes(scene) -> dict:
    out = {}
    fields = {"surface": scene.surface, "viscosity": scene.viscosity, "density": scene.density, "collision": scene.collision}
    for a in range(3):
        fields[f"vel{a}"] = scene.vel[a]
        fields[f"face_weights{a}"] = scene.face_weights[a]
        fields[f"collision_vel{a}"] = scene.collision_vel[a]
    for name, f in fields.items():
        out[name] = sha(f.data) if f.data is not None else f"const:{float(f.constant)!r}"
    return out


def classes(grid: np.ndarray) -> np.ndarray:
    """Index grids -> class grids: active entries read 0 (numbering is not part of the contract)."""
    return np.minimum(grid, 0).astype(np.int8)


def golden_vector(n: int) -> np.ndarray:
    """The seeded probe vector of the A @ v pin, in key order."""
    return np.random.default_rng(20261017).standard_normal(n)


def make_scene(case: dict):
    return buckling_sheet(**case["scene"]) if case.get("maker") == "buckling_sheet" else sphere_drop(**case["scene"])


def build(name: str, case: dict) -> dict:
    from oracle import avs_oracle as orc
    sc = make_scene(case)
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=case["levels"], tolerance=TOL, dt=case.get("dt", 1.0 / 24.0)))
    keys = ref.face_keys()
    order = key_order(keys)
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    A = ref.scipy_matrix()
    v = golden_vector(ref.n_face)
    v_oracle = v[inv]                      # probe value of oracle row r = v[position of r in key order]
    labels = {}
    for l in range(ref.levels):
        labels[f"cell{l}"] = sha(ref.labels(l))
        labels[f"center{l}"] = sha(classes(ref.center_index(l)))
        for a in range(3):
            labels[f"face{l}_{a}"] = sha(classes(ref.face_index(l, a)))
            labels[f"edge{l}_{a}"] = sha(classes(ref.edge_index(l, a)))
    labels["center_weights"] = sha(ref.center_weights())
    for a in range(3):
        labels[f"regular{a}"] = sha(classes(ref.regular_index(a)))
        labels[f"edge_weights{a}"] = sha(ref.edge_weights(a))
    out = dict(
        keys=keys[order].astype(np.int32),
        rhs=ref.rhs()[order], x0=ref.x0()[order], x=ref.solution()[order],
        diag=A.diagonal()[order], Av=(A @ v_oracle)[order],
        row_nnz=np.diff(A.indptr)[order].astype(np.int16),
    )
    meta = dict(
        case=case, tolerance=TOL, levels=int(ref.levels), octree_dofs=int(ref.n_face), edge_dofs=int(ref.n_edge),
        center_dofs=int(ref.n_center), regular_dofs=int(ref.regular_dofs), nnz=int(ref.nnz),
        iterations=int(ref.iterations), error=float(ref.error), interpolated_faces=int(ref.interpolated_faces),
        input_sha256=input_hashes(sc), label_sha256=labels, out_sha256={},
    )
    for a in range(3):
        o = ref.out_velocity(a)
        meta["out_sha256"][str(a)] = sha(o)
        changed = np.flatnonzero(o.ravel() != sc.vel[a].data.ravel())
        meta[f"out_changed_{a}"] = int(changed.size)
        if changed.size > MAX_OUT:
            changed = changed[:: int(np.ceil(changed.size / MAX_OUT))]
        out[f"out_idx{a}"] = changed.astype(np.int32)
        out[f"out_val{a}"] = o.ravel()[changed].astype(np.float32)
    out["meta"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8)
    return out


def main():
    for name, case in CASES.items():
        data = build(name, case)
        p = HERE / f"{name}.npz"
        np.savez_compressed(p, **data)
        meta = json.loads(bytes(data["meta"]).decode())
        print(f"{name}: N={meta['octree_dofs']} nnz={meta['nnz']} levels={meta['levels']} iters={meta['iterations']} "
              f"-> {p.name} ({p.stat().st_size / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
