"""Golden-vector tests (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU leg (`-m "not gpu"`): the oracle, re-run on the regenerated scene, still produces the committed vectors.
GPU leg (`-m gpu`): the CUDA path, through the C-ABI, reproduces the same vectors WITHOUT the oracle being
imported -- labels / weights / key sets bit-exact (sha256), matrix / rhs / restricted velocity <= 1e-12 relative,
solved velocity L-inf < 1e-6 (north star), regular-grid output < 1e-6.
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from tests.golden.make_golden import CASES, classes, golden_vector, input_hashes, key_order, make_scene, sha

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.mark.parametrize("name", list(CASES))
def test_fixture_matches_scene_generator(name):
    """The committed vectors belong to the scenes the generator produces today."""
    z, meta = load(name)
    assert meta["case"] == json.loads(json.dumps(CASES[name]))
    sc = make_scene(CASES[name])
    assert input_hashes(sc) == meta["input_sha256"]
    keys = z["keys"]
    assert keys.shape == (meta["octree_dofs"], 5)
    assert np.array_equal(key_order(keys), np.arange(keys.shape[0]))
    assert int(z["row_nnz"].astype(np.int64).sum()) == meta["nnz"]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name):
    from oracle import avs_oracle as orc
    z, meta = load(name)
    case = CASES[name]
    sc = make_scene(case)
    ref = orc.OracleRun(sc, orc.OracleParams(octree_levels=case["levels"], tolerance=meta["tolerance"], dt=case.get("dt", 1.0 / 24.0)))
    assert (ref.levels, ref.n_face, ref.n_edge, ref.n_center, ref.regular_dofs, ref.nnz, ref.interpolated_faces) == \
        (meta["levels"], meta["octree_dofs"], meta["edge_dofs"], meta["center_dofs"], meta["regular_dofs"], meta["nnz"],
         meta["interpolated_faces"])
    L = meta["label_sha256"]
    for l in range(ref.levels):
        assert sha(ref.labels(l)) == L[f"cell{l}"]
        assert sha(classes(ref.center_index(l))) == L[f"center{l}"]
        for a in range(3):
            assert sha(classes(ref.face_index(l, a))) == L[f"face{l}_{a}"]
            assert sha(classes(ref.edge_index(l, a))) == L[f"edge{l}_{a}"]
    assert sha(ref.center_weights()) == L["center_weights"]
    for a in range(3):
        assert sha(classes(ref.regular_index(a))) == L[f"regular{a}"]
        assert sha(ref.edge_weights(a)) == L[f"edge_weights{a}"]
    keys = ref.face_keys()
    order = key_order(keys)
    assert np.array_equal(keys[order], z["keys"])
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    A = ref.scipy_matrix()
    # assembly is deterministic (fixed summation order per row): exact
    assert np.array_equal(ref.rhs()[order], z["rhs"])
    assert np.array_equal(ref.x0()[order], z["x0"])
    assert np.array_equal(A.diagonal()[order], z["diag"])
    assert np.array_equal(np.diff(A.indptr)[order], z["row_nnz"])
    np.testing.assert_allclose((A @ golden_vector(ref.n_face)[inv])[order], z["Av"], rtol=1e-13, atol=1e-9 * np.abs(z["Av"]).max())
    # the CG's OpenMP reductions may re-associate: solution to well below the parity bar
    assert abs(ref.iterations - meta["iterations"]) <= 1
    assert np.abs(ref.solution()[order] - z["x"]).max() < 1e-9
    for a in range(3):
        o = ref.out_velocity(a).ravel()
        assert np.abs(o[z[f"out_idx{a}"]] - z[f"out_val{a}"]).max(initial=0.0) < 1e-7


@pytest.fixture(scope="module")
def solver():
    from adaptiveviscositysolver_b200.solver import Solver
    s = Solver(device=0)
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_reproduces_golden(solver, name):
    from adaptiveviscositysolver_b200.solver import Params
    z, meta = load(name)
    case = CASES[name]
    sc = make_scene(case)
    out = [v.data.copy() for v in sc.vel]
    info = solver.solve(sc, Params(octree_levels=case["levels"], tolerance=meta["tolerance"], dt=case.get("dt", 1.0 / 24.0)), out)
    assert (info.levels, info.octree_dofs, info.edge_dofs, info.center_dofs, info.regular_dofs, info.nnz,
            info.interpolated_faces) == \
        (meta["levels"], meta["octree_dofs"], meta["edge_dofs"], meta["center_dofs"], meta["regular_dofs"], meta["nnz"],
         meta["interpolated_faces"])
    # bit-exact quantities
    L = meta["label_sha256"]
    for l in range(info.levels):
        assert sha(solver.labels(l)) == L[f"cell{l}"], f"cell labels, level {l}"
        assert sha(classes(solver.center_labels(l))) == L[f"center{l}"], f"centre labels, level {l}"
        for a in range(3):
            assert sha(classes(solver.face_labels(l, a))) == L[f"face{l}_{a}"], f"face labels, level {l} axis {a}"
            assert sha(classes(solver.edge_labels(l, a))) == L[f"edge{l}_{a}"], f"edge labels, level {l} axis {a}"
    assert sha(solver.center_weights()) == L["center_weights"]
    for a in range(3):
        assert sha(classes(solver.regular_labels(a))) == L[f"regular{a}"]
        assert sha(solver.edge_weights(a)) == L[f"edge_weights{a}"]
    keys = solver.keys()
    order = key_order(keys)
    assert np.array_equal(keys[order], z["keys"]), "set of velocity DOFs differs from the golden key set"
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    # system: <= 1e-12 relative
    import scipy.sparse as sp
    ptr, col, val, rhs, x0 = solver.system()
    n = info.octree_dofs
    A = sp.csr_matrix((val, col, ptr), shape=(n, n))
    assert np.array_equal(np.diff(ptr)[order], z["row_nnz"])
    scale = np.abs(z["diag"]).max()
    assert np.abs(A.diagonal()[order] - z["diag"]).max() <= 1e-12 * scale
    assert np.abs((A @ golden_vector(n)[inv])[order] - z["Av"]).max() <= 1e-12 * np.abs(z["Av"]).max()
    assert np.abs(rhs[order] - z["rhs"]).max() <= 1e-12 * np.abs(z["rhs"]).max()
    assert np.abs(x0[order] - z["x0"]).max() <= 1e-12 * max(np.abs(z["x0"]).max(), 1e-300)
    # solve: velocity L-inf < 1e-6 (north star); both sides converged to 1e-10
    assert abs(info.iterations - meta["iterations"]) <= max(2, meta["iterations"] // 50)   # stiff buckling frame: +-2 %
    assert np.abs(solver.solution()[order] - z["x"]).max() < 1e-6
    # regular-grid output
    for a in range(3):
        o = out[a].ravel()
        assert np.abs(o[z[f"out_idx{a}"]] - z[f"out_val{a}"]).max(initial=0.0) < 1e-6
        assert int((o != sc.vel[a].data.ravel()).sum()) <= meta[f"out_changed_{a}"] * 1.001 + 8   # untouched faces stay untouched
