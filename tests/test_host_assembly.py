"""The CUDA library's row builder -- the device code of csrc/avs_system.cu: buildRow, buildSimpleRow, edgeStressFaces,
centerStressFaces, control volumes, stress weights, applyToMatrix, both row accumulators, and the leaf / hat-weight functions of the
restriction; of csrc/avs_prolong.cu: the interpolator's node pyramid and interpSPGrid; of csrc/avs_labels.cu: the weight sampler
and the face / edge / centre classification rules -- compiled for the HOST
(tests/host_assembly.cu, host_prolong.cu, host_labels.cu; nvcc -DAVS_HOST_TEST turns the product's AVS_DEV functions into __host__ __device__) and run on label grids
produced by the REFERENCE'S OWN CODE (oracle/_ref/libavs_ref.so).  Its rows are compared with the reference's matrix and right-hand
side, its restricted velocity and its regular-grid output with the reference's: the product's SOURCE against the reference, on the CPU, on the golden scenes and on random ones, without the restated
oracle in between.  (The GPU tests compare the same code as it runs on the device; this one runs in every CPU round.)

Also checked here: the closed-form pass of the split assembly (buildSimpleRow = k_assemble_simple) writes bit-identical rows, entry
for entry, to the generic builder it replaces, and the hashed accumulator equals the linear one."""
import ctypes as C
import importlib.util
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from adaptiveviscositysolver_b200 import scenes
from oracle import avs_oracle as orc
from oracle import avs_ref as ref

ROOT = Path(__file__).resolve().parent.parent
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
pytestmark = [pytest.mark.skipif(ref.build() is None, reason="oracle/_ref/libavs_ref.so not built and /root/reference not present to build it"),
              pytest.mark.skipif(not Path(NVCC).exists(), reason="nvcc not available")]
MAX_LEVELS = 10


class HostField(C.Structure):
    _fields_ = [("data", C.c_void_p), ("res", C.c_int32 * 3), ("org", C.c_double * 3), ("dx", C.c_double), ("constant", C.c_float)]


class HostSceneDesc(C.Structure):
    _fields_ = [("N", C.c_int32 * 3), ("levels", C.c_int32), ("origin", C.c_double * 3), ("dx", C.c_double), ("dt", C.c_double),
                ("extrapolation", C.c_double), ("enhanced", C.c_int32),
                ("viscosity", HostField), ("density", HostField), ("collisionVel", HostField * 3), ("faceW", HostField * 3),
                ("centerW", C.c_void_p), ("edgeW", C.c_void_p * 3),
                ("label", C.c_void_p * MAX_LEVELS), ("face", (C.c_void_p * 3) * MAX_LEVELS), ("edge", (C.c_void_p * 3) * MAX_LEVELS),
                ("center", C.c_void_p * MAX_LEVELS), ("vel", HostField * 3), ("regular", C.c_void_p * 3)]


def build_harness():
    """tests/_build/libhost_assembly.so, libhost_prolong.so: tests/host_*.cu (which #include the product's avs_system.cu /
    avs_prolong.cu) compiled by nvcc for host + sm_100a."""
    lib = ROOT / "adaptiveviscositysolver_b200" / "libavs_b200.so"
    if not lib.exists():
        import __graft_entry__
        __graft_entry__.build()
    csrc = ROOT / "adaptiveviscositysolver_b200" / "csrc"
    common = [ROOT / "tests" / "host_scene.h", csrc / "avs_common.cuh", csrc / "avs_rowacc.cuh", csrc / "avs_context.h"]

    def build(name, product_source):
        out = ROOT / "tests" / "_build" / f"lib{name}.so"
        out.parent.mkdir(exist_ok=True)
        deps = [ROOT / "tests" / f"{name}.cu", csrc / product_source] + common
        if not out.exists() or out.stat().st_mtime < max(d.stat().st_mtime for d in deps):
            cmd = [NVCC, "-DAVS_HOST_TEST", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false", "-ccbin", "/usr/bin/g++",
                   "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-o", str(out), str(deps[0]),
                   "-L", str(lib.parent), "-lavs_b200", "-Xlinker", "-rpath", "-Xlinker", str(lib.parent)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        return C.CDLL(str(out))

    L = build("host_assembly", "avs_system.cu")
    L.host_assemble_rows.restype = C.c_longlong
    L.host_assemble_rows.argtypes = [C.POINTER(HostSceneDesc), C.c_longlong, C.c_void_p, C.c_int] + [C.c_void_p] * 6
    L.host_restrict_rows.restype = None
    L.host_restrict_rows.argtypes = [C.POINTER(HostSceneDesc), C.c_longlong, C.c_void_p, C.c_void_p]
    P = build("host_prolong", "avs_prolong.cu")
    P.host_apply_regular.restype = C.c_longlong
    P.host_apply_regular.argtypes = [C.POINTER(HostSceneDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.prolong = P
    B = build("host_labels", "avs_labels.cu")
    B.host_weights.restype = None
    B.host_weights.argtypes = [C.POINTER(HostSceneDesc), C.POINTER(HostField), C.POINTER(HostField), C.c_int, C.c_int] + [C.c_void_p] * 4
    B.host_classify.restype = None
    B.host_classify.argtypes = [C.POINTER(HostSceneDesc), C.POINTER(HostField), C.POINTER(HostField)] + [C.c_void_p] * 4
    B.host_octree.restype = C.c_int
    B.host_octree.argtypes = [C.POINTER(HostSceneDesc), C.POINTER(HostField), C.POINTER(HostField), C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_int]
    L.labels = B
    return L


@pytest.fixture(scope="module")
def harness():
    return build_harness()


def _field(f, keep):
    h = HostField()
    if f.data is None:
        h.data = None
        h.res[:] = [1, 1, 1]
    else:
        a = np.ascontiguousarray(f.data, np.float32)
        keep.append(a)
        h.data = a.ctypes.data
        nz, ny, nx = a.shape
        h.res[:] = [nx, ny, nz]
    h.org[:] = list(f.org)
    h.dx = float(f.dx)
    h.constant = float(f.constant)
    return h


def _describe(sc, p, R, keep):
    """The reference's label grids and weights + the caller's fields, as the row builder reads them."""
    d = HostSceneDesc()
    d.N[:] = list(sc.res)
    d.levels = R.levels
    d.origin[:] = list(sc.origin)
    d.dx, d.dt, d.extrapolation, d.enhanced = float(sc.dx), float(p.dt), float(p.extrapolation), int(p.use_enhanced_gradients)
    d.viscosity, d.density = _field(sc.viscosity, keep), _field(sc.density, keep)
    for a in range(3):
        d.collisionVel[a] = _field(sc.collision_vel[a], keep)
        d.faceW[a] = _field(sc.face_weights[a], keep)
        d.vel[a] = _field(sc.vel[a], keep)
    pad = [1 << int(np.ceil(np.log2(n))) if n > 1 else 1 for n in sc.res]

    def put(arr, dtype, shape_xyz):
        a = np.ascontiguousarray(arr, dtype)
        assert a.shape == tuple(reversed(shape_xyz)), (a.shape, shape_xyz)
        keep.append(a)
        return a.ctypes.data

    d.centerW = put(R.center_weights(), np.float32, list(sc.res))
    for a in range(3):
        d.edgeW[a] = put(R.edge_weights(a), np.float32, [sc.res[k] + (k != a) for k in range(3)])
    for l in range(R.levels):
        cells = [pd >> l for pd in pad]
        d.label[l] = put(R.labels(l), np.uint8, cells)
        for a in range(3):
            fi = R.face_index(l, a)
            assert fi.max() < 2 ** 31
            d.face[l][a] = put(fi, np.int32, [cells[k] + (k == a) for k in range(3)])
            d.edge[l][a] = put(np.minimum(R.edge_index(l, a), 0), np.int8, [cells[k] + (k != a) for k in range(3)])
        d.center[l] = put(np.minimum(R.center_index(l), 0), np.int8, cells)
    for a in range(3):
        d.regular[a] = put(np.minimum(R.regular_index(a), 0), np.int8, [sc.res[k] + (k == a) for k in range(3)])
    return d


def _assemble(L, d, keys, mode):
    n = keys.shape[0]
    M = L.host_max_row()
    count, simple = np.zeros(n, np.int32), np.zeros(n, np.int32)
    col, val = np.full((n, M), -99, np.int32), np.zeros((n, M))
    rhs, mass = np.zeros(n), np.zeros(n)
    k32 = np.ascontiguousarray(keys, np.int32)
    rc = L.host_assemble_rows(C.byref(d), n, k32.ctypes.data, mode, count.ctypes.data, col.ctypes.data, val.ctypes.data, rhs.ctypes.data,
                              mass.ctypes.data, simple.ctypes.data)
    assert rc == 0, f"row {-rc - 1} overflowed MAX_ROW = {M}"
    return count, col, val, rhs, mass, simple


def check_product_rows_against_reference(L, sc, p):
    R = ref.RefRun(sc, p)
    assert R.returned_true and not R.errors and R.n_face > 0
    keep = []
    d = _describe(sc, p, R, keep)
    keys = R.face_keys()
    count, col, val, rhs, mass, _ = _assemble(L, d, keys, 0)
    ptr, rcol, rval = R.csr()
    # ---- the matrix: same columns per row, same values (the reference's setFromTriplets sums a row's duplicates per column in push
    # order -- the order the builder adds them in)
    assert np.array_equal(count, np.diff(ptr)), "entries per row"
    n = keys.shape[0]
    M = col.shape[1]
    mask = np.arange(M)[None, :] < count[:, None]
    order = np.argsort(np.where(mask, col, np.iinfo(np.int32).max), axis=1, kind="stable")
    scol = np.take_along_axis(col, order, axis=1)[mask]
    sval = np.take_along_axis(val, order, axis=1)[mask]
    assert np.array_equal(scol, rcol), "columns"
    assert np.array_equal(sval, rval), "matrix values (bit for bit)"
    # ---- the right-hand side: boundary terms + M_u u^n (the kernel k_finish_rhs adds the second part from the restricted velocity)
    assert np.array_equal(rhs + mass * R.x0(), R.rhs()), "right-hand side (bit for bit)"
    # ---- the restricted velocity (stage 8): leaves in the reference's order on levels 0 and 1 (bit for bit), closed-form hat weights above
    x0 = np.zeros(n)
    k32 = np.ascontiguousarray(keys, np.int32)
    L.host_restrict_rows(C.byref(d), n, k32.ctypes.data, x0.ctypes.data)
    fine = keys[:, 0] <= 1
    assert np.array_equal(x0[fine], R.x0()[fine]), "restriction, levels 0 and 1 (bit for bit)"
    assert np.abs(x0 - R.x0()).max() <= 1e-13 * max(1.0, np.abs(R.x0()).max()), "restriction, hat form"
    # ---- stage 1: the super-sampled integration weights (dense sampler; the GPU's sign-class shortcuts are not involved)
    surf, coll = _field(sc.surface, keep), _field(sc.collision, keep)
    wres = [list(sc.res)] + [[sc.res[k] + (k != a) for k in range(3)] for a in range(3)]
    w = [np.zeros(tuple(reversed(r)), np.float32) for r in wres]
    L.labels.host_weights(C.byref(d), C.byref(surf), C.byref(coll), int(p.number_super_samples), int(p.do_apply_solid_weights),
                          *[a.ctypes.data for a in w])
    assert np.array_equal(w[0], R.center_weights()), "centre weights (bit for bit)"
    for a in range(3):
        assert np.array_equal(w[1 + a], R.edge_weights(a)), f"edge weights, axis {a} (bit for bit)"
    # ---- stages 2 and 3: refinement mask + octree (cell labels of every level, number of levels built)
    pad = [1 << int(np.ceil(np.log2(n_))) if n_ > 1 else 1 for n_ in sc.res]
    alloc = max(1, min(int(p.octree_levels), min(int(np.log2(q)) for q in pad), MAX_LEVELS))
    for scalar_passes in (0, 1):     # the 16-cells-per-thread pass functions where the rows allow it (the stage's choice) / the scalar ones
        labs = [np.full(tuple(reversed([q >> l for q in pad])), 77, np.uint8) for l in range(alloc)]
        allocated = C.c_int(0)
        built = L.labels.host_octree(C.byref(d), C.byref(surf), C.byref(coll), int(p.octree_levels), int(p.fine_bandwidth),
                                     (C.c_void_p * alloc)(*[x.ctypes.data for x in labs]), C.byref(allocated), scalar_passes)
        assert allocated.value == alloc and built == R.levels, (allocated.value, alloc, built, R.levels)
        for l in range(R.levels):
            assert np.array_equal(labs[l], R.labels(l)), f"cell labels, level {l} (scalar passes: {scalar_passes})"
    # ---- stages 4 and 5: regular-grid, face, edge and centre labels from the reference's cell labels and weights
    reg = [np.full(tuple(reversed([sc.res[k] + (k == a) for k in range(3)])), 99, np.int8) for a in range(3)]
    faces, edges, centers = [], [], []
    for l in range(R.levels):
        for a in range(3):
            faces.append(np.full(R.face_index(l, a).shape, 99, np.int8))
            edges.append(np.full(R.edge_index(l, a).shape, 99, np.int8))
        centers.append(np.full(R.center_index(l).shape, 99, np.int8))
    ptrs = lambda arrs: (C.c_void_p * len(arrs))(*[x.ctypes.data for x in arrs])
    L.labels.host_classify(C.byref(d), C.byref(surf), C.byref(coll), ptrs(reg), ptrs(faces), ptrs(edges), ptrs(centers))
    for a in range(3):
        assert np.array_equal(reg[a], np.minimum(R.regular_index(a), 0)), f"regular-grid labels, axis {a}"
    for l in range(R.levels):
        for a in range(3):
            assert np.array_equal(faces[3 * l + a], np.minimum(R.face_index(l, a), 0)), f"face labels, level {l} axis {a}"
            assert np.array_equal(edges[3 * l + a], np.minimum(R.edge_index(l, a), 0)), f"edge labels, level {l} axis {a}"
        assert np.array_equal(centers[l], np.minimum(R.center_index(l), 0)), f"centre labels, level {l}"
    # ---- stage 11: the interpolator's node pyramid + interpSPGrid + the write-back rule, fed with the REFERENCE's solution vector
    sol = np.ascontiguousarray(R.solution(), np.float64)
    out = [np.ascontiguousarray(v.data.copy(), np.float32) for v in sc.vel]
    interpolated = L.prolong.host_apply_regular(C.byref(d), sol.ctypes.data, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data)
    for a in range(3):
        assert np.array_equal(out[a], R.out_velocity(a)), f"regular-grid velocity, axis {a} (bit for bit)"
    assert interpolated > 0 or R.levels == 1
    # ---- accumulators and the split assembly: same entries in the same order, bit for bit
    c1 = _assemble(L, d, keys, 1)
    c2 = _assemble(L, d, keys, 2)
    for other, what in ((c1, "linear accumulator"), (c2, "split assembly")):
        assert np.array_equal(other[0], count), what
        assert np.array_equal(np.where(mask, other[1], 0), np.where(mask, col, 0)), what
        assert np.array_equal(np.where(mask, other[2], 0.0), np.where(mask, val, 0.0)), what
        assert np.array_equal(other[3], rhs) and np.array_equal(other[4], mass), what
    simple = c2[5]
    assert simple[keys[:, 0] > 0].sum() == 0
    return R, float(simple.mean())


def _golden():
    spec = importlib.util.spec_from_file_location("make_golden", ROOT / "tests" / "golden" / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


@pytest.mark.parametrize("name", ["c1_uniform32", "sphere64_l3_noise", "padded_48x64x40_l5_varmu", "solid_ground32_l3", "buckling_f6_dx2mm"])
def test_product_row_builder_equals_reference_on_the_golden_scenes(harness, name):
    mg = _golden()
    case = mg.CASES[name]
    sc = getattr(scenes, case.get("maker", "sphere_drop"))(**case["scene"])
    p = orc.OracleParams(octree_levels=case["levels"], tolerance=mg.TOL, dt=case.get("dt", 1.0 / 24.0), max_iterations=1)
    R, simple_fraction = check_product_rows_against_reference(harness, sc, p)
    if name == "c1_uniform32":
        assert simple_fraction > 0.5          # a uniform grid: most rows take the closed-form pass


_spec = importlib.util.spec_from_file_location("fuzz_reference_pin", ROOT / "scripts" / "fuzz_reference_pin.py")
fz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fz)

# random scenes of scripts/fuzz_reference_pin.py (solids on their own collision grid, solid weights, variable viscosity / density,
# non-cubic grids, shifted origins, 65-entry rows at 25 / 48 / 135, deep and distorted-SDF variants)
# 4017 / 4022 / 4046: the solid's velocity as sampled fields on a grid of their own (variant 4)
SEEDS = [0, 4, 9, 15, 25, 29, 34, 43, 48, 57, 61, 70, 73, 77, 89, 100, 135, 192, 1003, 1011, 2005, 2010, 4017, 4022, 4046]


@pytest.mark.parametrize("seed", SEEDS)
def test_product_row_builder_equals_reference_on_a_random_scene(harness, seed):
    sc, p, desc = fz.fuzz_case(seed)
    p.max_iterations = 1                      # the solve is not the subject here
    check_product_rows_against_reference(harness, sc, p)


def test_variants_of_the_options(harness):
    for kw in (dict(use_enhanced_gradients=False), dict(do_apply_solid_weights=True), dict(fine_bandwidth=4, number_super_samples=2)):
        sc = scenes.sphere_drop(32, 9, center=(0.5, 0.34, 0.5), ground_height=0.125, ground_velocity=(0.1, 0.0, -0.2))
        check_product_rows_against_reference(harness, sc, orc.OracleParams(octree_levels=3, max_iterations=1, **kw))


def test_options_at_the_ends_of_their_ranges(harness):
    """The library's device functions on the option edge values tests/test_reference_pin.py holds the oracle to (1 and 9 levels, fine
    band 0 / 12, one super-sample, far extrapolation, dt 1e-6 / 100): weights, octree, labels, matrix, rhs, restriction, output bit for bit."""
    sc = scenes.sphere_drop(32, 10, noise=0.01)
    for kw in (dict(octree_levels=1), dict(octree_levels=9), dict(fine_bandwidth=0), dict(fine_bandwidth=12), dict(number_super_samples=1),
               dict(extrapolation=3.0), dict(dt=1e-6), dict(dt=100.0)):
        k = dict(octree_levels=3, tolerance=1e-6, max_iterations=1)
        k.update(kw)
        check_product_rows_against_reference(harness, sc, orc.OracleParams(**k))
