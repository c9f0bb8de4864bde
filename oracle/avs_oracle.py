"""ctypes front-end of the CPU oracle (``oracle/avs_oracle.cpp``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package never
imports this module.  Parity status of the oracle itself: pinned bit for bit to the reference's own
sources compiled against the stand-ins of oracle/mock_hdk (oracle/_ref, tests/test_reference_pin.py);
Houdini's and Eigen's behaviour is assumed -- see the header of avs_oracle.cpp.

Scenes are duck-typed: any object with the attributes used in ``_scene_struct`` works
(``adaptiveviscositysolver_b200.scenes.Scene`` is what the tests pass).  Arrays are numpy
float32 with shape (nz, ny, nx), C-contiguous, i.e. x-fastest like the reference's flat
voxel order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libavs_oracle.so"

INACTIVE, ACTIVE, UP, DOWN = 0, 1, 2, 3
FLUID, UNASSIGNED, SOLIDBOUNDARY, OUTSIDE = 0, -1, -2, -3


def build(force: bool = False) -> Path:
    """Compile the oracle with the Makefile next to this file (g++ -fopenmp)."""
    src = _HERE / "avs_oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


class _Field(C.Structure):
    _fields_ = [("data", C.c_void_p), ("res", C.c_int * 3), ("org", C.c_double * 3),
                ("dx", C.c_double), ("constant", C.c_float)]


class _Scene(C.Structure):
    _fields_ = [("res", C.c_int * 3), ("origin", C.c_double * 3), ("dx", C.c_double),
                ("surface", _Field), ("vel", _Field * 3), ("faceWeights", _Field * 3),
                ("viscosity", _Field), ("density", _Field), ("collision", _Field),
                ("collisionVel", _Field * 3)]


class _Params(C.Structure):
    _fields_ = [("dt", C.c_double), ("tolerance", C.c_double), ("extrapolation", C.c_double),
                ("maxIterations", C.c_int), ("numberSuperSamples", C.c_int), ("octreeLevels", C.c_int),
                ("fineBandwidth", C.c_int), ("useEnhancedGradients", C.c_int),
                ("doApplySolidWeights", C.c_int), ("singlePrecision", C.c_int), ("stopAfterStage", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Scene), C.POINTER(_Params)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_run.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_weight_shortcut.argtypes = [C.c_void_p, C.c_int]
        L.orc_levels.argtypes = [C.c_void_p]
        L.orc_levels_allocated.argtypes = [C.c_void_p]
        L.orc_padded_res.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_count.restype = C.c_int64
        L.orc_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_error.restype = C.c_double
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_get_float.restype = C.c_int64
        L.orc_get_float.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_get_labels.restype = C.c_int64
        L.orc_get_labels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_get_index_grid.restype = C.c_int64
        L.orc_get_index_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_get_face_keys.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_vector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_out_velocity.restype = C.c_int64
        L.orc_get_out_velocity.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_get_node_grid.restype = C.c_int64
        L.orc_get_node_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_interpolated_faces.restype = C.c_int64
        L.orc_interpolated_faces.argtypes = [C.c_void_p]
        L.orc_get_octree_points.restype = C.c_int64
        L.orc_get_octree_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stencil.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                                                C.c_void_p, C.POINTER(C.c_double)]
        L.orc_spmv_f64.argtypes = [C.c_int64] + [C.c_void_p] * 5
        L.orc_spmv_f32.argtypes = [C.c_int64] + [C.c_void_p] * 5
        for f in (L.orc_cg_f64, L.orc_cg_f32):
            f.argtypes = [C.c_int64] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _field_struct(f, keep):
    s = _Field()
    if f.data is None:
        s.data = None
        s.res[:] = (1, 1, 1)
        s.constant = float(f.constant)
    else:
        a = np.ascontiguousarray(f.data, dtype=np.float32)
        keep.append(a)
        s.data = a.ctypes.data
        nz, ny, nx = a.shape
        s.res[:] = (nx, ny, nz)
        s.constant = 0.0
    s.org[:] = tuple(float(v) for v in f.org)
    s.dx = float(f.dx)
    return s


def _scene_struct(scene, keep):
    s = _Scene()
    s.res[:] = tuple(int(v) for v in scene.res)
    s.origin[:] = tuple(float(v) for v in scene.origin)
    s.dx = float(scene.dx)
    s.surface = _field_struct(scene.surface, keep)
    for a in range(3):
        s.vel[a] = _field_struct(scene.vel[a], keep)
        s.faceWeights[a] = _field_struct(scene.face_weights[a], keep)
        s.collisionVel[a] = _field_struct(scene.collision_vel[a], keep)
    s.viscosity = _field_struct(scene.viscosity, keep)
    s.density = _field_struct(scene.density, keep)
    s.collision = _field_struct(scene.collision, keep)
    return s


@dataclass
class OracleParams:
    """Mirror of the reference's option getters (AV.h:28-41) with its effective defaults."""
    dt: float = 1.0 / 24.0
    tolerance: float = 1e-3
    max_iterations: int = 2500
    number_super_samples: int = 3
    octree_levels: int = 4
    fine_bandwidth: int = 0
    use_enhanced_gradients: bool = True
    do_apply_solid_weights: bool = False
    extrapolation: float = 0.5
    single_precision: bool = False


class OracleRun:
    """One execution of the restated reference pipeline; exposes every intermediate."""

    def __init__(self, scene, params: OracleParams | None = None, stop_after_stage: int = 0, weight_shortcut: bool = True):
        params = params or OracleParams()
        self._L = lib()
        keep = []
        sc = _scene_struct(scene, keep)
        p = _Params(params.dt, params.tolerance, params.extrapolation, params.max_iterations,
                    params.number_super_samples, params.octree_levels, params.fine_bandwidth,
                    int(params.use_enhanced_gradients), int(params.do_apply_solid_weights),
                    int(params.single_precision), stop_after_stage)
        self._h = self._L.orc_create(C.byref(sc), C.byref(p))
        del keep  # the oracle copies its inputs
        self._L.orc_set_weight_shortcut(self._h, int(weight_shortcut))
        self._L.orc_run(self._h, stop_after_stage)
        self.stage = stop_after_stage if stop_after_stage > 0 else 11

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    # -- scalars
    @property
    def levels(self):
        return self._L.orc_levels(self._h)

    @property
    def padded_res(self):
        r = (C.c_int * 3)()
        self._L.orc_padded_res(self._h, r)
        return tuple(r)

    @property
    def n_face(self):
        return self._L.orc_count(self._h, 0)

    @property
    def n_edge(self):
        return self._L.orc_count(self._h, 1)

    @property
    def n_center(self):
        return self._L.orc_count(self._h, 2)

    @property
    def regular_dofs(self):
        return self._L.orc_count(self._h, 3)

    @property
    def nnz(self):
        return self._L.orc_count(self._h, 4)

    @property
    def iterations(self):
        return self._L.orc_count(self._h, 5)

    @property
    def error(self):
        return self._L.orc_error(self._h)

    # -- arrays (all returned with shape (nz, ny, nx))
    def _float(self, kind):
        r = (C.c_int * 3)()
        n = self._L.orc_get_float(self._h, kind, None, r)
        out = np.empty(n, np.float32)
        self._L.orc_get_float(self._h, kind, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def center_weights(self):
        return self._float(0)

    def edge_weights(self, axis):
        return self._float(1 + axis)

    def mask(self):
        return self._float(4)

    def labels(self, level):
        r = (C.c_int * 3)()
        n = self._L.orc_get_labels(self._h, level, None, r)
        out = np.empty(n, np.uint8)
        self._L.orc_get_labels(self._h, level, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def _grid(self, kind, level, axis):
        r = (C.c_int * 3)()
        n = self._L.orc_get_index_grid(self._h, kind, level, axis, None, r)
        out = np.empty(n, np.int64)
        self._L.orc_get_index_grid(self._h, kind, level, axis, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def face_index(self, level, axis):
        return self._grid(0, level, axis)

    def edge_index(self, level, axis):
        return self._grid(1, level, axis)

    def center_index(self, level):
        return self._grid(2, level, 0)

    def regular_index(self, axis):
        return self._grid(3, 0, axis)

    def face_keys(self):
        """(n_face, 5) int32: level, axis, i, j, k of every octree velocity DOF."""
        out = np.empty((self.n_face, 5), np.int32)
        self._L.orc_get_face_keys(self._h, out.ctypes.data)
        return out

    def _vec(self, what):
        out = np.empty(self.n_face, np.float64)
        self._L.orc_get_vector(self._h, what, out.ctypes.data)
        return out

    def x0(self):
        return self._vec(0)

    def rhs(self):
        return self._vec(1)

    def solution(self):
        return self._vec(2)

    def out_velocity(self, axis):
        """Regular-grid velocity component after the solve (what solveGasSubclass leaves in ``vel``)."""
        r = (C.c_int * 3)()
        n = self._L.orc_get_out_velocity(self._h, axis, None, r)
        out = np.empty(n, np.float32)
        self._L.orc_get_out_velocity(self._h, axis, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def node_grid(self, kind, level):
        r = (C.c_int * 3)()
        n = self._L.orc_get_node_grid(self._h, kind, level, None, r)
        out = np.empty(n, np.float32)
        self._L.orc_get_node_grid(self._h, kind, level, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    @property
    def interpolated_faces(self):
        return self._L.orc_interpolated_faces(self._h)

    def octree_points(self):
        """HDK_OctreeGrid::outputOctreeGeometry (OG.cpp:245-308): (P (n,3) f32, pscale (n,) f32, octreeLevel (n,) i32)."""
        n = self._L.orc_get_octree_points(self._h, None, None, None)
        pos = np.empty((n, 3), np.float32)
        pscale = np.empty(n, np.float32)
        level = np.empty(n, np.int32)
        if n:
            self._L.orc_get_octree_points(self._h, pos.ctypes.data, pscale.ctypes.data, level.ctypes.data)
        return pos, pscale, level

    def csr(self):
        n, nnz = self.n_face, self.nnz
        ptr = np.empty(n + 1, np.int64)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        self._L.orc_get_csr(self._h, ptr.ctypes.data, col.ctypes.data, val.ctypes.data)
        return ptr, col, val

    def scipy_matrix(self):
        import scipy.sparse as sp
        ptr, col, val = self.csr()
        return sp.csr_matrix((val, col, ptr), shape=(self.n_face, self.n_face))

    def stencil(self, kind, level, axis, i, j, k):
        """Row of D for an edge (kind=0) or centre (kind=1) stress: (idx, coef, boundary, weight)."""
        idx = np.empty(40, np.int64)
        coef = np.empty(40, np.float64)
        bnd = np.empty(8, np.float64)
        nb = C.c_int()
        w = C.c_double()
        n = self._L.orc_stencil(self._h, kind, level, axis, i, j, k, idx.ctypes.data, coef.ctypes.data,
                                C.byref(nb), bnd.ctypes.data, C.byref(w))
        return idx[:n].copy(), coef[:n].copy(), bnd[:nb.value].copy(), w.value


# ---- stand-alone linear algebra (the CPU baseline of the CG hot loop) ------------------------------
def spmv(ptr, col, val, x):
    L = lib()
    ptr = np.ascontiguousarray(ptr, np.int64)
    col = np.ascontiguousarray(col, np.int32)
    n = ptr.size - 1
    if val.dtype == np.float32:
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty(n, np.float32)
        L.orc_spmv_f32(n, ptr.ctypes.data, col.ctypes.data, val.ctypes.data, x.ctypes.data, y.ctypes.data)
    else:
        val = np.ascontiguousarray(val, np.float64)
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(n, np.float64)
        L.orc_spmv_f64(n, ptr.ctypes.data, col.ctypes.data, val.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y


def cg(ptr, col, val, b, x0, tol, max_iters):
    """Eigen-equivalent Jacobi-PCG (AV.cpp:611-630). Returns (x, iterations, error)."""
    L = lib()
    ptr = np.ascontiguousarray(ptr, np.int64)
    col = np.ascontiguousarray(col, np.int32)
    n = ptr.size - 1
    it = C.c_int()
    err = C.c_double()
    if val.dtype == np.float32:
        b = np.ascontiguousarray(b, np.float32)
        x = np.array(x0, np.float32, copy=True)
        L.orc_cg_f32(n, ptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data, x.ctypes.data,
                     tol, max_iters, C.byref(it), C.byref(err))
    else:
        val = np.ascontiguousarray(val, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        x = np.array(x0, np.float64, copy=True)
        L.orc_cg_f64(n, ptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data, x.ctypes.data,
                     tol, max_iters, C.byref(it), C.byref(err))
    return x, it.value, err.value


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
