"""ctypes front-end of oracle/_ref/libavs_ref.so: the REFERENCE's own sources (/root/reference/Source/*.cpp, unchanged)
compiled against the Houdini / Eigen stand-ins in oracle/mock_hdk (see mock_hdk.h for what is real and what is assumed).

TEST INFRASTRUCTURE ONLY -- same rule as oracle/avs_oracle.py: only tests/ and bench.py's CPU arm may import this.
/root/reference exists only in the build container, so the library is built THERE (oracle/Makefile target `ref`,
__graft_entry__.build()) and travels to the GPU box as a binary; nothing here reads /root/reference at run time.

``RefRun`` mirrors ``avs_oracle.OracleRun`` accessor for accessor, so a test can run the restated oracle and the compiled
reference side by side on the same scene.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from .avs_oracle import OracleParams, _Params, _Scene, _scene_struct

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_ref" / "libavs_ref.so"
REFERENCE_SOURCES = Path("/root/reference/Source")


def available() -> bool:
    return _LIB_PATH.exists()


def build(force: bool = False) -> Path | None:
    """Compiles the reference from where it lies (needs /root/reference); returns None when neither sources nor a prebuilt
    library are present (the GPU box without a prebuilt file)."""
    if _LIB_PATH.exists() and not force:
        return _LIB_PATH
    if not REFERENCE_SOURCES.exists():
        return None
    r = subprocess.run(["make", "-C", str(_HERE), "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return _LIB_PATH


_SHIM_PATH = _HERE / "_ref" / "libavs_shim.so"


def build_shim(force: bool = False) -> Path:
    """oracle/_ref/libavs_shim.so: THIS repository's Houdini-side shim (integration/hdk/HDK_AdaptiveViscosityB200.cpp) compiled
    against the same stand-ins and linked to libavs_b200.so -- needs no reference source, only g++ and the built CUDA library."""
    if _SHIM_PATH.exists() and not force:
        return _SHIM_PATH
    r = subprocess.run(["make", "-C", str(_HERE), "shim"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref shim build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return _SHIM_PATH


_lib = None
_shim = None


def _bind(L):
    _bind_signatures(L)
    return L


def shim_lib():
    global _shim
    if _shim is None:
        _shim = _bind(C.CDLL(str(build_shim())))
    return _shim


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libavs_ref.so is missing and /root/reference is not available to build it")
        _lib = _bind(C.CDLL(str(_LIB_PATH)))
    return _lib


def _bind_signatures(L):
    L.ref_create.restype = C.c_void_p
    L.ref_create.argtypes = [C.POINTER(_Scene), C.POINTER(_Params)]
    L.ref_destroy.argtypes = [C.c_void_p]
    L.ref_run.argtypes = [C.c_void_p]
    L.ref_tamper.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.ref_set_threads.argtypes = [C.c_int]
    L.ref_set_weight_shortcut.argtypes = [C.c_int]
    L.ref_error_count.argtypes = [C.c_void_p]
    L.ref_error_text.restype = C.c_char_p
    L.ref_error_text.argtypes = [C.c_void_p, C.c_int]
    L.ref_extra_info.restype = C.c_char_p
    L.ref_extra_info.argtypes = [C.c_void_p]
    L.ref_levels.argtypes = [C.c_void_p]
    L.ref_count.restype = C.c_int64
    L.ref_count.argtypes = [C.c_void_p, C.c_int]
    L.ref_error.restype = C.c_double
    L.ref_error.argtypes = [C.c_void_p]
    L.ref_get_float.restype = C.c_int64
    L.ref_get_float.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    L.ref_get_labels.restype = C.c_int64
    L.ref_get_labels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    L.ref_get_index_grid.restype = C.c_int64
    L.ref_get_index_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    L.ref_get_face_keys.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_get_vector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_get_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_get_out_velocity.restype = C.c_int64
    L.ref_get_out_velocity.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    L.ref_get_octree_points.restype = C.c_int64
    L.ref_get_octree_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def set_threads(n: int) -> None:
    """Jobs per THREADED_METHOD of the reference (UT_ThreadedAlgorithm stand-in); 1 = deterministic triplet order."""
    lib().ref_set_threads(int(n))


class RefRun:
    """One call of the reference's HDK_AdaptiveViscosity::solveGasSubclass (HDK_AdaptiveViscosity.cpp:126-707) on the stand-in
    fields; exposes what the reference computed on the way."""

    def _library(self):
        return lib()

    _precision_word = 0

    # names the reference looks its fields up by (GAS_NAME_* of the stand-ins; AV.cpp:138-144, 203, 218)
    FIELD_NAMES = {"surface": b"surface", "velocity": b"velocity", "collision": b"collision", "collisionvel": b"collisionvel",
                   "faceWeights": b"faceWeights", "viscosity": b"viscosity", "density": b"density"}
    _TAMPER_OPS = {"remove": 0, "misalign": 1, "unstagger": 2}

    def __init__(self, scene, params: OracleParams | None = None, octree_only: bool = False, weight_shortcut: bool = True, tamper=()):
        """``tamper``: (op, field) pairs applied to the object before solveGasSubclass runs -- op "remove", "misalign" or
        "unstagger", field a key of FIELD_NAMES -- for the validation paths AV.cpp:152-229."""
        params = params or OracleParams()
        if params.single_precision and self._precision_word == 0:
            raise ValueError("libavs_ref.so is built without USESINGLEPRECISION")
        self._L = self._library()
        keep = []
        sc = _scene_struct(scene, keep)
        p = _Params(params.dt, params.tolerance, params.extrapolation, params.max_iterations,
                    params.number_super_samples, params.octree_levels, params.fine_bandwidth,
                    int(params.use_enhanced_gradients), int(params.do_apply_solid_weights),
                    self._precision_word | int(params.single_precision), 3 if octree_only else 0)
        self._L.ref_set_weight_shortcut(int(weight_shortcut))
        self._h = self._L.ref_create(C.byref(sc), C.byref(p))
        del keep
        for op, field in tamper:
            if self._L.ref_tamper(self._h, self._TAMPER_OPS[op], self.FIELD_NAMES[field]) != 0:
                raise ValueError(f"cannot {op} field {field}")
        self.returned_true = self._L.ref_run(self._h) == 0
        self.errors = [self._L.ref_error_text(self._h, i).decode() for i in range(self._L.ref_error_count(self._h))]
        self.extra_info = self._L.ref_extra_info(self._h).decode()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_destroy(self._h)
            self._h = None

    levels = property(lambda self: self._L.ref_levels(self._h))
    n_face = property(lambda self: self._L.ref_count(self._h, 0))
    n_edge = property(lambda self: self._L.ref_count(self._h, 1))
    n_center = property(lambda self: self._L.ref_count(self._h, 2))
    regular_dofs = property(lambda self: self._L.ref_count(self._h, 3))
    nnz = property(lambda self: self._L.ref_count(self._h, 4))
    iterations = property(lambda self: self._L.ref_count(self._h, 5))
    error = property(lambda self: self._L.ref_error(self._h))

    def _float(self, kind):
        r = (C.c_int * 3)()
        n = self._L.ref_get_float(self._h, kind, None, r)
        out = np.empty(n, np.float32)
        self._L.ref_get_float(self._h, kind, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def center_weights(self):
        return self._float(0)

    def edge_weights(self, axis):
        return self._float(1 + axis)

    def labels(self, level):
        r = (C.c_int * 3)()
        n = self._L.ref_get_labels(self._h, level, None, r)
        out = np.empty(n, np.uint8)
        self._L.ref_get_labels(self._h, level, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def _grid(self, kind, level, axis):
        r = (C.c_int * 3)()
        n = self._L.ref_get_index_grid(self._h, kind, level, axis, None, r)
        out = np.empty(n, np.int64)
        self._L.ref_get_index_grid(self._h, kind, level, axis, out.ctypes.data, r)
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
        out = np.empty((self.n_face, 5), np.int32)
        self._L.ref_get_face_keys(self._h, out.ctypes.data)
        return out

    def _vec(self, what):
        out = np.empty(self.n_face, np.float64)
        self._L.ref_get_vector(self._h, what, out.ctypes.data)
        return out

    def x0(self):
        return self._vec(0)

    def rhs(self):
        return self._vec(1)

    def solution(self):
        return self._vec(2)

    def out_velocity(self, axis):
        r = (C.c_int * 3)()
        n = self._L.ref_get_out_velocity(self._h, axis, None, r)
        out = np.empty(n, np.float32)
        self._L.ref_get_out_velocity(self._h, axis, out.ctypes.data, r)
        return out.reshape(r[2], r[1], r[0])

    def octree_points(self):
        n = self._L.ref_get_octree_points(self._h, None, None, None)
        pos = np.empty((n, 3), np.float32)
        pscale = np.empty(n, np.float32)
        level = np.empty(n, np.int32)
        if n:
            self._L.ref_get_octree_points(self._h, pos.ctypes.data, pscale.ctypes.data, level.ctypes.data)
        return pos, pscale, level

    def csr(self):
        n, nnz = self.n_face, self.nnz
        ptr = np.empty(n + 1, np.int64)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        self._L.ref_get_csr(self._h, ptr.ctypes.data, col.ctypes.data, val.ctypes.data)
        return ptr, col, val

    def scipy_matrix(self):
        import scipy.sparse as sp
        ptr, col, val = self.csr()
        return sp.csr_matrix((val, col, ptr), shape=(self.n_face, self.n_face))


_LIB32_PATH = _HERE / "_ref" / "libavs_ref_f32.so"
_lib32 = None


def build_f32(force: bool = False) -> Path | None:
    """oracle/_ref/libavs_ref_f32.so: the reference compiled with USESINGLEPRECISION (needs /root/reference; None without it)."""
    if _LIB32_PATH.exists() and not force:
        return _LIB32_PATH
    if not REFERENCE_SOURCES.exists():
        return None
    r = subprocess.run(["make", "-C", str(_HERE), "ref-f32"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref fp32 build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return _LIB32_PATH


class RefRun32(RefRun):
    """The reference's USESINGLEPRECISION build (SolveType = fpreal32: float triplets summed in float by setFromTriplets, float
    right-hand side, float conjugate gradients, HDK_Utilities.h:25-30).  The harness widens what it snapshots to double."""
    _precision_word = 1

    def _library(self):
        global _lib32
        if _lib32 is None:
            if build_f32() is None:
                raise RuntimeError("oracle/_ref/libavs_ref_f32.so is missing and /root/reference is not available to build it")
            _lib32 = _bind(C.CDLL(str(_LIB32_PATH)))
        return _lib32


class ShimRun(RefRun):
    """The same harness around THIS repository's Houdini-side shim (integration/hdk/HDK_AdaptiveViscosityB200.cpp): the stand-in
    "Houdini" calls the shim's solveGasSubclass, the shim flattens the fields and calls avs_solve / avs_solve_multi in
    libavs_b200.so, and writes the velocity back -- the drop-in boundary end to end.  Needs a GPU (its field validation, which comes
    first, does not: tests/test_validation_contract.py).  Only what the DOP itself
    exposes is available afterwards: ``out_velocity``, ``octree_points``, ``errors``, ``extra_info`` (the PerfMon string)."""

    def __init__(self, scene, params: OracleParams | None = None, octree_only: bool = False, gpus: int = 1, tamper=()):
        self._precision_word = (int(gpus) if gpus > 1 else 0) << 8
        super().__init__(scene, params, octree_only, tamper=tamper)

    def _library(self):
        return shim_lib()

    def info(self) -> dict:
        """iterations / error / DOF counts parsed from the PerfMon extra-info string (AV.cpp:645-652)."""
        out = {}
        for part in self.extra_info.split(","):
            if "=" in part:
                k, v = part.strip().split("=")
                out[k.strip()] = float(v)
        return out
