"""Host-side mirror of the reference's operator interface for the viscosity-solve path.

``HDK_AdaptiveViscosity`` keeps the reference's names, argument meaning and error behaviour
(HDK_AdaptiveViscosity.h:28-58, HDK_AdaptiveViscosity.cpp:126-231): a sub-solver object with the DOP's
options and ``solveGasSubclass(engine, obj, time, timestep) -> bool`` that looks its fields up by name
on a simulation object, reports problems through ``addError`` and returns ``False`` -- and then hands
everything after validation to the CUDA library through the C-ABI (``avs_solve``).

``Solver`` is the thin, explicit wrapper over the staged C-ABI entry points used by the parity tests
and the benchmark.  Neither class has a CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (AVS_OK, AVS_PRECISION_F32, AVS_PRECISION_F64, AvsDeviceConfig, AvsError, AvsField, AvsFields,
                   AvsParams, AvsResult, AvsVelocityOut, STAGE_NAMES)
from .scenes import SampledField, Scene


@dataclass
class Params:
    """The DOP's options (HDK_AdaptiveViscosity.h:28-41) with the reference's effective defaults."""
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
    check_every: int = 0
    cancel: Optional[object] = None   # ctypes.c_int32 polled between CG batches (UT_Interrupt::opInterrupt)

    def to_c(self) -> AvsParams:
        p = AvsParams()
        _lib.load().avs_default_params(C.byref(p))
        p.dt = self.dt
        p.tolerance = self.tolerance
        p.extrapolation = self.extrapolation
        p.max_iterations = int(self.max_iterations)
        p.number_super_samples = int(self.number_super_samples)
        p.octree_levels = int(self.octree_levels)
        p.fine_bandwidth = int(self.fine_bandwidth)
        p.use_enhanced_gradients = int(self.use_enhanced_gradients)
        p.do_apply_solid_weights = int(self.do_apply_solid_weights)
        p.precision = AVS_PRECISION_F32 if self.single_precision else AVS_PRECISION_F64
        p.check_every = int(self.check_every)
        p.cancel = C.addressof(self.cancel) if self.cancel is not None else None
        return p


@dataclass
class SolveInfo:
    status: int
    iterations: int
    error: float
    levels: int
    octree_dofs: int
    regular_dofs: int
    edge_dofs: int
    center_dofs: int
    nnz: int
    local_rows: int
    spmv_launches: int
    kernel_launches: int
    stage_ms: Dict[str, float]
    spmv_ms: float
    interpolated_faces: int
    cg_update_xr_ms: float = 0.0
    cg_update_p_ms: float = 0.0
    dist_mode: int = 0
    halo_columns: int = 0
    cg_kernel_ms: float = 0.0       # persistent CG kernel: CUDA events around its cooperative launch(es)
    cg_kernel_launches: int = 0

    @staticmethod
    def from_c(r: AvsResult) -> "SolveInfo":
        return SolveInfo(r.status, r.iterations, r.error, r.levels, r.octree_dofs, r.regular_dofs, r.edge_dofs,
                         r.center_dofs, r.nnz, r.local_rows, r.spmv_launches, r.kernel_launches,
                         {STAGE_NAMES[i]: float(r.stage_ms[i]) for i in range(11)}, float(r.spmv_ms),
                         r.interpolated_faces, float(r.cg_update_xr_ms), float(r.cg_update_p_ms), r.dist_mode,
                         r.halo_columns, float(r.cg_kernel_ms), int(r.cg_kernel_launches))


def _new_result() -> AvsResult:
    r = AvsResult()
    r.size = C.sizeof(AvsResult)
    return r


def _data_ptr(arr):
    """numpy array or torch tensor -> (pointer, on_device, keepalive)."""
    if isinstance(arr, np.ndarray):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        return a.ctypes.data, 0, a
    # torch tensor (plumbing only: device memory / pinned host memory)
    t = arr.contiguous()
    return t.data_ptr(), int(t.is_cuda), t


def _field_c(f: SampledField, keep: list) -> AvsField:
    s = AvsField()
    if f.data is None:
        s.data = None
        s.res[:] = (1, 1, 1)
        s.constant = float(f.constant)
        s.on_device = 0
        s.org[:] = (0.0, 0.0, 0.0)
        s.dx = 1.0
        return s
    ptr, dev, k = _data_ptr(f.data)
    keep.append(k)
    nz, ny, nx = f.data.shape
    s.data = ptr
    s.res[:] = (nx, ny, nz)
    s.org[:] = tuple(float(v) for v in f.org)
    s.dx = float(f.dx)
    s.constant = 0.0
    s.on_device = dev
    return s


def fields_to_c(scene: Scene, keep: list) -> AvsFields:
    s = AvsFields()
    s.size = C.sizeof(AvsFields)
    s.res[:] = tuple(int(v) for v in scene.res)
    s.origin[:] = tuple(float(v) for v in scene.origin)
    s.dx = float(scene.dx)
    s.surface = _field_c(scene.surface, keep)
    for a in range(3):
        s.vel[a] = _field_c(scene.vel[a], keep)
        s.face_weights[a] = _field_c(scene.face_weights[a], keep)
        s.collision_vel[a] = _field_c(scene.collision_vel[a], keep)
    s.viscosity = _field_c(scene.viscosity, keep)
    s.density = _field_c(scene.density, keep)
    s.collision = _field_c(scene.collision, keep)
    return s


def nccl_unique_id() -> bytes:
    """128-byte NCCL unique id (call on rank 0, broadcast to every rank, pass to ``Solver``)."""
    buf = C.create_string_buffer(128)
    rc = _lib.load().avs_nccl_unique_id(buf)
    if rc != AVS_OK:
        raise AvsError(rc, "avs_nccl_unique_id")
    return buf.raw


class Solver:
    """One ``AvsContext``: a single-threaded solver bound to one GPU (include/avs.h)."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, time_spmv: bool = False, stream: int = 0,
                 nccl_unique_id: Optional[bytes] = None, distributed_output: bool = False):
        self._L = _lib.load()
        cfg = AvsDeviceConfig()
        cfg.size = C.sizeof(AvsDeviceConfig)
        cfg.device, cfg.rank, cfg.nranks = device, rank, nranks
        self._idbuf = None
        if nranks > 1:
            if nccl_unique_id is None or len(nccl_unique_id) != 128:
                raise ValueError("nranks > 1 needs the 128-byte id from nccl_unique_id() on rank 0")
            self._idbuf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            cfg.nccl_unique_id = C.cast(self._idbuf, C.c_void_p)
        else:
            cfg.nccl_unique_id = None
        cfg.stream = stream or None
        cfg.time_spmv = int(time_spmv)
        cfg.distributed_output = int(distributed_output)
        h = C.c_void_p()
        rc = self._L.avs_create(C.byref(cfg), C.byref(h))
        if rc != AVS_OK:
            raise AvsError(rc, "avs_create", _lib.last_error())
        self._h = h

    _owned = True

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._L.avs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc != AVS_OK:
            raise AvsError(rc, where, _lib.last_error() if rc in (-5, -10) else "")   # AVS_ERR_CUDA / AVS_ERR_UNSUPPORTED carry a message

    # ---- the pipeline
    def solve(self, scene: Scene, params: Params, out: Optional[List] = None) -> SolveInfo:
        """avs_solve: stages 1-11 in one call. ``out`` = three float32 arrays (numpy or torch) shaped like
        ``scene.vel``; they are updated in place on the faces the reference would write."""
        keep: list = []
        f = fields_to_c(scene, keep)
        p = params.to_c()
        r = _new_result()
        o = None
        if out is not None:
            o = AvsVelocityOut()
            dev = 0
            for a in range(3):
                if isinstance(out[a], np.ndarray):
                    assert out[a].dtype == np.float32 and out[a].flags.c_contiguous
                    o.vel[a] = out[a].ctypes.data
                else:
                    o.vel[a] = out[a].data_ptr()
                    dev = int(out[a].is_cuda)
            o.on_device = dev
        rc = self._L.avs_solve(self._h, C.byref(f), C.byref(p), C.byref(o) if o is not None else None, C.byref(r))
        self._check(rc, "avs_solve")
        return SolveInfo.from_c(r)

    def assemble(self, scene: Scene, params: Params) -> SolveInfo:
        keep: list = []
        f = fields_to_c(scene, keep)
        p = params.to_c()
        r = _new_result()
        self._check(self._L.avs_assemble(self._h, C.byref(f), C.byref(p), C.byref(r)), "avs_assemble")
        return SolveInfo.from_c(r)

    def build_octree(self, scene: Scene, params: Params) -> SolveInfo:
        """avs_build_octree: stages 1-3 only (the reference's ``onlyPrintOctree`` early return, AV.cpp:292-293)."""
        keep: list = []
        f = fields_to_c(scene, keep)
        p = params.to_c()
        r = _new_result()
        self._check(self._L.avs_build_octree(self._h, C.byref(f), C.byref(p), C.byref(r)), "avs_build_octree")
        return SolveInfo.from_c(r)

    def octree_points(self):
        """The octree geometry dump (HDK_OctreeGrid::outputOctreeGeometry, OG.cpp:245-308): ``(P, pscale, octreeLevel)``
        with one point per ACTIVE cell -- positions (n, 3) float32, pscale (n,) float32, octreeLevel (n,) int32."""
        n = C.c_int64()
        self._check(self._L.avs_get_octree_points(self._h, C.byref(n), None, None, None), "avs_get_octree_points")
        pos = np.empty((n.value, 3), np.float32)
        pscale = np.empty(n.value, np.float32)
        level = np.empty(n.value, np.int32)
        if n.value:
            self._check(self._L.avs_get_octree_points(self._h, C.byref(n), pos.ctypes.data, pscale.ctypes.data, level.ctypes.data),
                        "avs_get_octree_points")
        return pos, pscale, level

    def solve_resident(self, params: Params) -> SolveInfo:
        p = params.to_c()
        r = _new_result()
        self._check(self._L.avs_solve_resident(self._h, C.byref(p), C.byref(r)), "avs_solve_resident")
        return SolveInfo.from_c(r)

    def apply(self, out: List) -> SolveInfo:
        o = AvsVelocityOut()
        dev = 0
        for a in range(3):
            if isinstance(out[a], np.ndarray):
                assert out[a].dtype == np.float32 and out[a].flags.c_contiguous
                o.vel[a] = out[a].ctypes.data
            else:
                o.vel[a] = out[a].data_ptr()
                dev = int(out[a].is_cuda)
        o.on_device = dev
        r = _new_result()
        self._check(self._L.avs_apply(self._h, C.byref(o), C.byref(r)), "avs_apply")
        return SolveInfo.from_c(r)

    # ---- read-back (host numpy)
    def sizes(self):
        n, nnz, lv = C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self._L.avs_get_sizes(self._h, C.byref(n), C.byref(nnz), C.byref(lv)), "avs_get_sizes")
        return n.value, nnz.value, lv.value

    def local_range(self):
        b, e = C.c_int64(), C.c_int64()
        self._check(self._L.avs_get_local_range(self._h, C.byref(b), C.byref(e)), "avs_get_local_range")
        return b.value, e.value

    def row_starts(self, nranks: int) -> list:
        a = (C.c_int64 * (nranks + 1))()
        self._check(self._L.avs_get_row_starts(self._h, a), "avs_get_row_starts")
        return list(a)

    def keys(self) -> np.ndarray:
        n, _, _ = self.sizes()
        k = np.empty((n, 5), np.int32)
        self._check(self._L.avs_get_keys(self._h, k.ctypes.data), "avs_get_keys")
        return k

    def system(self):
        """(row_ptr, col, val, rhs, x0) of the resident system (canonical CSR, columns sorted)."""
        n, nnz, _ = self.sizes()
        b, e = self.local_range()
        ptr = np.empty(e - b + 1, np.int64)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        rhs = np.empty(e - b, np.float64)
        x0 = np.empty(n, np.float64)
        self._check(self._L.avs_get_system_csr(self._h, ptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                                               rhs.ctypes.data, x0.ctypes.data), "avs_get_system_csr")
        return ptr, col, val, rhs, x0

    def output_slab(self, axis: int):
        """z-planes [z0, z1) of velocity component ``axis`` this rank computes (and, with distributed output, fills)."""
        z0, z1 = C.c_int32(), C.c_int32()
        self._check(self._L.avs_get_output_slab(self._h, int(axis), C.byref(z0), C.byref(z1)), "avs_get_output_slab")
        return z0.value, z1.value

    def solution(self) -> np.ndarray:
        b, e = self.local_range()
        x = np.empty(e - b, np.float64)
        self._check(self._L.avs_get_solution(self._h, x.ctypes.data), "avs_get_solution")
        return x

    _GRID_DTYPES = {0: np.uint8, 1: np.int32, 2: np.int8, 3: np.int8, 4: np.int8, 5: np.float32, 6: np.float32}

    def grid(self, kind: int, level: int = 0, axis: int = 0) -> np.ndarray:
        res = (C.c_int32 * 3)()
        nb = C.c_int64()
        self._check(self._L.avs_get_grid(self._h, kind, level, axis, None, res, C.byref(nb)), "avs_get_grid")
        out = np.empty(res[0] * res[1] * res[2], self._GRID_DTYPES[kind])
        assert out.nbytes == nb.value
        self._check(self._L.avs_get_grid(self._h, kind, level, axis, out.ctypes.data, res, C.byref(nb)), "avs_get_grid")
        return out.reshape(res[2], res[1], res[0])

    def labels(self, level):
        return self.grid(0, level)

    def face_labels(self, level, axis):
        return self.grid(1, level, axis)

    def edge_labels(self, level, axis):
        return self.grid(2, level, axis)

    def center_labels(self, level):
        return self.grid(3, level)

    def regular_labels(self, axis):
        return self.grid(4, 0, axis)

    def center_weights(self):
        return self.grid(5)

    def edge_weights(self, axis):
        return self.grid(6, 0, axis)

    # ---- stand-alone linear algebra
    def cg_csr(self, ptr, col, val, rhs, x0, params: Params):
        ptr = np.ascontiguousarray(ptr, np.int64)
        col = np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        rhs = np.ascontiguousarray(rhs, np.float64)
        x = np.array(x0, np.float64, copy=True)
        p = params.to_c()
        r = _new_result()
        rc = self._L.avs_cg_csr(self._h, ptr.size - 1, ptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                                rhs.ctypes.data, x.ctypes.data, C.byref(p), C.byref(r))
        self._check(rc, "avs_cg_csr")
        return x, SolveInfo.from_c(r)

    def spmv_csr(self, ptr, col, val, x, single_precision=False, repeats=0):
        ptr = np.ascontiguousarray(ptr, np.int64)
        col = np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty_like(x)
        ms = C.c_float(0)
        rc = self._L.avs_spmv_csr(self._h, ptr.size - 1, ptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                                  x.ctypes.data, y.ctypes.data, int(single_precision), repeats, C.byref(ms))
        self._check(rc, "avs_spmv_csr")
        return y, ms.value

    def time_spmv_resident(self, repeats=20):
        ms = C.c_float(0)
        nbytes = C.c_double(0)
        self._check(self._L.avs_time_spmv_resident(self._h, 0, repeats, C.byref(ms), C.byref(nbytes)), "avs_time_spmv_resident")
        return ms.value, nbytes.value


# -------------------------------------------------------------------------------------------------
# The reference-facing operator
# -------------------------------------------------------------------------------------------------
class SIM_Object:
    """Stand-in for Houdini's SIM_Object: named fields the sub-solver looks up (AV.cpp:138-231)."""

    def __init__(self, scalar_fields: Dict[str, Scene] = None, **fields):
        self.fields = dict(fields)
        self.geometry: Dict[str, tuple] = {}     # getOrCreateGeometry(obj, "octreeGeometry") (AV.cpp:285)

    @staticmethod
    def from_scene(scene: Scene) -> "SIM_Object":
        o = SIM_Object(surface=scene.surface, vel=scene.vel, collision=scene.collision, collisionvel=scene.collision_vel,
                       surfaceweights=scene.face_weights, viscosity=scene.viscosity, massdensity=scene.density)
        o.res, o.origin, o.dx = scene.res, scene.origin, scene.dx
        return o


class MultiSolver:
    """``AvsMulti``: ONE host thread drives the row-partitioned solve on several GPUs (``avs_create_multi`` /
    ``avs_solve_multi``, include/avs.h) -- the shape of the DOP, whose ``solveGasSubclass`` runs on Houdini's cook thread
    (HDK_AdaptiveViscosity.cpp:126-128).  ``devices`` may repeat an ordinal: the ranks then share that GPU (how the multi-rank
    path is exercised on a one-GPU box).  Fields and ``out`` must be host (numpy) arrays."""

    def __init__(self, devices: Sequence[int], time_spmv: bool = False):
        self._L = _lib.load()
        n = len(devices)
        arr = (C.c_int32 * n)(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self._L.avs_create_multi(arr, n, int(time_spmv), C.byref(h))
        if rc != AVS_OK:
            raise AvsError(rc, "avs_create_multi", _lib.last_error())
        self._h = h
        self.nranks = n

    def close(self):
        if getattr(self, "_h", None):
            self._L.avs_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rank(self, r: int) -> "Solver":
        """Read-back view of rank ``r``'s context (keys, local system, local solution, label grids)."""
        v = Solver.__new__(Solver)
        v._L = self._L
        v._h = C.c_void_p(self._L.avs_multi_context(self._h, int(r)))
        v._owned = False
        return v

    def solve(self, scene: Scene, params: Params, out: Optional[List] = None) -> SolveInfo:
        keep: list = []
        f = fields_to_c(scene, keep)
        p = params.to_c()
        r = _new_result()
        o = None
        if out is not None:
            o = AvsVelocityOut()
            for a in range(3):
                assert isinstance(out[a], np.ndarray) and out[a].dtype == np.float32 and out[a].flags.c_contiguous
                o.vel[a] = out[a].ctypes.data
            o.on_device = 0
        rc = self._L.avs_solve_multi(self._h, C.byref(f), C.byref(p), C.byref(o) if o is not None else None, C.byref(r))
        if rc != AVS_OK:
            raise AvsError(rc, "avs_solve_multi", _lib.last_error())
        return SolveInfo.from_c(r)

    def solution(self) -> np.ndarray:
        """The full solution vector, concatenated over the ranks' row blocks."""
        return np.concatenate([self.rank(r).solution() for r in range(self.nranks)])

    def keys(self) -> np.ndarray:
        return self.rank(0).keys()


class HDK_AdaptiveViscosity:
    """Same option names as the DOP (HDK_AdaptiveViscosity.h:28-41, HDK_AdaptiveViscosity.cpp:36-116).

    ``fineBandwidth`` and ``doApplySolidWeights`` are the names the reference *reads*; the DOP exposes
    ``fineLayerBandwidth`` / ``applySolidWeights`` instead, so in a real scene they stay at 0 / False
    (SURVEY.md section 5).  The same defaults apply here.
    """

    FIELD_NAMES = dict(surface="surface", faceWeights="surfaceweights", velocity="vel", viscosity="viscosity",
                       density="massdensity", collision="collision", collisionvel="collisionvel")

    def __init__(self, tolerance=1e-3, maxIterations=2500, numberSuperSamples=3, octreeLevels=4, fineBandwidth=0,
                 useEnhancedGradients=True, doApplySolidWeights=False, doPrintOctree=False, onlyPrintOctree=False,
                 extrapolation=0.5, singlePrecision=False, device=0, **field_names):
        self.tolerance, self.maxIterations, self.numberSuperSamples = tolerance, int(maxIterations), numberSuperSamples
        self.octreeLevels, self.fineBandwidth = octreeLevels, fineBandwidth
        self.useEnhancedGradients, self.doApplySolidWeights = useEnhancedGradients, doApplySolidWeights
        self.doPrintOctree, self.onlyPrintOctree = doPrintOctree, onlyPrintOctree
        self.extrapolation, self.singlePrecision = extrapolation, singlePrecision
        self.names = dict(self.FIELD_NAMES, **field_names)
        self.errors: List[str] = []
        self.extra_info = ""
        self.info: Optional[SolveInfo] = None
        self._solver: Optional[Solver] = None
        self._device = device

    # addError(obj, SIM_MESSAGE, text, UT_ERROR_WARNING)
    def addError(self, obj, text):
        self.errors.append(text)

    def _params(self, timestep) -> Params:
        return Params(dt=float(timestep), tolerance=self.tolerance, max_iterations=self.maxIterations,
                      number_super_samples=self.numberSuperSamples, octree_levels=self.octreeLevels,
                      fine_bandwidth=self.fineBandwidth, use_enhanced_gradients=self.useEnhancedGradients,
                      do_apply_solid_weights=self.doApplySolidWeights, extrapolation=self.extrapolation,
                      single_precision=self.singlePrecision)

    def solveGasSubclass(self, engine, obj: SIM_Object, time, timestep) -> bool:
        """HDK_AdaptiveViscosity::solveGasSubclass (AV.cpp:126-710). Updates obj's ``vel`` in place."""
        self.errors.clear()
        f = obj.fields
        vel = f.get(self.names["velocity"])
        fw = f.get(self.names["faceWeights"])
        if vel is None:
            self.addError(obj, "Liquid velocity field missing")                      # AV.cpp:154
            return False
        if not _is_face_sampled(vel, obj):
            self.addError(obj, "Liquid velocity field must be a staggered grid")     # AV.cpp:159
            return False
        if fw is None:
            self.addError(obj, "Face weights field missing")                         # AV.cpp:165
            return False
        if not _aligned_vec(fw, vel):
            self.addError(obj, "Face weights must align with velocity samples")      # AV.cpp:171
            return False
        coll = f.get(self.names["collision"])
        if coll is None:
            self.addError(obj, "Solid surface field missing")                        # AV.cpp:177
            return False
        cvel = f.get(self.names["collisionvel"])
        if cvel is None:
            self.addError(obj, "Solid velocity field missing")                       # AV.cpp:185
            return False
        surf = f.get(self.names["surface"])
        if surf is None:
            self.addError(obj, "Liquid surface field is missing")                    # AV.cpp:191
            return False
        visc = f.get(self.names["viscosity"])
        if visc is None:
            self.addError(obj, "Viscosity field is missing")                         # AV.cpp:207
            return False
        if not _aligned(visc, surf):
            self.addError(obj, "Viscosity field must align with the surface volume")  # AV.cpp:212
            return False
        dens = f.get(self.names["density"])
        if dens is None:
            self.addError(obj, "Density field is missing")                           # AV.cpp:222
            return False
        if not _aligned(dens, surf):
            self.addError(obj, "Density field must align with the surface volume")   # AV.cpp:227
            return False

        scene = Scene(obj.res, obj.origin, obj.dx, surf, vel, fw, visc, dens, coll, cvel)
        if self._solver is None:
            try:
                self._solver = Solver(device=self._device)
            except AvsError as e:      # no GPU / no library: reported like the C++ shim does (addError, UT_ERROR_ABORT) -- no CPU path
                self.addError(obj, str(e))
                return False
        if self.doPrintOctree and self.onlyPrintOctree:
            # AV.cpp:283-294: dump the octree geometry ("octreeGeometry" SIM_GeometryCopy: P, pscale, octreeLevel) and return
            try:
                self.info = self._solver.build_octree(scene, self._params(timestep))
                obj.geometry["octreeGeometry"] = self._solver.octree_points()
            except AvsError as e:
                self.addError(obj, str(e))
                return False
            return True
        out = [v.data for v in vel]
        try:
            self.info = self._solver.solve(scene, self._params(timestep), out)
        except AvsError as e:
            self.addError(obj, str(e))
            return False
        if self.doPrintOctree:
            obj.geometry["octreeGeometry"] = self._solver.octree_points()
        i = self.info
        # event.setExtraInfo (AV.cpp:645-652)
        self.extra_info = "iterations=%d, error=%.6f, octree DOFS=%d, regular DOFs=%d" % (
            i.iterations, i.error, i.octree_dofs, i.regular_dofs)
        return True


def _is_face_sampled(vel, obj) -> bool:
    if not isinstance(vel, (list, tuple)) or len(vel) != 3:
        return False
    for a, c in enumerate(vel):
        if c.data is None:
            return False
        nz, ny, nx = c.data.shape
        want = list(obj.res)
        want[a] += 1
        if (nx, ny, nz) != tuple(want):
            return False
    return True


def _aligned(a: SampledField, b: SampledField) -> bool:
    if a.data is None or b.data is None:
        return True
    return tuple(a.data.shape) == tuple(b.data.shape) and np.allclose(a.org, b.org, rtol=0, atol=1e-6 * b.dx) \
        and abs(a.dx - b.dx) <= 1e-9 * b.dx


def _aligned_vec(a, b) -> bool:
    return isinstance(a, (list, tuple)) and len(a) == 3 and all(_aligned(x, y) for x, y in zip(a, b))
