"""ctypes binding of ``libavs_b200.so`` (the C-ABI declared in ``include/avs.h``).

There is NO CPU fallback: if the CUDA library has not been built, loading raises; if no GPU is
visible, ``avs_create`` returns ``AVS_ERR_NO_DEVICE`` and the Python wrapper raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libavs_b200.so"

AVS_OK = 0
AVS_PRECISION_F64, AVS_PRECISION_F32 = 0, 1
AVS_STAGE_COUNT = 12
STAGE_NAMES = ["upload", "surface_weights", "octree", "regular_labels", "octree_labels", "restriction",
               "system", "solve", "apply", "download", "total", "_"]

EXPORTS = [
    "avs_abi_version", "avs_device_count", "avs_nccl_unique_id", "avs_get_local_range", "avs_get_row_starts", "avs_get_output_slab", "avs_create", "avs_destroy", "avs_status_string", "avs_last_error", "avs_default_params",
    "avs_solve", "avs_assemble", "avs_solve_resident", "avs_apply", "avs_get_sizes", "avs_get_keys",
    "avs_get_system_csr", "avs_get_solution", "avs_get_grid", "avs_cg_csr", "avs_spmv_csr", "avs_time_spmv_resident",
    "avs_build_octree", "avs_get_octree_points",
    "avs_create_multi", "avs_destroy_multi", "avs_multi_size", "avs_multi_context", "avs_solve_multi",
]


class AvsField(C.Structure):
    _fields_ = [("data", C.c_void_p), ("res", C.c_int32 * 3), ("org", C.c_double * 3), ("dx", C.c_double),
                ("constant", C.c_float), ("on_device", C.c_int32)]


class AvsFields(C.Structure):
    _fields_ = [("size", C.c_uint32), ("res", C.c_int32 * 3), ("origin", C.c_double * 3), ("dx", C.c_double),
                ("surface", AvsField), ("vel", AvsField * 3), ("face_weights", AvsField * 3),
                ("viscosity", AvsField), ("density", AvsField), ("collision", AvsField),
                ("collision_vel", AvsField * 3)]


class AvsParams(C.Structure):
    _fields_ = [("size", C.c_uint32), ("dt", C.c_double), ("tolerance", C.c_double), ("extrapolation", C.c_double),
                ("max_iterations", C.c_int32), ("number_super_samples", C.c_int32), ("octree_levels", C.c_int32),
                ("fine_bandwidth", C.c_int32), ("use_enhanced_gradients", C.c_int32),
                ("do_apply_solid_weights", C.c_int32), ("precision", C.c_int32), ("check_every", C.c_int32),
                ("cancel", C.c_void_p)]


class AvsVelocityOut(C.Structure):
    _fields_ = [("vel", C.c_void_p * 3), ("on_device", C.c_int32)]


class AvsResult(C.Structure):
    _fields_ = [("size", C.c_uint32), ("status", C.c_int32), ("iterations", C.c_int32), ("levels", C.c_int32),
                ("error", C.c_double), ("octree_dofs", C.c_int64), ("regular_dofs", C.c_int64),
                ("edge_dofs", C.c_int64), ("center_dofs", C.c_int64), ("nnz", C.c_int64), ("local_rows", C.c_int64),
                ("spmv_launches", C.c_int64), ("kernel_launches", C.c_int64),
                ("stage_ms", C.c_float * AVS_STAGE_COUNT), ("spmv_ms", C.c_float),
                ("cg_update_xr_ms", C.c_float), ("cg_update_p_ms", C.c_float),
                ("dist_mode", C.c_int32), ("reserved0", C.c_int32), ("halo_columns", C.c_int64),
                ("interpolated_faces", C.c_int64), ("cg_kernel_ms", C.c_float), ("cg_kernel_launches", C.c_int32)]


class AvsDeviceConfig(C.Structure):
    _fields_ = [("size", C.c_uint32), ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("stream", C.c_void_p), ("time_spmv", C.c_int32), ("distributed_output", C.c_int32)]


class AvsError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: {status_string(status)} ({status}){' - ' + detail if detail else ''}")


_lib = None


def load():
    """Load the CUDA library; raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). This package has no CPU or PyTorch fallback path.")
    L = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    L.avs_abi_version.restype = C.c_int
    L.avs_create.argtypes = [C.POINTER(AvsDeviceConfig), C.POINTER(vp)]
    L.avs_destroy.argtypes = [vp]
    L.avs_destroy.restype = None
    L.avs_status_string.restype = C.c_char_p
    L.avs_status_string.argtypes = [C.c_int]
    L.avs_last_error.restype = C.c_char_p
    L.avs_default_params.argtypes = [C.POINTER(AvsParams)]
    L.avs_default_params.restype = None
    L.avs_solve.argtypes = [vp, C.POINTER(AvsFields), C.POINTER(AvsParams), C.POINTER(AvsVelocityOut), C.POINTER(AvsResult)]
    L.avs_assemble.argtypes = [vp, C.POINTER(AvsFields), C.POINTER(AvsParams), C.POINTER(AvsResult)]
    L.avs_solve_resident.argtypes = [vp, C.POINTER(AvsParams), C.POINTER(AvsResult)]
    L.avs_apply.argtypes = [vp, C.POINTER(AvsVelocityOut), C.POINTER(AvsResult)]
    L.avs_get_sizes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.avs_get_keys.argtypes = [vp, vp]
    L.avs_get_local_range.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.avs_nccl_unique_id.argtypes = [vp]
    L.avs_get_row_starts.argtypes = [vp, vp]
    L.avs_get_output_slab.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.avs_get_system_csr.argtypes = [vp, vp, vp, vp, vp, vp]
    L.avs_get_solution.argtypes = [vp, vp]
    L.avs_get_grid.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.avs_cg_csr.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(AvsParams), C.POINTER(AvsResult)]
    L.avs_spmv_csr.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.avs_time_spmv_resident.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]
    L.avs_build_octree.argtypes = [vp, C.POINTER(AvsFields), C.POINTER(AvsParams), C.POINTER(AvsResult)]
    L.avs_get_octree_points.argtypes = [vp, C.POINTER(C.c_int64), vp, vp, vp]
    L.avs_create_multi.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(vp)]
    L.avs_destroy_multi.argtypes = [vp]
    L.avs_destroy_multi.restype = None
    L.avs_multi_size.argtypes = [vp]
    L.avs_multi_context.argtypes = [vp, C.c_int]
    L.avs_multi_context.restype = vp
    L.avs_solve_multi.argtypes = [vp, C.POINTER(AvsFields), C.POINTER(AvsParams), C.POINTER(AvsVelocityOut), C.POINTER(AvsResult)]
    for name in EXPORTS:
        getattr(L, name)  # every symbol of include/avs.h must be exported
    _lib = L
    return L


def status_string(status: int) -> str:
    try:
        return load().avs_status_string(int(status)).decode()
    except Exception:
        return f"status {status}"


def last_error() -> str:
    return load().avs_last_error().decode()
