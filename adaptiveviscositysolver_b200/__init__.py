"""B200-native adaptive-octree viscosity solve (one hot path of rgoldade/AdaptiveViscositySolver).

``scenes``  caller-side field containers + synthetic scenes (numpy only)
``solver``  host mirror of the reference operator (``HDK_AdaptiveViscosity.solveGasSubclass``) and the
            staged ``Solver`` wrapper over the C-ABI in ``include/avs.h``
``_lib``    ctypes binding of ``libavs_b200.so`` -- hand-written sm_100a CUDA, no CPU fallback
"""
from .scenes import SampledField, Scene, sphere_drop  # noqa: F401


def __getattr__(name):  # lazy: importing the package must work on a box without the built library
    if name in ("Solver", "MultiSolver", "Params", "HDK_AdaptiveViscosity", "SIM_Object", "SolveInfo"):
        from . import solver
        return getattr(solver, name)
    raise AttributeError(name)
