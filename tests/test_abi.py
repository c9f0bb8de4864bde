"""CPU-side checks of the boundary: the library loads and exports every symbol include/avs.h declares,
struct sizes agree between the header and the ctypes mirror, and there is no silent CPU fallback."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from adaptiveviscositysolver_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(built):
    header = (ROOT / "include" / "avs.h").read_text()
    declared = set(re.findall(r"^(?:int|void|const char \*|AvsContext \*)\s*\*?\s*(avs_\w+)\s*\(", header, re.M))
    assert len(declared) >= 18
    L = built.load()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(built.EXPORTS)
    assert L.avs_abi_version() == 3


def test_struct_layout_matches_header(built, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "avs.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(AvsField),'
                   'sizeof(AvsFields), sizeof(AvsParams), sizeof(AvsVelocityOut), sizeof(AvsResult), sizeof(AvsDeviceConfig));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mirror = [C.sizeof(t) for t in (built.AvsField, built.AvsFields, built.AvsParams, built.AvsVelocityOut,
                                   built.AvsResult, built.AvsDeviceConfig)]
    assert sizes == mirror


def test_defaults_are_the_reference_defaults(built):
    p = built.AvsParams()
    built.load().avs_default_params(C.byref(p))
    assert p.size == C.sizeof(built.AvsParams)
    assert (p.tolerance, p.max_iterations, p.number_super_samples, p.octree_levels) == (1e-3, 2500, 3, 4)
    assert (p.fine_bandwidth, p.use_enhanced_gradients, p.do_apply_solid_weights, p.extrapolation) == (0, 1, 0, 0.5)
    assert built.status_string(0) == "ok" and "cancel" in built.status_string(-7)


def test_no_cpu_fallback_without_gpu(built):
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from adaptiveviscositysolver_b200.solver import Solver
    with pytest.raises(built.AvsError) as e:
        Solver(device=0)
    assert e.value.status == -9


def test_product_never_imports_the_oracle():
    for p in (ROOT / "adaptiveviscositysolver_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            assert "oracle" not in p.read_text().replace("CPU oracle", "").replace("the oracle", "").lower() or \
                "import" not in "".join(l for l in p.read_text().splitlines() if "oracle" in l.lower()), p
