"""Sweep of tests/test_host_assembly.py's check over many seeds of scripts/fuzz_reference_pin.py (CPU; needs /root/reference for
oracle/_ref): the CUDA library's row builder, restriction functions and prolongation, compiled for the host, against the compiled
reference -- matrix, right-hand side, restricted velocity (levels 0-1) and regular-grid output bit for bit.
    python scripts/sweep_host_product_source.py first count
Seeds whose scene is outside the reference's contract (out-of-range columns, see profiles/r2_fuzz.md) are skipped and counted."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import test_host_assembly as t  # noqa: E402
from oracle import avs_ref as ref  # noqa: E402

if __name__ == "__main__":
    first, count = int(sys.argv[1]), int(sys.argv[2])
    L = t.build_harness()
    ok, skipped, bad = 0, [], []
    for seed in range(first, first + count):
        sc, p, desc = t.fz.fuzz_case(seed)
        p.max_iterations = 1
        if "TOUCHING-THE-BOUNDARY" in desc:     # candidates for the scenes the reference itself asserts / crashes on
            skipped.append(seed)
            continue
        R = ref.RefRun(sc, p)
        if R.n_face == 0 or R.csr()[1].min() < 0:
            skipped.append(seed)
            continue
        try:
            t.check_product_rows_against_reference(L, sc, p)
            ok += 1
        except AssertionError as e:
            bad.append(seed)
            print(f"FAIL {desc}\n     {str(e)[:300]}", flush=True)
    print(f"seeds {first}..{first + count - 1}: {ok} agree bit for bit, {len(skipped)} skipped (empty or outside the reference's contract) {skipped}, failing: {bad}")
    sys.exit(1 if bad else 0)
