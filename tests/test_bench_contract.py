"""bench.py's JSON contract, checked where it can be checked without a GPU: the reference arm (the CPU restatement timed on the
host cores) prints ONE JSON line with the contract's keys, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=e)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0", "--gpus", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "viscosity_solve_dof_iters_per_s" and d["unit"] == "DOF*iters/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("C1") and d["config"]["N"] > 10000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ws = d["config"]["whole_solve"]
    assert ws["iterations"] > 0 and ws["rel_error"] < 1e-3 and 0 < ws["dof_iters_per_s"] < d["value"] * 1.5
    # the compiled reference timed beside the port on one whole C2-literal solve (only where oracle/_ref is present)
    cr = d["config"]["compiled_reference"]
    if "unavailable" not in cr:
        assert cr["reference"]["iterations"] == cr["port"]["iterations"] == 161 and cr["N"] == 495108
        assert cr["reference"]["whole_solve_s"] > 0 and cr["port"]["whole_solve_s"] > 0 and cr["port_over_reference"] > 0


def test_reference_arm_only_rank0_works_under_torchrun_env():
    r = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--workload", "c1", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "needs a CUDA device" in (r.stderr + r.stdout)
