# Round 2, call R (1 GPU): full suite after the vectorised octree passes, split coarse restriction, cheaper tile marking; C3, C4, C2, C5.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2r_pytest_gpu.log; tail -6 gpurun_out/r2r_pytest_gpu.log
run() { name=$1; shift; "$@" > gpurun_out/r2r_$name.json 2> gpurun_out/r2r_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2r_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}; e=j.get("e2e") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "spmv", sp.get("avg_ms"), "frac", r.get("frac"), "stages", c.get("stage_ms"))
    if e.get("ms_per_step"): print("   e2e ms", e["ms_per_step"], "h2d", e["h2d_bytes_per_step"], "d2h", e["d2h_bytes_per_step"], e.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2r_$name.err").read()[-1500:])
PY
}
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run c3 python bench.py --workload c3 $B
run c4 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
run c2 python bench.py --workload c2 $B
run c5 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline
