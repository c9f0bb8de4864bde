# Round 2, call U (1 GPU): validation of the latest labelling changes -- full suite + C3 + C4.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2u_pytest_gpu.log; tail -6 gpurun_out/r2u_pytest_gpu.log
run() { name=$1; shift; "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2u_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "err", c.get("rel_error"), "spmv", sp.get("avg_ms"), "frac", r.get("frac"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2u_$name.err").read()[-1500:])
PY
}
run c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
run c4 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
