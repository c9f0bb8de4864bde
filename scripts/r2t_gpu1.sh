# Round 2, call T (1 GPU): work-item restriction of the coarse rows; register budget of the generic assembly pass.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2t_pytest_gpu.log; tail -6 gpurun_out/r2t_pytest_gpu.log
run() { name=$1; shift; "$@" > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2t_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "spmv", sp.get("avg_ms"), "frac", r.get("frac"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2t_$name.err").read()[-1500:])
PY
}
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run c3 python bench.py --workload c3 $B
AVS_ASM_MINB=4 run c3_minb4 python bench.py --workload c3 $B
AVS_ASM_MINB=6 run c3_minb6 python bench.py --workload c3 $B
run c4 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
