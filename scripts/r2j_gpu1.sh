# Round 2, call J (1 GPU): full GPU suite after the reverts (thread-per-row assembly, one row per CTA) and the p-pointer fix; v1 vs v2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2j_pytest_gpu.log; tail -8 gpurun_out/r2j_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2j_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "spmv", sp.get("avg_ms"), "xr", r.get("xr_phase_ms_per_iter"), "p", r.get("p_phase_ms_per_iter"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2j_$name.err").read()[-1500:])
PY
}
run c3 python bench.py --workload c3 $B
AVS_PCG_KERNEL=v2 run c3_v2 python bench.py --workload c3 $B
AVS_PCG_KERNEL=v2 AVS_SPMV_MODE=pf run c3_v2_pf python bench.py --workload c3 $B
AVS_PCG_KERNEL=v2 run c3_fp32_v2 python bench.py --workload c3 --fp32 $B
run c3_fp32 python bench.py --workload c3 --fp32 $B
AVS_PCG_KERNEL=v2 run c2_v2 python bench.py --workload c2 $B
run c2 python bench.py --workload c2 $B
