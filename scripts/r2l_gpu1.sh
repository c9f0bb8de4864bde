# Round 2, call L (1 GPU): split assembly, 4-samples-per-thread labelling, fused weight classification, parallel level lookup in the
# interpolator, z-windowed node pyramid / own-row restriction (multi-rank paths run as 2 ranks on the one GPU in the test-suite).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -30 > gpurun_out/r2l_pytest_gpu.log; tail -8 gpurun_out/r2l_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2l_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "spmv", sp.get("avg_ms"), "xr", r.get("xr_phase_ms_per_iter"), "p", r.get("p_phase_ms_per_iter"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2l_$name.err").read()[-1500:])
PY
}
run c3 python bench.py --workload c3 $B
AVS_ASM=generic run c3_asm_generic python bench.py --workload c3 $B
AVS_LABELS=1 run c3_labels1 python bench.py --workload c3 $B
AVS_PCG_KERNEL=v2 AVS_SPMV_MODE=pf run c3_fp32_v2_pf python bench.py --workload c3 --fp32 $B
run c2 python bench.py --workload c2 $B
run c4 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
