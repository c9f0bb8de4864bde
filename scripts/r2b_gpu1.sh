# Round 2, call B (1 GPU): the new persistent kernel (k_cg_persistent2) and the SpMV slice-loop variants.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 > gpurun_out/r2b_pytest_gpu.log; tail -3 gpurun_out/r2b_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2b_$name.json")); r=j["roofline"]; c=j["config"]
    print("$name", "ms/step %.1f solve %.1f it %d"%(j["ms_per_step"], c["stage_ms"]["solve"], c["iterations"]), "cg frac %.3f"%r["frac"], "spmv %.4f (%.3f) standalone %.4f (%.3f) xr %.4f p %.4f"%(r["spmv_phase"]["avg_ms"], r["spmv_phase"]["frac"], r["spmv_standalone"]["avg_ms"], r["spmv_standalone"]["frac"], r["xr_phase_ms_per_iter"], r["p_phase_ms_per_iter"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2b_$name.err").read()[-1500:])
PY
}
run c3_v2_base python bench.py --workload c3 $B
AVS_PCG_KERNEL=v1 run c3_v1 python bench.py --workload c3 $B
AVS_SPMV_MODE=pf run c3_v2_pf python bench.py --workload c3 $B
AVS_SPMV_MODE=ring run c3_v2_ring python bench.py --workload c3 $B
AVS_SPMV_MODE=ring4 run c3_v2_ring4 python bench.py --workload c3 $B
AVS_SPMV_MODE=ring run c3_fp32_ring python bench.py --workload c3 --fp32 $B
AVS_SPMV_MODE=pf run c3_fp32_pf python bench.py --workload c3 --fp32 $B
run c2_v2_base python bench.py --workload c2 $B
AVS_PCG_KERNEL=v1 run c2_v1 python bench.py --workload c2 $B
AVS_SPMV_MODE=ring run c2_v2_ring python bench.py --workload c2 $B
