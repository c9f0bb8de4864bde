# Round 2, call D (1 GPU): full GPU test suite (incl. the tests against the compiled reference) + benches after the fence / halo-flag changes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 > gpurun_out/r2d_pytest_gpu.log; tail -4 gpurun_out/r2d_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2d_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    print("$name", "ms/step %.1f solve %.1f it %d"%(j["ms_per_step"], c["stage_ms"]["solve"], c["iterations"]), "cg frac %.3f"%r["frac"], "spmv %.4f (%.3f) standalone %.4f (%.3f) xr %.4f p %.4f"%(r["spmv_phase"]["avg_ms"], r["spmv_phase"]["frac"], r["spmv_standalone"]["avg_ms"], r["spmv_standalone"]["frac"], r["xr_phase_ms_per_iter"], r["p_phase_ms_per_iter"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2d_$name.err").read()[-1500:])
PY
}
run c3_v2 python bench.py --workload c3 $B
run c2_v2 python bench.py --workload c2 $B
AVS_PCG_KERNEL=v1 run c2_v1 python bench.py --workload c2 $B
run c5_v2 python bench.py --workload c5 $B
AVS_PCG_KERNEL=v1 run c5_v1 python bench.py --workload c5 $B
