# Round 2, call E (1 GPU): in-process multi-rank tests (ranks share the GPU), v2 kernel with counter-based local barrier.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inprocess_multi.py tests/test_gpu_cg_modes.py tests/test_gpu_parity.py -q --tb=short -x 2>&1 | tail -25 > gpurun_out/r2e_pytest_gpu.log; tail -6 gpurun_out/r2e_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2e_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    print("$name", "ms/step %.1f"%j["ms_per_step"], "avg_launch_ms", r.get("avg_launch_ms"), (c.get("stage_ms") or {}).get("solve"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2e_$name.err").read()[-1500:])
PY
}
run c5_v2 python bench.py --workload c5 $B
AVS_PCG_KERNEL=v1 run c5_v1 python bench.py --workload c5 $B
run c2_v2 python bench.py --workload c2 $B
AVS_PCG_KERNEL=v1 run c2_v1 python bench.py --workload c2 $B
run c3_v2 python bench.py --workload c3 $B
