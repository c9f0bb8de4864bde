# Round 2, call F (1 GPU): in-process multi-rank tests (shared GPU) with the stream-ordered allocator, HDK shim run on the stand-ins,
# noinline slice loop + reworked syncSum.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_inprocess_multi.py tests/test_gpu_hdk_shim.py -q --tb=short -x 2>&1 | tail -30 > gpurun_out/r2f_pytest_multi.log; tail -8 gpurun_out/r2f_pytest_multi.log
timeout 300 python -m pytest tests/test_gpu_cg_modes.py tests/test_gpu_parity.py -q --tb=short -x 2>&1 | tail -8 > gpurun_out/r2f_pytest_gpu.log; tail -3 gpurun_out/r2f_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2f_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "solve", (c.get("stage_ms") or {}).get("solve"), "spmv", sp.get("avg_ms"), "xr", r.get("xr_phase_ms_per_iter"), "p", r.get("p_phase_ms_per_iter"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2f_$name.err").read()[-1500:])
PY
}
run c3_v2 python bench.py --workload c3 $B
AVS_SPMV_MODE=inline run c3_v2_inline python bench.py --workload c3 $B
run c2_v2 python bench.py --workload c2 $B
AVS_PCG_KERNEL=v1 run c2_v1 python bench.py --workload c2 $B
run c5_v2 python bench.py --workload c5 $B
run c3_fp32 python bench.py --workload c3 --fp32 $B
