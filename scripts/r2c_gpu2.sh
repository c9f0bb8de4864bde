# Round 2, call C (N GPUs, default 2): multi-GPU parity tests + k_cg_persistent2 against round 1's kernel, C3 and C4.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -25 > gpurun_out/r2c_pytest_multi.log; tail -4 gpurun_out/r2c_pytest_multi.log
fi
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N "$@" > gpurun_out/r2c_${name}_g$N.json 2> gpurun_out/r2c_${name}_g$N.err; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2c_${name}_g$N.json")); r=j["roofline"]; c=j["config"]
    print("$name g$N", "ms/step %.1f solve %.1f it %d err %.6e"%(j["ms_per_step"], c["stage_ms"]["solve"], c["iterations"], c["rel_error"]), "spmv %.4f xr %.4f p %.4f"%(r["spmv_phase"]["avg_ms"], r["xr_phase_ms_per_iter"], r["p_phase_ms_per_iter"]), "e2e", (j.get("e2e") or {}).get("ms_per_step"))
    print("   stages", c["stage_ms"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2c_${name}_g$N.err").read()[-2500:])
PY
}
run c3_v2 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
AVS_PCG_KERNEL=v1 run c3_v1 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
run c4_v2 --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
AVS_PCG_KERNEL=v1 run c4_v1 --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
