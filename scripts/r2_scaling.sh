# Round 2: strong scaling on N GPUs of one box (N = number of GPUs of the gpurun call):
#   gpurun --gpus N --timeout 900 -- 'bash scripts/r2_scaling.sh N'
# C3 (default bench workload, with the end-to-end leg), C3 with round 1's kernel, C4 (BASELINE configs[3], ~40 M DOF), and at 8 GPUs C5.
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N "$@" > gpurun_out/r2s_${name}_g$N.json 2> gpurun_out/r2s_${name}_g$N.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2s_${name}_g$N.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name g$N", "ms/step %.1f value %.3e"%(j["ms_per_step"], j["value"]), "it", c.get("iterations"), "err", c.get("rel_error"), "spmv", sp.get("avg_ms"), "xr", r.get("xr_phase_ms_per_iter"), "p", r.get("p_phase_ms_per_iter"), "e2e", (j.get("e2e") or {}).get("ms_per_step"))
    print("   stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2s_${name}_g$N.err").read()[-2500:])
PY
}
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_inprocess_multi.py tests/test_gpu_hdk_shim.py -q --tb=short 2>&1 | tail -12 > gpurun_out/r2s_pytest_multi_g2.log; tail -4 gpurun_out/r2s_pytest_multi_g2.log
fi
run c3 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
if [ "$N" = "2" ]; then AVS_PCG_KERNEL=v1 run c3_v1 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e; fi
run c4 --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
if [ "$N" = "2" ]; then AVS_PCG_KERNEL=v1 run c4_v1 --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e; fi
if [ "$N" = "8" ]; then run c5 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline; run c3_gather --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gather-output; fi
