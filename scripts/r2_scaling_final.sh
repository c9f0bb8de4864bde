# Round 2, final build: strong scaling on N GPUs, C3 (with the end-to-end leg) and C4 only.
#   gpurun --gpus N --timeout 600 -- 'bash scripts/r2_scaling_final.sh N'
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N "$@" > gpurun_out/r2s_${name}_g$N.json 2> gpurun_out/r2s_${name}_g$N.err; python - <<PY
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
run c3 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
run c4 --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
