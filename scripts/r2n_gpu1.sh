# Round 2, call N (1 GPU): tile-culled level-0 labelling + node pyramid, hashed accumulator in the split assembly; ncu full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2n_pytest_gpu.log; tail -12 gpurun_out/r2n_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2n_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "spmv", sp.get("avg_ms"), "frac", r.get("frac"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2n_$name.err").read()[-1500:])
PY
}
AVS_TRACE=1 run c3 python bench.py --workload c3 $B
grep -m2 "rows through" gpurun_out/r2n_c3.err
AVS_LABELS=dense run c3_dense python bench.py --workload c3 $B
run c2 python bench.py --workload c2 $B
run c5 python bench.py --workload c5 $B
run c4 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_assemble<|k_apply_regular|k_weights_classify4" -c 5 -o gpurun_out/r2n_stages python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-spmv-events > gpurun_out/r2n_ncu_stages.log 2>&1
tail -3 gpurun_out/r2n_ncu_stages.log | cut -c1-300
