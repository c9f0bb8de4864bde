# Round 2, call P (1 GPU): stress of the 2-ranks-on-one-GPU tests with eager module loading, full suite, C3 with the e2e leg, fp32,
# ncu launch list + full captures of the stage kernels.
mkdir -p gpurun_out
fails=0
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  timeout 300 python -m pytest tests/test_gpu_hdk_shim.py tests/test_gpu_inprocess_multi.py -q --tb=line -x > gpurun_out/r2p_stress_$i.log 2>&1 || fails=$((fails+1))
done
echo "stress: $fails of 12 runs failed"; grep -h "passed\|failed" gpurun_out/r2p_stress_*.log | sort | uniq -c
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2p_pytest_gpu.log; tail -6 gpurun_out/r2p_pytest_gpu.log
run() { name=$1; shift; "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2p_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}; e=j.get("e2e") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "spmv", sp.get("avg_ms"), "frac", r.get("frac"), "stages", c.get("stage_ms"))
    if e.get("ms_per_step"): print("   e2e ms", e["ms_per_step"], "h2d", e["h2d_bytes_per_step"], "d2h", e["d2h_bytes_per_step"], e.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2p_$name.err").read()[-1500:])
PY
}
run c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
run c3_fp32 python bench.py --workload c3 --fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2p_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_assemble<|k_apply_regular|k_weights_classify4" -c 5 -o gpurun_out/r2p_stages python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-spmv-events > gpurun_out/r2p_ncu_stages.log 2>&1
tail -2 gpurun_out/r2p_ncu_stages.log | cut -c1-200
