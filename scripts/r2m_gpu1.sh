# Round 2, call M (1 GPU): full GPU suite after the restriction fix; fp32 default (v2 kernel + prefetching slice loop); ncu launch list
# of the C3 step and full captures of the assembly / write-back kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2m_pytest_gpu.log; tail -12 gpurun_out/r2m_pytest_gpu.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err; python - <<PY
import json
try:
    txt=open("gpurun_out/r2m_$name.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); r=j["roofline"]; c=j["config"]
    sp=r.get("spmv_phase") or {}
    print("$name", "ms/step %.1f"%j["ms_per_step"], "it", c.get("iterations"), "spmv", sp.get("avg_ms"), "frac", r.get("frac"), "xr", r.get("xr_phase_ms_per_iter"), "p", r.get("p_phase_ms_per_iter"), "stages", c.get("stage_ms"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2m_$name.err").read()[-1500:])
PY
}
run c3_fp32 python bench.py --workload c3 --fp32 $B
run c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
AVS_FACEW_UPLOAD=bulk run c3_facew_bulk python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
AVS_ASM_ROW=hash run c3_hashrow python bench.py --workload c3 $B
python - <<'PY'
import json
for n in ("c3", "c3_facew_bulk"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/r2m_{n}.json").read().splitlines() if l.startswith("{")][-1]); e = j["e2e"]
        print(n, "e2e ms", e["ms_per_step"], "h2d", e["h2d_bytes_per_step"], "d2h", e["d2h_bytes_per_step"], e.get("stage_ms"))
    except Exception as ex:
        print(n, "e2e FAILED", ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2m_ncu_launches.log 2>&1
tail -2 gpurun_out/r2m_ncu_launches.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_assemble|k_apply_regular|k_classify_faces|k_node_sample" -c 12 -o gpurun_out/r2m_stages python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-spmv-events > gpurun_out/r2m_ncu_stages.log 2>&1
tail -3 gpurun_out/r2m_ncu_stages.log | cut -c1-300
ls -la gpurun_out/r2m_*
