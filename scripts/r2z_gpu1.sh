# Round 2, last call (1 GPU, final build): smoke, the default bench line, launch list, sanitizers, fp32 ncu capture (in this order: the
# call is clamped to the GPU budget that is left).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; tail -c 1200 gpurun_out/r2z_bench_default.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2z_ncu_launches.log 2>&1
cat > /tmp/avs_small.py <<'PY'
import sys; sys.path.insert(0, '.')
from adaptiveviscositysolver_b200 import Params, Solver, sphere_drop
sc = sphere_drop(32, 10); s = Solver(device=0)
out = [v.data.copy() for v in sc.vel]
print(s.solve(sc, Params(octree_levels=4, tolerance=1e-6), out).iterations)
PY
for tool in memcheck synccheck racecheck; do
  timeout 120 compute-sanitizer --tool $tool python /tmp/avs_small.py > gpurun_out/r2z_$tool.log 2>&1; echo "$tool: $(tail -1 gpurun_out/r2z_$tool.log)"
done
AVS_PCG_KERNEL=v2 timeout 100 compute-sanitizer --tool memcheck python /tmp/avs_small.py > gpurun_out/r2z_memcheck_v2.log 2>&1; echo "memcheck v2 kernel: $(tail -1 gpurun_out/r2z_memcheck_v2.log)"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent2 -c 1 -o gpurun_out/r2z_pcg2_fp32 python bench.py --workload c3 --fp32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r2z_ncu_pcg2_fp32.log 2>&1; tail -1 gpurun_out/r2z_ncu_pcg2_fp32.log | cut -c1-200
