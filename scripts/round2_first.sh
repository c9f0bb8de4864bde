# Round 2, first GPU call (1 GPU): time the experimental persistent kernel against the default one, then racecheck a small solve.
#   gpurun --timeout 900 -- 'bash scripts/round2_first.sh'
mkdir -p gpurun_out
AVS_PCG_KERNEL=x python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_x.log; tail -2 gpurun_out/r2_pytest_gpu_x.log
for w in c3 c2; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_${w}_default.json 2> gpurun_out/r2_bench_${w}_default.err
  AVS_PCG_KERNEL=x python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_${w}_x.json 2> gpurun_out/r2_bench_${w}_x.err
done
AVS_ASM_ROW=hash python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu_hash.log; tail -2 gpurun_out/r2_pytest_gpu_hash.log
AVS_ASM_ROW=hash python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_c3_hashrow.json 2> gpurun_out/r2_bench_c3_hashrow.err
python bench.py --workload c4 --steps 2 --warmup 2 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
# memcheck + racecheck of one small solve through the C-ABI (sanitizers in CI on the small config)
cat > /tmp/avs_small.py <<'PY'
import sys; sys.path.insert(0, '.')
from adaptiveviscositysolver_b200 import Params, Solver, sphere_drop
sc = sphere_drop(32, 10); s = Solver(device=0)
out = [v.data.copy() for v in sc.vel]
print(s.solve(sc, Params(octree_levels=4, tolerance=1e-6), out).iterations)
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/avs_small.py > gpurun_out/r2_memcheck.log 2>&1; tail -3 gpurun_out/r2_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python /tmp/avs_small.py > gpurun_out/r2_racecheck.log 2>&1; tail -3 gpurun_out/r2_racecheck.log
