# Round 2, call A (1 GPU): measure what round 1 left unmeasured.
#   gpurun --timeout 1500 -- 'bash scripts/r2a_gpu1.sh'
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; "$@" > gpurun_out/r2a_$name.json 2> gpurun_out/r2a_$name.err; tail -c 300 gpurun_out/r2a_$name.json | head -c 300; echo; }
run c3_default python bench.py --workload c3 $B
AVS_PCG_KERNEL=x run c3_x python bench.py --workload c3 $B
run c3_fp32 python bench.py --workload c3 --fp32 $B
AVS_PCG_KERNEL=x run c3_fp32_x python bench.py --workload c3 --fp32 $B
AVS_ASM_ROW=hash run c3_hashrow python bench.py --workload c3 $B
run c4_g1 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
AVS_PCG_KERNEL=x python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/r2a_pytest_gpu_x.log; tail -2 gpurun_out/r2a_pytest_gpu_x.log
AVS_ASM_ROW=hash python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/r2a_pytest_gpu_hash.log; tail -2 gpurun_out/r2a_pytest_gpu_hash.log
# fp32 ncu captures (BASELINE configs[2]): stand-alone SpMV and the persistent CG kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sjds --launch-skip 3 -c 1 -o gpurun_out/r2a_spmv_fp32 python bench.py --workload c3 --fp32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu_spmv_fp32.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -c 1 -o gpurun_out/r2a_pcg_fp32 python bench.py --workload c3 --fp32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu_pcg_fp32.log 2>&1
# sanitizers on one small solve through the C-ABI
cat > /tmp/avs_small.py <<'PY'
import sys; sys.path.insert(0, '.')
from adaptiveviscositysolver_b200 import Params, Solver, sphere_drop
sc = sphere_drop(32, 10); s = Solver(device=0)
out = [v.data.copy() for v in sc.vel]
print(s.solve(sc, Params(octree_levels=4, tolerance=1e-6), out).iterations)
PY
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool python /tmp/avs_small.py > gpurun_out/r2a_$tool.log 2>&1; tail -2 gpurun_out/r2a_$tool.log
done
AVS_CG_MODE=launch timeout 300 compute-sanitizer --tool memcheck python /tmp/avs_small.py > gpurun_out/r2a_memcheck_launch.log 2>&1; tail -2 gpurun_out/r2a_memcheck_launch.log
ls -la gpurun_out | grep r2a
