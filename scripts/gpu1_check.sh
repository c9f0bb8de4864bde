mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1g_pytest_gpu.log
tail -4 gpurun_out/r1g_pytest_gpu.log
AVS_SPMV_TMA=1 AVS_CG_MODE=launch python -m pytest tests/test_gpu_parity.py -x -q -k "standalone or solve_parity" 2>&1 | tail -3
AVS_CG_MODE=launch python -m pytest tests/test_gpu_parity.py -x -q -k "standalone or solve_parity or max_iter" 2>&1 | tail -3
for w in c3 c2 c5; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_$w.json 2> gpurun_out/r1g_bench_$w.err; done
AVS_CG_MODE=launch python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1g_bench_c3_launch.json 2> gpurun_out/r1g_bench_c3_launch.err
