mkdir -p gpurun_out
for v in 0 1; do AVS_PCG_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1m_bench_c3_v$v.json 2> gpurun_out/r1m_bench_c3_v$v.err; done
AVS_PCG_VARIANT=1 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1m_bench_c2_v1.json 2> gpurun_out/r1m_bench_c2_v1.err
AVS_PCG_VARIANT=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
