# Round 2, ncu full capture of the two assembly kernels on the final build (C3).
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_assemble -c 2 -o gpurun_out/r2zz_asm python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-spmv-events > gpurun_out/r2zz_ncu_asm.log 2>&1; tail -2 gpurun_out/r2zz_ncu_asm.log | cut -c1-200
