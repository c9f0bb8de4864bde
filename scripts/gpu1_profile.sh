mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in c3 c2 c5; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r1h_bench_$w.json 2> gpurun_out/r1h_bench_$w.err; done
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1h_bench_c3_reference.json 2> gpurun_out/r1h_bench_c3_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1h_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1h_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -c 1 -o gpurun_out/r1h_pcg python bench.py --workload c3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r1h_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
