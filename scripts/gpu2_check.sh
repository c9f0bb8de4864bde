mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30 > gpurun_out/r1k_pytest_multi.log
tail -5 gpurun_out/r1k_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r1k_bench_c3_g2.json 2> gpurun_out/r1k_bench_c3_g2.err
AVS_PCG_INTERIOR_FIRST=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r1k_bench_c3_g2_nointerior.json 2> gpurun_out/r1k_bench_c3_g2_nointerior.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --workload c2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r1k_bench_c2_g2.json 2> gpurun_out/r1k_bench_c2_g2.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1k_bench_c3_g1.json 2> gpurun_out/r1k_bench_c3_g1.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
