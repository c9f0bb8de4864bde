mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30 > gpurun_out/r1n_pytest_multi.log
tail -3 gpurun_out/r1n_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1n_bench_c3_g2.json 2> gpurun_out/r1n_bench_c3_g2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload c5 --steps 2 --warmup 2 > gpurun_out/r1n_bench_c5_g2.json 2> gpurun_out/r1n_bench_c5_g2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r1n_bench_ref_g2.json 2> gpurun_out/r1n_bench_ref_g2.err
