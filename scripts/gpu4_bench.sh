mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 2 --warmup 2 --no-e2e > gpurun_out/r1o_bench_c3_g4.json 2> gpurun_out/r1o_bench_c3_g4.err
tail -c 300 gpurun_out/r1o_bench_c3_g4.err
