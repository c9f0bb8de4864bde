# Round 2: strong scaling of C3 on N GPUs, default and experimental CG kernel.  N = number of GPUs of the gpurun call:
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/round2_scaling.sh 8'
N=${1:-2}
mkdir -p gpurun_out
for k in default x; do
  if [ $k = x ]; then export AVS_PCG_KERNEL=x; else unset AVS_PCG_KERNEL; fi
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_bench_c3_g${N}_$k.json 2> gpurun_out/r2_bench_c3_g${N}_$k.err
done
