mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; }
run base AVS_X=0
run dyn AVS_PCG_DYN=1
run minb5 AVS_PCG_MINB=5
run minb5dyn AVS_PCG_MINB=5 AVS_PCG_DYN=1
run nc AVS_PCG_NC=1
run ncdyn AVS_PCG_NC=1 AVS_PCG_DYN=1
run minb3 AVS_PCG_MINB=3
run launch AVS_CG_MODE=launch
