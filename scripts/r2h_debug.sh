mkdir -p gpurun_out
AVS_TRACE=1 timeout 200 python -m pytest tests/test_gpu_inprocess_multi.py -q --tb=short -x -k "rank_solve and 2" > gpurun_out/r2h_multi.log 2>&1; grep -v "avs_stage\|sjds build" gpurun_out/r2h_multi.log | tail -60
