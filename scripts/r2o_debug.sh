mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8 9 10; do
  AVS_TRACE=1 timeout 300 python -m pytest tests/test_gpu_hdk_shim.py tests/test_gpu_inprocess_multi.py -q --tb=short -x > gpurun_out/r2o_debug_$i.log 2>&1
  tail -1 gpurun_out/r2o_debug_$i.log
  if grep -q "failed" gpurun_out/r2o_debug_$i.log; then grep -n "Error\|error\|rank\|E  " gpurun_out/r2o_debug_$i.log | head -30 | cut -c1-400; break; fi
done
