#!/bin/bash
# Very last GPU call of round 2 (~45 GPU-seconds left): the multi-rank error agreement added to avs_stage_system (2 ranks sharing the GPU).
export PYTHONUNBUFFERED=1
timeout 25 python -m pytest tests/test_gpu_inprocess_multi.py -v --tb=short -p no:cacheprovider \
    -k "refuses_a_scene" > gpurun_out/r2final2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2final2_pytest.log
tail -25 gpurun_out/r2final2_pytest.log
