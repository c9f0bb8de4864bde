#!/bin/bash
# Last GPU call of round 2 (about 80 GPU-seconds were left): the two tests added after the budget was spent -- the CUDA path against
# the compiled reference on random scenes (incl. the 65-entry rows that needed MAX_ROW > 64) and the refusal of a scene outside the
# reference's contract -- then smoke() if time remains.  Logs in gpurun_out/r2final_*.
export PYTHONUNBUFFERED=1
timeout 40 python -m pytest tests/test_gpu_reference.py -v --tb=short -p no:cacheprovider \
    -k "random_scenes or outside_the_reference" > gpurun_out/r2final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
tail -25 gpurun_out/r2final_pytest.log
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2final_smoke.log
tail -3 gpurun_out/r2final_smoke.log
