mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cin.py -q -x 2>&1 | tail -120 > gpurun_out/pytest_3_cin.log; tail -50 gpurun_out/pytest_3_cin.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_cin.py 2>&1 | tail -150 > gpurun_out/pytest_3.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_3.log | tail -30
