timeout -k 5 1200 python -m pytest tests -q -m gpu 2>&1 > gpurun_out/pytest_gpu_full.log; grep -E "^E  |passed|failed|error" gpurun_out/pytest_gpu_full.log | head
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
