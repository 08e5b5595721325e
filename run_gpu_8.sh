mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -k 5 300 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -40 > gpurun_out/pytest_8_sharded.log; tail -25 gpurun_out/pytest_8_sharded.log
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cin_bwd|split_" -c 12 --csv --log-file gpurun_out/launches_cin_bwd.csv python tools/bench_layers.py --only cin --iters 2 > /dev/null 2>&1; grep -v "^==" gpurun_out/launches_cin_bwd.csv | awk -F'","' '{print $5, $NF}' | tail -12
