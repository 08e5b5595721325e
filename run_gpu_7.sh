mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_cin.py -q -k "bwd" 2>&1 | tail -70 > gpurun_out/pytest_7_cin.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_7_cin.log | tail -30
timeout -k 5 240 python tools/bench_layers.py --only cin --iters 5 > gpurun_out/bench_layers_r1_d.jsonl 2> gpurun_out/bench_layers.err; cut -c1-200 gpurun_out/bench_layers_r1_d.jsonl | tail -3; tail -3 gpurun_out/bench_layers.err
