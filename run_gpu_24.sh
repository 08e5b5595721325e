mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_din.py tests/test_gpu_layers.py tests/test_gpu_cin.py -q 2>&1 | tail -40 > gpurun_out/pytest_24.log; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_24.log | tail -12
timeout -k 5 240 python tools/bench_layers.py --only din,cin --iters 10 > gpurun_out/bench_layers_r1_j.jsonl 2> gpurun_out/bench_layers.err; cut -c1-170 gpurun_out/bench_layers_r1_j.jsonl; tail -3 gpurun_out/bench_layers.err
timeout -k 5 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_din.py -q -k "test_din_fwd_bwd and (9-7-8 or 2-3-4 or 3-20-5)" 2>&1 | tail -5
