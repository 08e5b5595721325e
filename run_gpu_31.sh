timeout -k 5 300 python -m pytest tests/test_gpu_adam.py -q 2>&1 | tail -25 > gpurun_out/pytest_adam.log; tail -5 gpurun_out/pytest_adam.log
timeout -k 5 300 python tools/bench_layers.py --only adam > gpurun_out/bench_adam.jsonl 2> gpurun_out/bench_adam.err; cat gpurun_out/bench_adam.jsonl | cut -c1-400; tail -3 gpurun_out/bench_adam.err
timeout -k 5 400 python bench.py > gpurun_out/bench_31.json 2> gpurun_out/bench_31.err; cat gpurun_out/bench_31.json; tail -3 gpurun_out/bench_31.err
