mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cin.py -q 2>&1 | tail -80 > gpurun_out/pytest_4_cin.log; grep -E "passed|failed|Error" gpurun_out/pytest_4_cin.log | tail -20
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_cin.py 2>&1 | tail -100 > gpurun_out/pytest_4.log; grep -E "passed|failed|Error" gpurun_out/pytest_4.log | tail -20
timeout 600 python tools/bench_layers.py > gpurun_out/bench_layers_r1_a.jsonl 2> gpurun_out/bench_layers.err; cat gpurun_out/bench_layers_r1_a.jsonl | cut -c1-400; tail -3 gpurun_out/bench_layers.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cin_fwd_tc -c 4 -o gpurun_out/prof_cin_r1_a -f python tools/bench_layers.py --only cin --iters 3 > gpurun_out/ncu_cin.log 2>&1; tail -3 gpurun_out/ncu_cin.log
