mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_cin.py -q -x 2>&1 | tail -60 > gpurun_out/pytest_6_cin.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_6_cin.log | tail -12
timeout -k 5 240 python tools/bench_layers.py --only cin > gpurun_out/bench_layers_r1_c.jsonl 2> gpurun_out/bench_layers.err; cut -c1-260 gpurun_out/bench_layers_r1_c.jsonl | head -4; tail -3 gpurun_out/bench_layers.err
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:cin_fwd_tc -s 4 -c 3 -o gpurun_out/prof_cin_r1_b -f python tools/bench_layers.py --only cin --iters 2 > gpurun_out/ncu_cin.log 2>&1; tail -2 gpurun_out/ncu_cin.log
