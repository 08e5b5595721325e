mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_cin.py -q 2>&1 | tail -40 > gpurun_out/pytest_23_cin.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_23_cin.log | tail -12
timeout -k 5 240 python tools/bench_layers.py --only cin --iters 5 > gpurun_out/bench_layers_r1_i.jsonl 2> gpurun_out/bench_layers.err; cut -c1-170 gpurun_out/bench_layers_r1_i.jsonl; tail -3 gpurun_out/bench_layers.err
timeout -k 5 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"cin_bwd_d" -s 2 -c 2 --csv --log-file gpurun_out/launches_cin_bwd2.csv python tools/prof_cin_dx.py > /dev/null 2>&1; grep -v "^==" gpurun_out/launches_cin_bwd2.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $NF}' | tail -4
