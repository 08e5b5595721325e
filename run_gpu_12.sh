mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_cin.py -q -k "bwd" 2>&1 | tail -70 > gpurun_out/pytest_12_cin.log; grep -E "passed|failed|Error|error" gpurun_out/pytest_12_cin.log | tail -30
timeout -k 5 240 python tools/bench_layers.py --only cin --iters 5 > gpurun_out/bench_layers_r1_e.jsonl 2> gpurun_out/bench_layers.err; cut -c1-200 gpurun_out/bench_layers_r1_e.jsonl | tail -6; tail -3 gpurun_out/bench_layers.err
CTR_EMBED_OCC4=1 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('OCC4 value',d['value'],'fwd frac',d['roofline']['frac'],'bwd frac',d['roofline_bwd']['frac'],'e2e',d['e2e']['value'])"
python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('base value',d['value'],'fwd frac',d['roofline']['frac'],'bwd frac',d['roofline_bwd']['frac'],'e2e',d['e2e']['value'])"
