timeout -k 5 400 python tools/bench_e2e_tfrecord.py 2>&1 | tail -2 | tee gpurun_out/bench_e2e_tfrecord.json
