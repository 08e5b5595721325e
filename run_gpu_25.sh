mkdir -p gpurun_out
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; grep "^{" gpurun_out/bench_n2.json | cut -c1-700; grep -vE "^W0|OMP_NUM|^\*\*\*|frame #" gpurun_out/bench_n2.err | tail -4
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>/dev/null | grep "^{" | cut -c1-300
timeout -k 5 300 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_sharded.py -q 2>&1 | tail -5
