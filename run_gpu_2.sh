mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_2.log; tail -60 gpurun_out/pytest_2.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_din.py -q -k "test_din_fwd_bwd and (9-7-8 or 2-3-4)" 2>&1 | tail -25 > gpurun_out/sanitizer_din.log; tail -12 gpurun_out/sanitizer_din.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fibinet.py -q -k "4-8-8 or 3-8-8-4" 2>&1 | tail -25 > gpurun_out/sanitizer_fib.log; tail -12 gpurun_out/sanitizer_fib.log
