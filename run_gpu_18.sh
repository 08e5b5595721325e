mkdir -p gpurun_out
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_sharded_gpu.py > gpurun_out/sharded_18.log 2>&1; grep -vE "^W0|OMP_NUM|^\*\*\*|frame #" gpurun_out/sharded_18.log | grep -E "SHARDED_OK|Error|error|assert|line " | head -12
for R in 12500000; do
CTR_BENCH_ROWS=$R timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --workload deepfm_cfg5_sharded --steps 30 2> gpurun_out/bench_sh.err | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('SYMM rows/field', d['config']['rows_per_field'], 'table GB', d['config']['table_bytes_total']/1e9, 'fwd_ms', r['fwd_ms'], 'bwd_push_ms', r['bwd_push_ms'], 'pull GB/s', r['achieved_pull_GBps'])"
grep -vE "^W0|OMP_NUM|^\*\*\*|frame #" gpurun_out/bench_sh.err | tail -4
done
