mkdir -p gpurun_out
for R in 200000 2000000; do
CTR_BENCH_ROWS=$R timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --workload deepfm_cfg5_sharded --steps 30 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('rows/field', d['config']['rows_per_field'], 'table GB', d['config']['table_bytes_total']/1e9, 'fwd_ms', r['fwd_ms'], 'bwd_push_ms', r['bwd_push_ms'], 'pull GB/s', r['achieved_pull_GBps'])"
done
