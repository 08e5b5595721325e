mkdir -p gpurun_out
for R in 6250000 12500000; do
CTR_BENCH_ROWS=$R timeout -k 5 300 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('rows/field', d['config']['rows_per_field'], 'fwd_ms', r['avg_launch_ms'], 'frac', r['frac'], 'bwd frac', d['roofline_bwd']['frac'])"
done
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:"cross_bwd|din_attention_bwd" -c 4 -o gpurun_out/prof_small -f python tools/prof_small.py > gpurun_out/ncu_small.log 2>&1; tail -2 gpurun_out/ncu_small.log
