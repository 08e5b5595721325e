for v in "" "CTR_EMBED_NOPREFETCH=1" "CTR_EMBED_OCC1=1" "CTR_EMBED_OCC1=1 CTR_EMBED_NOPREFETCH=1"; do
env $v python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value',round(d['value']/1e6,1),'fwd frac',round(d['roofline']['frac'],4),'fwd ms',round(d['roofline']['avg_launch_ms'],4),'bwd frac',round(d['roofline_bwd']['frac'],4))"
done
python -m pytest tests/test_gpu_embed_fm2.py -q 2>&1 | tail -2
