mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_sharded_gpu.py > gpurun_out/sharded_10.log 2>&1; grep -E "rank [01]\]|SHARDED_OK|Error|error:|assert" gpurun_out/sharded_10.log | head -20
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:cin_bwd_dx -s 1 -c 1 -o gpurun_out/prof_cin_dx -f python tools/prof_cin_dx.py > gpurun_out/ncu_dx.log 2>&1; tail -2 gpurun_out/ncu_dx.log
