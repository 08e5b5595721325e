mkdir -p gpurun_out
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:"cin_bwd_d" -s 2 -c 2 -o gpurun_out/prof_cin_bwd -f python tools/prof_cin_dx.py > gpurun_out/ncu_cin_bwd.log 2>&1; tail -2 gpurun_out/ncu_cin_bwd.log
