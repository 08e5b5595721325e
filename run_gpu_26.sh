mkdir -p gpurun_out
timeout -k 5 300 python tools/bench_layers.py --configs > gpurun_out/bench_configs_r1.jsonl 2> gpurun_out/bench_configs.err; cut -c1-330 gpurun_out/bench_configs_r1.jsonl; tail -4 gpurun_out/bench_configs.err
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:embed_fm2 -s 6 -c 2 -o gpurun_out/prof_embed_r1_final -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_embed.log 2>&1; tail -1 gpurun_out/ncu_embed.log | cut -c1-200
