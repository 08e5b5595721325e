mkdir -p gpurun_out
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29535 tools/dbg_peer.py > gpurun_out/dbg_peer.log 2>&1
grep -vE "^W0|OMP_NUM|^\*\*\*|frame #" gpurun_out/dbg_peer.log | head -40 | cut -c1-250
