timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29535 tools/dbg_peer.py 2>&1 | grep -vE "^W0|OMP_NUM|^\*\*\*|frame #" | tail -30
nvidia-smi topo -m | head -8
