import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from recalgorithm_b200 import _lib, ops
dev = torch.device("cuda", local)
w = torch.full((1000, 32), float(rank + 1), device=dev)
h = [None] * world
dist.all_gather_object(h, reduce_tensor(w))
peer = 1 - rank
rc = _lib.lib().ctr_enable_peer_access(peer); print(rank, "enable_peer rc", rc, flush=True)
fn, args = h[peer]
pt = fn(*args)
print(rank, "peer tensor device", pt.device, hex(pt.data_ptr()), flush=True)
x = pt[:2, :2].to(dev); torch.cuda.synchronize(); print(rank, "torch p2p copy ok", x.flatten().tolist(), flush=True)
# my kernel with the peer table directly (ld.global.nc on peer memory)
off = torch.tensor([0, 1000], device=dev); ids = torch.arange(8, device=dev).reshape(8, 1)
try:
    tile, _ = ops.embed_fm2_fwd(pt, off, ids, want_fm2=False); torch.cuda.synchronize()
    print(rank, "nc-load on peer ok", tile[0, 0, :2].tolist(), flush=True)
except Exception as e:
    print(rank, "nc-load on peer FAILED", repr(e)[:200], flush=True)
dist.barrier(); dist.destroy_process_group()
