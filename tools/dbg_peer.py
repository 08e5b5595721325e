import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from recalgorithm_b200 import _lib, ops
L = _lib.lib()
dev = torch.device("cuda", local)
peer = 1 - rank

class _Raw:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

def view(ptr, shape, dtype=torch.float32):
    ts = {torch.float32: "<f4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_Raw(ptr, shape, ts), device=dev)

# ---- (B) own IPC: cudaMalloc + cudaIpcGetMemHandle / OpenMemHandle in the consumer's device context
p = ctypes.c_void_p()
_lib.check(L.ctr_peer_alloc(1000 * 32 * 4, ctypes.byref(p)))
mine = view(p.value, (1000, 32)); mine.fill_(float(rank + 1)); torch.cuda.synchronize()
hbuf = ctypes.create_string_buffer(64)
_lib.check(L.ctr_ipc_export(p, hbuf))
hs = [None] * world
dist.all_gather_object(hs, bytes(hbuf.raw))
q = ctypes.c_void_p()
_lib.check(L.ctr_ipc_import(hs[peer], ctypes.byref(q)))
print(rank, "own-ipc import ok", hex(q.value), flush=True)
pt = view(q.value, (1000, 32))
off = torch.tensor([0, 1000], device=dev); ids = torch.arange(8, device=dev).reshape(8, 1)
try:
    tile, _ = ops.embed_fm2_fwd(pt, off, ids, want_fm2=False); torch.cuda.synchronize()
    print(rank, "own-ipc: nc-load on peer ok", tile[0, 0, :2].tolist(), flush=True)
except Exception as e:
    print(rank, "own-ipc: nc-load on peer FAILED", repr(e)[:300], flush=True)
try:
    pt[5].fill_(42.0); torch.cuda.synchronize(); dist.barrier()
    print(rank, "own-ipc: peer store ok; my row5 now", mine[5, 0].item(), flush=True)
except Exception as e:
    print(rank, "own-ipc: peer store FAILED", repr(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
