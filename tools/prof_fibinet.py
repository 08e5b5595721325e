"""Profile helper: FiBiNET bilinear forward + backward, F = 30, K = 16, B = 4096, all three types (ncu -k regex:bilinear)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recalgorithm_b200 import ops
torch.manual_seed(0)
B, F, K = 4096, 30, 16
P = (F - 1) * (F - 2) // 2
x = torch.randn((B, F, K), device="cuda") * 0.25
g = torch.randn((B, P, K), device="cuda")
for typ in ("all", "each", "interaction"):
    w = torch.randn(ops.bilinear_w_shape(F, K, typ), device="cuda") * 0.2
    for _ in range(2):
        ops.bilinear_fwd(x, w, typ)
        ops.bilinear_bwd(x, w, typ, g)
torch.cuda.synchronize()
