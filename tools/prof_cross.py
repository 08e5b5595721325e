"""Profile helper: DCN cross forward + backward at d = 480, L = 3, B = 65536 (ncu -k regex:cross_)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recalgorithm_b200 import ops
torch.manual_seed(0)
B, d, L = 65536, 480, 3
x0 = torch.randn((B, d), device="cuda")
w = torch.randn((L, d), device="cuda") * 0.05
b = torch.randn((L, d), device="cuda") * 0.05
g = torch.randn((B, d), device="cuda")
for _ in range(2):
    ops.cross_fwd(x0, w, b)
    ops.cross_bwd(x0, w, b, g)
torch.cuda.synchronize()
