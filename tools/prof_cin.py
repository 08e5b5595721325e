"""Profile helper: CIN forward + backward at BASELINE config 3, layer 2 (ncu -k regex:cin_(fwd_tc|bwd_d))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recalgorithm_b200 import ops
torch.manual_seed(0)
B, m, D, H = 8192, 30, 16, 128
x0 = torch.randn((B, m, D), device="cuda") * 0.25
xk = torch.randn((B, H, D), device="cuda") * 0.25
w = torch.randn((H * m, H), device="cuda") * 0.05
g = torch.randn((B, H, D), device="cuda")
for _ in range(2):
    ops.cin_fwd(x0, xk, w, want_pooled=True)
    ops.cin_bwd(x0, xk, w, g)
torch.cuda.synchronize()
