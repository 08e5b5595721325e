"""Profile helper: cross backward (B=65536, d=480, L=3) and DIN backward (config 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recalgorithm_b200 import ops
torch.manual_seed(0)
B, d, L = 65536, 480, 3
x0 = torch.randn((B, d), device="cuda"); w = torch.randn((L, d), device="cuda") * 0.05
b = torch.randn((L, d), device="cuda") * 0.05; g = torch.randn((B, d), device="cuda")
for _ in range(2):
    ops.cross_bwd(x0, w, b, g)
B, T, H = 4096, 50, 16
q = torch.randn((B, H), device="cuda") * 0.25; k = torch.randn((B, T, H), device="cuda") * 0.25
lens = torch.randint(0, T + 1, (B,), device="cuda")
ws = [torch.randn(s, device="cuda") * 0.2 for s in ((4 * H, 64), (64,), (64, 32), (32,), (32, 1), (1,))]
go = torch.randn((B, H), device="cuda")
for _ in range(2):
    ops.din_attention_bwd(q, k, lens, *ws, go)
torch.cuda.synchronize()
