"""Profile helper (run under ncu): one launch each of the 8f kernels at their bench shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recalgorithm_b200 import autograd, ops, optim
torch.manual_seed(0)
rn = lambda *s, std=1.0: torch.randn(s, device="cuda") * std
# FwFM / AFM
B, F, K = 65536, 40, 32
x, r, g1 = rn(B, F, K, std=0.2), rn(F * (F - 1) // 2, std=0.3), rn(B)
B2, F2, K2, T2 = 65536, 7, 8, 128
xa, w, b, h, gk = rn(B2, F2, K2, std=0.5), rn(K2, T2, std=0.3), rn(T2, std=0.1), rn(T2, std=0.3), rn(B2, K2)
# BST
Bb, Tb, db, Hb = 4096, 51, 8, 3
xb, gb = rn(Bb, Tb, db), rn(Bb, Tb, db)
klen = torch.randint(1, Tb + 1, (Bb,), device="cuda")
packed = rn(int(ops._lib.lib().ctr_bst_param_count(db, Hb, Tb)), std=0.3)
# lazy Adam at a reduced table (same access pattern per id)
rows = 500_000
tables = autograd.EmbeddingTables([rows] * 40, 32, device="cuda", init=None)
tables.weight.normal_(0, 0.2)
ids = torch.randint(0, rows, (65536, 40), device="cuda")
vals = rn(65536, 40, 32)
opt = optim.TableAdam(tables, lr=1e-3, lazy=True)
for _ in range(2):
    ops.fwfm_fwd(x, r); ops.fwfm_bwd(x, r, g1)
    ops.afm_fwd(xa, w, b, h); ops.afm_bwd(xa, w, b, h, gk)
    ops.bst_transformer_fwd(xb, xb, xb, klen, packed, Hb, Tb); ops.bst_transformer_bwd(xb, xb, xb, klen, packed, gb, Hb, Tb)
    tables.grad_slices.append(autograd.IndexedSlices(vals.clone(), ids, tables.field_row_offset)); opt.step()
torch.cuda.synchronize()
