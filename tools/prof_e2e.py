#!/usr/bin/env python
"""Component timings of the public-API e2e step at BASELINE config 5 (CUDA events, 20 iterations each)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from recalgorithm_b200 import autograd, ops  # noqa: E402


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B, F, D, rows = 65536, 40, 32, 2_500_000
    gen = torch.Generator(device="cuda").manual_seed(0)
    tables = autograd.EmbeddingTables([rows] * F, D, device="cuda", init=None)
    tables.weight.normal_(0, D ** -0.5, generator=gen)
    ids = [torch.randint(0, rows, (B, F), device="cuda", generator=gen) for _ in range(4)]
    ids32 = [i.int() for i in ids]
    w = (torch.randn((F * D, 1), device="cuda", generator=gen) * 0.01).requires_grad_()
    lab = (torch.rand((B, 1), device="cuda", generator=gen) < 0.03).float()
    k = [0]
    res = {}

    def nxt(lst):
        k[0] += 1
        return lst[k[0] % 4]
    res["fwd_plain_i64"] = t(lambda: ops.embed_fm2_fwd(tables.weight, tables.field_row_offset, nxt(ids)))
    ids64 = torch.empty((B, F), dtype=torch.int64, device="cuda")
    res["fwd_lin_i64"] = t(lambda: ops.embed_fm2_lin_fwd(tables.weight, tables.field_row_offset, nxt(ids), w.detach()))
    res["fwd_lin_i32"] = t(lambda: ops.embed_fm2_lin_fwd(tables.weight, tables.field_row_offset, nxt(ids32), w.detach(), ids64_out=ids64))
    tile, fm2, lin = ops.embed_fm2_lin_fwd(tables.weight, tables.field_row_offset, ids[0], w.detach())
    g = torch.randn((B,), device="cuda") * 0.01
    res["lin_bwd"] = t(lambda: ops.embed_fm2_lin_bwd(tile, w.detach(), g, g))
    res["sigmoid_ce_kernel"] = t(lambda: ops.sigmoid_ce(fm2, lin, lab))

    def ce_autograd():
        a = fm2.detach().requires_grad_(); b = lin.detach().requires_grad_()
        autograd.sigmoid_cross_entropy_mean(a, lab, logit_b=b).backward()
    res["sigmoid_ce_autograd_fwd_bwd"] = t(ce_autograd)

    def bce_torch():
        a = fm2.detach().requires_grad_(); b = lin.detach().requires_grad_()
        torch.nn.functional.binary_cross_entropy_with_logits(a + b, lab).backward()
    res["torch_bce_fwd_bwd"] = t(bce_torch)

    def step_fused():
        tables.zero_grad(); w.grad = None
        f_, l_ = autograd.lookup_fm2_linear(tables, nxt(ids32), w)
        autograd.sigmoid_cross_entropy_mean(f_, lab, logit_b=l_).backward()
    res["step_fused_device_only"] = t(step_fused)

    def step_fused_torch_loss():
        tables.zero_grad(); w.grad = None
        f_, l_ = autograd.lookup_fm2_linear(tables, nxt(ids32), w)
        torch.nn.functional.binary_cross_entropy_with_logits(f_ + l_, lab).backward()
    res["step_fused_torch_loss_device_only"] = t(step_fused_torch_loss)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
