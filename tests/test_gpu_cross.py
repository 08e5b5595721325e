"""GPU parity: DCN cross stack (row CROSS) through the C ABI vs golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev, golden, trunc_normal
from oracle import layers_np as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cross_d82_L1", "cross_d82_L3", "cross_d480_L3"])
def test_cross_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    out = ops.cross_fwd(dev(g["x0"]), dev(g["ws"]), dev(g["bs"]))
    assert_close(out, g["out_f64"], TOL, "cross fwd vs reference(float64)")
    assert_close(out, g["out_f32"], TOL, "cross fwd vs reference(float32)")


@pytest.mark.parametrize("B,d,L", [(1, 1, 1), (9, 82, 3), (64, 480, 3), (33, 480, 1), (17, 130, 8), (5, 1024, 2),
                                   (260, 66, 4), (8, 36, 2), (40, 1000, 3), (3, 7, 5)])
def test_cross_fwd_bwd(B, d, L):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + d + L)
    x0 = trunc_normal(rng, (B, d), 0.5)
    lim = (6.0 / (d + 1)) ** 0.5
    ws = rng.uniform(-lim, lim, (L, d)).astype(np.float32)
    bs = rng.uniform(-lim, lim, (L, d)).astype(np.float32)
    g = trunc_normal(rng, (B, d), 1.0)
    x64, w64, b64, g64 = (a.astype(np.float64) for a in (x0, ws, bs, g))
    assert_close(ops.cross_fwd(dev(x0), dev(ws), dev(bs)), O.cross_stack_fwd(x64, w64, b64)[-1], TOL, "fwd")
    dx0, dxl, dw, db = ops.cross_bwd(dev(x0), dev(ws), dev(bs), dev(g))
    ex0, ew, eb = O.cross_stack_bwd(x64, w64, b64, g64)
    assert dxl is None
    assert_close(dx0, ex0, TOL, "dx0"); assert_close(dw, ew, TOL, "dw"); assert_close(db, eb, TOL, "db")


def test_cross_single_layer_signature():
    """cross_layer(x0, xl, index): one layer with an explicit xl -- composes to the same stack."""
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(5)
    B, d, L = 21, 82, 3
    x0 = trunc_normal(rng, (B, d), 0.5)
    ws = trunc_normal(rng, (L, d), 0.2); bs = trunc_normal(rng, (L, d), 0.2)
    x = dev(x0)
    for l in range(L):
        x = ops.cross_fwd(dev(x0), dev(ws[l:l + 1]), dev(bs[l:l + 1]), xl_in=x)
    assert_close(x, O.cross_stack_fwd(x0.astype(np.float64), ws.astype(np.float64), bs.astype(np.float64))[-1], TOL)
    # backward of one layer with explicit xl vs torch autograd (float64)
    xl = trunc_normal(rng, (B, d), 0.5); g = trunc_normal(rng, (B, d), 1.0)
    dx0, dxl, dw, db = ops.cross_bwd(dev(x0), dev(ws[:1]), dev(bs[:1]), dev(g), xl_in=dev(xl))
    t = lambda a: torch.tensor(a.astype(np.float64), requires_grad=True)
    x0t, xlt, wt, bt = t(x0), t(xl), t(ws[0]), t(bs[0])
    (x0t * (xlt @ wt)[:, None] + bt[None, :] + xlt).backward(torch.tensor(g.astype(np.float64)))
    assert_close(dx0, x0t.grad, TOL, "dx0"); assert_close(dxl, xlt.grad, TOL, "dxl")
    assert_close(dw[0], wt.grad, TOL, "dw"); assert_close(db[0], bt.grad, TOL, "db")


def test_cross_config2_properties():
    """BASELINE config 2 (B=4096, d=480, L=3): zero weights => identity + biases; determinism of fwd."""
    from recalgorithm_b200 import ops
    B, d, L = 4096, 480, 3
    gen = torch.Generator(device="cuda").manual_seed(7)
    x0 = torch.randn((B, d), device="cuda", generator=gen)
    b = torch.randn((L, d), device="cuda", generator=gen) * 0.1
    out = ops.cross_fwd(x0, torch.zeros((L, d), device="cuda"), b)
    assert_close(out, x0.double() + b.double().sum(0)[None, :], 1e-6, "w=0 -> x0 + sum b")
    w = torch.randn((L, d), device="cuda", generator=gen) * 0.05
    o1, o2 = ops.cross_fwd(x0, w, b), ops.cross_fwd(x0, w, b)
    assert torch.equal(o1, o2)
    x, x64 = x0.double(), x0.double()
    for l in range(L):
        x = x64 * (x @ w[l].double())[:, None] + b[l].double()[None, :] + x
    assert_close(o1, x, TOL, "config-2 forward vs float64 torch")


@pytest.mark.parametrize("B,F,D,L,rows,i32", [(1, 1, 4, 1, 3, False), (9, 6, 8, 3, 11, False), (64, 30, 16, 3, 50, True),
                                              (260, 8, 12, 4, 7, False), (37, 16, 32, 2, 5, True), (5, 3, 4, 4, 2, False)])
def test_lookup_cross_fused_forward(B, F, D, L, rows, i32):
    """ctr_embed_cross_fwd = input_layer gather + the cross loop (DCN/dcn.py:153-160) in one launch: x0 is the bit-exact gather
    (OOV / out-of-range ids -> zero vector), x_L within 1e-5 of the float64 oracle."""
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B * 7 + F + D + L)
    d = F * D
    table = trunc_normal(rng, (rows * F, D), D ** -0.5)
    ids = rng.integers(-1, rows + 1, (B, F))                       # -1 (OOV) and `rows` (out of range) included
    lim = (6.0 / (d + 1)) ** 0.5
    ws = rng.uniform(-lim, lim, (L, d)).astype(np.float32)
    bs = rng.uniform(-lim, lim, (L, d)).astype(np.float32)
    off = torch.arange(F + 1, device="cuda", dtype=torch.int64) * rows
    ids_t = torch.from_numpy(ids).cuda()
    if i32:
        ids_t = ids_t.int()
    x0, out = ops.embed_cross_fwd(dev(table), off, ids_t, dev(ws), dev(bs))
    valid = (ids >= 0) & (ids < rows)
    ex0 = np.where(valid[..., None], table[(np.arange(F)[None, :] * rows + np.clip(ids, 0, rows - 1))], 0.0).reshape(B, d)
    assert np.array_equal(x0.cpu().numpy(), ex0.astype(np.float32))          # the gather is a copy: bit-exact
    e = O.cross_stack_fwd(ex0.astype(np.float64), ws.astype(np.float64), bs.astype(np.float64))[-1]
    assert_close(out, e, TOL, "fused lookup+cross fwd")
    # and equal to the two-launch form within rounding of the reformulated recurrence
    assert_close(out, ops.cross_fwd(x0, dev(ws), dev(bs)).double().cpu().numpy(), TOL, "fused vs two launches")


def test_lookup_cross_unsupported_shapes_use_two_launches():
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(5)
    B, F, D, L, rows = 7, 40, 16, 5, 9                               # d = 640 > 512 and L = 5 > 4
    assert not ops.embed_cross_supported(F, D, L)
    table = trunc_normal(rng, (rows * F, D), 0.25)
    ids = torch.from_numpy(rng.integers(0, rows, (B, F))).cuda()
    ws, bs = trunc_normal(rng, (L, F * D), 0.05), trunc_normal(rng, (L, F * D), 0.05)
    off = torch.arange(F + 1, device="cuda", dtype=torch.int64) * rows
    x0, out = ops.embed_cross_fwd(dev(table), off, ids, dev(ws), dev(bs))
    e = O.cross_stack_fwd(x0.double().cpu().numpy(), ws.astype(np.float64), bs.astype(np.float64))[-1]
    assert_close(out, e, TOL, "fallback chain")


def test_lookup_cross_autograd_matches_unfused_chain():
    """autograd.lookup_cross (one launch forward; both outputs differentiable) == lookup + cross_stack, including the table
    gradient when x0 also feeds a deep tower (DCN/dcn.py:163)."""
    from recalgorithm_b200 import autograd
    torch.manual_seed(3)
    B, F, D, L, rows = 50, 30, 16, 3, 40
    d = F * D
    ids = torch.randint(-1, rows, (B, F), device="cuda")
    w0 = torch.randn((L, d), device="cuda") * 0.05
    b0 = torch.randn((L, d), device="cuda") * 0.05
    deep = torch.randn((d, 1), device="cuda") * 0.1
    gy = torch.randn((B, d), device="cuda")
    res = []
    for fused in (True, False):
        tables = autograd.EmbeddingTables([rows] * F, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(11))
        w, b = w0.clone().requires_grad_(), b0.clone().requires_grad_()
        if fused:
            xl, x0 = autograd.lookup_cross(tables, ids, w, b)
        else:
            x0 = autograd.lookup(tables, ids).reshape(B, d)
            xl = autograd.cross_stack(x0, w, b)
        loss = (xl * gy).sum() + (x0 @ deep).sum()
        loss.backward()
        dense = sum(s.to_dense(tables.num_rows) for s in tables.grad_slices)
        res.append((xl.detach(), w.grad, b.grad, dense))
    for a, e, name in zip(res[0], res[1], ("x_L", "dw", "db", "table grad")):
        assert_close(a, e.double().cpu().numpy(), TOL, name)
