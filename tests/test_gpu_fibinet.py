"""GPU parity: FiBiNET SENET + bilinear (rows SENET, BILINEAR) vs golden vectors and oracle."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev, golden, trunc_normal
from oracle import layers_np as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["fibinet_F8_K8", "fibinet_F30_K16"])
def test_fibinet_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    x = dev(g["x"])
    assert_close(ops.senet_fwd(x, dev(g["senet_w1"]), dev(g["senet_w2"])), g["senet_f64"], TOL, "senet")
    F = g["x"].shape[1]
    for mask in (4, 7, 8):                                     # default, tournament kernels for every type, round-1 kernels
        prev = ops.bilinear_set_tournament(mask)
        try:
            for typ in ("all", "each", "interaction"):
                out = ops.bilinear_fwd(x, dev(g[f"w_{typ}"]), typ)
                assert out.shape[1] == (F - 1) * (F - 2) // 2          # the reference's range(F-1) quirk
                assert_close(out, g[f"bilinear_{typ}_f64"], TOL, f"bilinear {typ}")
        finally:
            ops.bilinear_set_tournament(prev)
    with pytest.raises(ValueError):
        ops.bilinear_fwd(x, dev(g["w_all"]), "nope")


@pytest.mark.parametrize("B,F,K,r", [(3, 8, 8, 4), (70, 30, 16, 8), (9, 5, 4, 2), (33, 12, 32, 16), (2, 40, 16, 3)])
def test_senet_fwd_bwd(B, F, K, r):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + F + K)
    x = trunc_normal(rng, (B, F, K), 1.0)
    w1 = trunc_normal(rng, (F, r), 0.5); w2 = trunc_normal(rng, (r, F), 0.5)
    g = trunc_normal(rng, (B, F, K), 1.0)
    d = lambda a: a.astype(np.float64)
    assert_close(ops.senet_fwd(dev(x), dev(w1), dev(w2)), O.senet_fwd(d(x), d(w1), d(w2)), TOL, "fwd")
    dx, dw1, dw2 = ops.senet_bwd(dev(x), dev(w1), dev(w2), dev(g))
    ex, e1, e2 = O.senet_bwd(d(x), d(w1), d(w2), d(g))
    assert_close(dx, ex, TOL, "dx"); assert_close(dw1, e1, TOL, "dw1"); assert_close(dw2, e2, TOL, "dw2")
    with pytest.raises(ValueError):          # reference: assert reduction_dim < embedding_dim (senet.py:19)
        ops.senet_fwd(dev(x), dev(trunc_normal(rng, (F, K), 1.0)), dev(trunc_normal(rng, (K, F), 1.0)))


@pytest.fixture(params=[4, 7, 7 | (1 << 10), 7 | (4 << 10) | (16 << 4), 8, 0],
                ids=["default", "tournament", "tournament_kt1", "tournament_kt4_tile16", "round1_kernels", "staged_all_each_round1_interaction"])
def bilinear_impl(request):
    """Every kernel family: default (staged per-sample kernels for 'all'/'each', tournament kernels for 'interaction'), the
    tournament kernels for all three types in three register-blocking variants, and the round-1 CTA-per-sample kernels."""
    from recalgorithm_b200 import ops
    prev = ops.bilinear_set_tournament(request.param)
    yield request.param
    ops.bilinear_set_tournament(prev)


@pytest.mark.parametrize("B,F,K", [(4, 8, 8), (37, 30, 16), (3, 3, 4), (5, 4, 16), (300, 9, 8), (2, 2, 4), (19, 3, 16), (50, 7, 32),
                                   (65, 41, 16), (130, 70, 8), (1, 30, 16)])
@pytest.mark.parametrize("typ", ["all", "each", "interaction"])
def test_bilinear_fwd_bwd(B, F, K, typ, bilinear_impl):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + F + K)
    x = trunc_normal(rng, (B, F, K), 1.0)
    w = trunc_normal(rng, ops.bilinear_w_shape(F, K, typ), 0.4)
    P = (F - 1) * (F - 2) // 2
    g = trunc_normal(rng, (B, P, K), 1.0)
    d = lambda a: a.astype(np.float64)
    out = ops.bilinear_fwd(dev(x), dev(w), typ)
    assert out.shape == (B, P, K)
    if P == 0:
        return
    assert_close(out, O.bilinear_fwd(d(x), d(w), typ), TOL, "fwd")
    dx, dw = ops.bilinear_bwd(dev(x), dev(w), typ, dev(g))
    ex, ew = O.bilinear_bwd(d(x), d(w), typ, d(g))
    assert_close(dx, ex, TOL, "dx"); assert_close(dw, ew, TOL, "dw")
    assert torch.all(dx[:, F - 1, :] == 0)       # the last field never takes part
