"""GPU parity: BST transformer block vs fixtures produced by executing BST/transformer_layer.py (float32, the
precision whose mask arithmetic defines the result) and vs the float64 oracle / its autograd backward."""
import numpy as np
import pytest
import torch

from _util import TOL, assert_close, dev, golden, trunc_normal
from oracle import bst_torch
from oracle import layers_np as O

pytestmark = pytest.mark.gpu


def _pack(p, d, H, maxlen):
    from recalgorithm_b200 import ops
    return ops.bst_pack_params({k: dev(v) for k, v in p.items()}, d, H, maxlen)


@pytest.mark.parametrize("name", ["bst_T3_smoke", "bst_T51_d8_h3", "bst_T20_d16_h2"])
def test_bst_golden(name):
    from recalgorithm_b200 import ops
    g = golden(name)
    p = {k[2:]: g[k] for k in g.files if k.startswith("p_")}
    B, T, d = g["x"].shape
    H, maxlen = int(g["heads"]), int(g["max_length"])
    x = dev(g["x"])
    out = ops.bst_transformer_fwd(x, x, x, dev(g["keys_length"]), _pack(p, d, H, maxlen), H, maxlen)
    assert_close(out, g["out_f32"], TOL, "vs transformer_layer.py executed in float32")
    assert_close(out, O.bst_transformer_fwd(g["x"], g["x"], g["x"], g["keys_length"], p, H), TOL, "vs oracle")


@pytest.mark.parametrize("B,T,d,H,maxlen,pos", [(5, 51, 8, 3, 51, True), (3, 7, 4, 1, 9, True), (4, 20, 16, 2, 20, False),
                                                 (2, 33, 32, 2, 40, True), (9, 1, 8, 3, 1, True), (3, 64, 8, 4, 64, True)])
def test_bst_fwd_bwd(B, T, d, H, maxlen, pos):
    from recalgorithm_b200 import ops
    rng = np.random.default_rng(B + T + d + H)
    q, k, v = (trunc_normal(rng, (B, T, d), 1.0) for _ in range(3))            # distinct queries / keys / values
    lengths = rng.integers(0, T + 1, size=B).astype(np.int64)
    shapes = O.bst_param_shapes(d, H, maxlen)
    p = {n: trunc_normal(rng, s, 0.4) for n, s in shapes.items()}
    p["ln1_gamma"] = 1 + p["ln1_gamma"]; p["ln2_gamma"] = 1 + p["ln2_gamma"]
    g = trunc_normal(rng, (B, T, d), 1.0)
    want, dq, dk, dv, dp = bst_torch.bst_transformer_bwd(q, k, v, lengths, p, H, g, use_position_embedding=pos)
    packed = _pack(p, d, H, maxlen)
    out = ops.bst_transformer_fwd(dev(q), dev(k), dev(v), dev(lengths), packed, H, maxlen, pos)
    assert_close(out, want, TOL, "fwd")
    gq, gk, gv, gp = ops.bst_transformer_bwd(dev(q), dev(k), dev(v), dev(lengths), packed, dev(g), H, maxlen, pos)
    assert_close(gq, dq, TOL, "d_queries"); assert_close(gk, dk, TOL, "d_keys"); assert_close(gv, dv, TOL, "d_values")
    got = ops.bst_unpack_params(gp, d, H, maxlen)
    for n in O.BST_PARAM_ORDER:
        if np.abs(dp[n]).max() == 0:
            assert float(got[n].abs().max()) == 0.0, n
        else:
            assert_close(got[n], dp[n], TOL, f"d_{n}")


def test_bst_layers_api_and_errors():
    from recalgorithm_b200 import _lib, layers as L, ops
    store = L.set_default_store(L.VariableStore(device="cuda", seed=5))
    x = torch.randn((3, 6, 8), device="cuda", requires_grad=True)
    klen = torch.tensor([6, 2, 0], device="cuda")
    with L.variable_scope("transformer_part"):
        y = L.bst_transformer(queries=x, keys=x, values=x, keys_length=klen, heads=3, index=0, max_length=6)
        y = L.bst_transformer(queries=y, keys=y, values=y, keys_length=klen, heads=3, index=1, max_length=6)
    names = set(store.vars)
    assert {"transformer_part/position_embedding", "transformer_part/w_q_0", "transformer_part/w_o_1", "transformer_part/LayerNorm/beta",
            "transformer_part/LayerNorm_1/gamma", "transformer_part/LayerNorm_2/beta", "transformer_part/LayerNorm_3/gamma",
            "transformer_part/dense/kernel", "transformer_part/dense_1/bias"} <= names
    y.sum().backward()
    assert x.grad is not None and store.vars["transformer_part/position_embedding"].grad is not None
    with pytest.raises(_lib.CtrError):           # position table shorter than the sequence
        ops.bst_transformer_fwd(x.detach(), x.detach(), x.detach(), klen, torch.zeros(int(_lib.lib().ctr_bst_param_count(8, 3, 4)), device="cuda"), 3, 4)
